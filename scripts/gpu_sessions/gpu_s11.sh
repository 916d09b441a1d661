#!/bin/bash
O=gpurun_out/r02s11; mkdir -p $O
export PARITY=0
timeout 600 python -m pytest tests -m gpu -q --timeout 120 2>&1 | tail -4 > $O/tests.log; tail -2 $O/tests.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:aggregate_kernel -s 2 -c 1 -o /tmp/ncu_agg -f python scripts/c2_probe.py > /tmp/ncu_agg.log 2>&1
ncu -i /tmp/ncu_agg.ncu-rep --page source --csv > /tmp/agg_src.csv 2>/dev/null && python scripts/ncu_opcodes.py /tmp/agg_src.csv 45 > $O/ncu_aggregate_opcodes.txt 2>&1; cat $O/ncu_aggregate_opcodes.txt
timeout 500 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x \
    -k "nodata_matches or ragged or exact_zero" 2>&1 | grep -E "passed|failed|RACECHECK SUMMARY|Race reported|Error" | cut -c1-220 | head -12 > $O/racecheck.log; cat $O/racecheck.log
