#!/bin/bash
O=gpurun_out/r02s5; mkdir -p $O
export PARITY=0
timeout 900 python -m pytest tests -m gpu -q --timeout 120 2>&1 | tail -25 > $O/tests.log; tail -8 $O/tests.log
summ() { python - "$1" <<'P'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print("value %.1f e2e %.1f agg %.3f verified %s" % (d["value"], d["e2e"]["value"], d["roofline"]["kernel_ms"], d.get("outputs_verified")))
except Exception as e:
    print("bench line unreadable:", e)
P
}
S2PB200_LIB=$PWD/s2p_b200/libs2pb200_tma.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 120 2>&1 | tail -6 > $O/tests_tma.log; tail -3 $O/tests_tma.log
S2PB200_LIB=$PWD/s2p_b200/libs2pb200_tma.so timeout 400 python bench.py --no-cpu --no-extra --steps 6 --warmup 3 > $O/bench_tma.json 2> $O/bench_tma.err; tail -c 300 $O/bench_tma.err; summ $O/bench_tma.json
S2PB200_LIB=$PWD/s2p_b200/libs2pb200_tma.so timeout 300 ncu --set full --clock-control none --import-source on -k regex:aggregate -s 2 -c 1 -o /tmp/ncu_agg_tma -f python scripts/c2_probe.py > /tmp/ncu_agg_tma.log 2>&1
python scripts/ncu_summary.py /tmp/ncu_agg_tma.ncu-rep $O/ncu_aggregate_tma_selfissue.txt > /dev/null 2>&1
ncu -i /tmp/ncu_agg_tma.ncu-rep --page source --csv > /tmp/agg_src.csv 2>/dev/null && python scripts/ncu_hot.py /tmp/agg_src.csv 25 > $O/ncu_aggregate_tma_selfissue_hot_sass.txt 2>&1
head -32 $O/ncu_aggregate_tma_selfissue.txt | tail -24
