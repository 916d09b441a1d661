#!/bin/bash
# sanitizers and fuzz after the WTA / cost / two-aggregation changes and the threaded mgm_multi batch
O=gpurun_out/r02s30; mkdir -p $O
timeout 500 python scripts/fuzz_gpu.py 120 7 2>&1 | tail -4 > $O/fuzz.log; cat $O/fuzz.log
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -m gpu -q -x \
    -k "costvolume or census or end_to_end or batch or eight_tiles or nodata_matches or options" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|error" | head -8 > $O/memcheck.log; cat $O/memcheck.log
timeout 500 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x \
    -k "costvolume or eight_tiles or multi_batch" 2>&1 | grep -E "passed|failed|RACECHECK SUMMARY|Race reported|Error" | cut -c1-220 | head -12 > $O/racecheck.log; cat $O/racecheck.log
