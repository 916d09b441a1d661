#!/bin/bash
O=gpurun_out/r02s10; mkdir -p $O
export PARITY=0
timeout 900 python scripts/fuzz_gpu.py 150 11 > $O/fuzz.txt 2>&1; tail -6 $O/fuzz.txt | cut -c1-300
# sanitizers on the kernels added this round
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_parity.py tests/test_gpu_homography.py tests/test_triangulation.py -m gpu -q -x \
    -k "nodata_matches or exact_zero or ragged or wider or more_than_512 or pkr or fused or real_rpc or test_mgm_multi" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|error" | head -8 > $O/memcheck.log; cat $O/memcheck.log
timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x \
    -k "nodata_matches or ragged or more_than_512 or eight_tiles" 2>&1 | grep -E "passed|failed|RACECHECK SUMMARY|Race reported|Error" | cut -c1-220 | head -12 > $O/racecheck.log; cat $O/racecheck.log
