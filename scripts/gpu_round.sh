#!/bin/bash
# One GPU session for the round's evidence: tests, sanitizers, launch list, ncu captures of the three heavy kernels
# (summarised on the box; only text comes back), a metrics pass over every kernel of the library, both bench arms.
# usage (from the repo root, on the GPU box):  bash scripts/gpu_round.sh [tag]
TAG=${1:-r01}
O=gpurun_out/$TAG
mkdir -p $O
export PARITY=0
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -4 > $O/tests.log; tail -1 $O/tests.log
# sanitizers: memcheck on the kernels added this round, racecheck on the aggregation (incl. the 5..8 labels-per-lane path)
timeout 500 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_distances.py tests/test_gpu_homography.py -m gpu -q -x \
    -k "costvolume or aggregate_general or mgm_weighted or mgm_distances or homography" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|error" | head -6 > $O/memcheck.log; cat $O/memcheck.log
timeout 500 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x \
    -k "test_aggregate or test_mgm_end_to_end" 2>&1 | grep -E "passed|failed|RACECHECK SUMMARY|Race reported|Error" | cut -c1-220 | head -12 > $O/racecheck.log; cat $O/racecheck.log
# launch list of one C2 tile
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_c2_probe.csv python scripts/c2_probe.py > /dev/null 2>&1
# full captures, summarised here
for k in aggregate wta cost_kernel; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -o /tmp/ncu_$k -f python scripts/c2_probe.py > /tmp/ncu_$k.log 2>&1
  python scripts/ncu_summary.py /tmp/ncu_$k.ncu-rep $O/ncu_$k.txt > /dev/null 2>&1
done
ncu -i /tmp/ncu_aggregate.ncu-rep --page source --csv > /tmp/agg_src.csv 2>/dev/null && python scripts/ncu_hot.py /tmp/agg_src.csv 25 > $O/ncu_aggregate_hot_sass.txt 2>&1
# every kernel of the library: time, DRAM bytes, issue activity
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size \
    --clock-control none -o /tmp/ncu_all -f python scripts/all_kernels_probe.py > /tmp/ncu_all.log 2>&1
ncu -i /tmp/ncu_all.ncu-rep --page raw --csv 2>/dev/null | python scripts/kernel_table.py > $O/all_kernels_table.md 2> $O/all_kernels_table.err
# both bench arms, the way the driver runs them
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference_n1.json 2> $O/bench_ref.err
timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench_n1.json 2> $O/bench.err
timeout 300 python bench.py --no-cpu --steps 4 --warmup 3 --slots 4 --tiles 8 --dmin -128 --dmax 127 > $O/bench_d256_n1.json 2>/dev/null
timeout 300 python bench.py --no-cpu --steps 4 --warmup 3 --slots 4 --tiles 8 --dmin -96 --dmax 95 > $O/bench_d192_n1.json 2>/dev/null
BIG=1 WATCHDOG=200 timeout 250 python scripts/multi_probe.py 2>&1 | grep -A1 "(532, 768)" | tail -1 > $O/mgm_multi_c3.txt
timeout 100 python scripts/hom_probe.py > $O/homography.txt 2>&1
timeout 100 python scripts/general_probe.py > $O/general_flavour.txt 2>&1
tail -c 1500 $O/bench_n1.json; echo; tail -c 500 $O/bench_reference_n1.json; echo; cat $O/mgm_multi_c3.txt $O/homography.txt; ls -la $O
