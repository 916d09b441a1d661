#!/bin/bash
# One GPU session: tests, launch list, ncu captures of the three heavy kernels, bench line.
set -x
mkdir -p gpurun_out
export PARITY=0
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python scripts/c2_probe.py > /dev/null 2>&1
for k in aggregate wta cost_kernel; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -o gpurun_out/ncu_$k -f python scripts/c2_probe.py > gpurun_out/ncu_$k.log 2>&1
done
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 3000 gpurun_out/bench.json; tail -c 600 gpurun_out/bench_ref.json; tail -5 gpurun_out/bench.err
