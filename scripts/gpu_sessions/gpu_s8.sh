#!/bin/bash
# evidence for profiles/: every kernel of the library (one ncu metrics pass), the full default bench line with CPU baselines, the reference arm
O=gpurun_out/r02s8; mkdir -p $O
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size \
    --clock-control none -o /tmp/ncu_all -f python scripts/all_kernels_probe.py > /tmp/ncu_all.log 2>&1
ncu -i /tmp/ncu_all.ncu-rep --page raw --csv 2>/dev/null | python scripts/kernel_table.py > $O/all_kernels_table.md 2> $O/all_kernels_table.err; head -30 $O/all_kernels_table.md | cut -c1-200
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench.err; tail -c 300 $O/bench.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference_n1.json 2> $O/bench_ref.err; tail -c 400 $O/bench_reference_n1.json
python - <<'P'
import json
d = json.load(open("gpurun_out/r02s8/bench_n1.json"))
print("value %.1f e2e %.1f agg %.3f verified %s cpu %s" % (d["value"], d["e2e"]["value"], d["roofline"]["kernel_ms"], d.get("outputs_verified"), d["cpu_baseline"]["value"]))
for k, v in d.get("extra_configs", {}).items():
    print(k, {x: (round(v[x], 2) if isinstance(v[x], float) else v[x]) for x in ("value", "error", "ms_per_tile", "seconds", "ms_per_warp") if x in v}, "cpu", (v.get("cpu_baseline") or {}).get("value"))
P
# traffic of the aggregation kernel for bench.py's roofline.traffic (ncu --set full of one C2 tile)
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"aggregate|wta_kernel" -s 4 -c 3 --csv --log-file $O/traffic.csv python scripts/c2_probe.py > /dev/null 2>&1; cat $O/traffic.csv | tail -12 | cut -c1-200
