#!/bin/bash
# session 23: evidence after the WTA series -- suite, every kernel in one ncu pass, full captures of the aggregation and the WTA,
# launch list of one C2 tile, the default bench line and the reference arm
O=gpurun_out/r02s35; mkdir -p $O
export PARITY=0
timeout 900 python -m pytest tests -m gpu -q --timeout 180 2>&1 | tail -6 > $O/tests.log; tail -2 $O/tests.log
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size \
    --clock-control none -o /tmp/ncu_all -f python scripts/all_kernels_probe.py > /tmp/ncu_all.log 2>&1
ncu -i /tmp/ncu_all.ncu-rep --page raw --csv 2>/dev/null | python scripts/kernel_table.py > $O/all_kernels_table.md 2> $O/all_kernels_table.err; head -16 $O/all_kernels_table.md | cut -c1-180
timeout 300 ncu --set full --clock-control none --import-source on -k regex:aggregate_kernel -s 2 -c 1 -o /tmp/ncu_agg -f python scripts/c2_probe.py > /tmp/ncu_agg.log 2>&1
python scripts/ncu_summary.py /tmp/ncu_agg.ncu-rep $O/ncu_aggregate.txt > /dev/null 2>&1
ncu -i /tmp/ncu_agg.ncu-rep --page source --csv > /tmp/agg_src.csv 2>/dev/null && python scripts/ncu_hot.py /tmp/agg_src.csv 40 > $O/ncu_aggregate_hot_sass.txt 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wta_kernel -s 4 -c 1 -o /tmp/ncu_wta -f python scripts/c2_probe.py > /tmp/ncu_wta.log 2>&1
python scripts/ncu_summary.py /tmp/ncu_wta.ncu-rep $O/ncu_wta.txt > /dev/null 2>&1
head -30 $O/ncu_aggregate.txt | cut -c1-150; head -30 $O/ncu_wta.txt | cut -c1-150
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/launches_c2_probe.csv python scripts/c2_probe.py > /dev/null 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench.err; tail -c 300 $O/bench.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference_n1.json 2> $O/bench_ref.err
python - <<'P'
import json
d = json.load(open("gpurun_out/r02s35/bench_n1.json"))
print("value %.1f e2e %.1f agg %.3f verified %s cpu %s launches %s" % (d["value"], d["e2e"]["value"], d["roofline"]["kernel_ms"], d.get("outputs_verified"), d["cpu_baseline"]["value"], d["gpu_launches"]))
for k, v in d.get("extra_configs", {}).items():
    print(k, {x: (round(v[x], 2) if isinstance(v[x], float) else v[x]) for x in ("value", "error", "ms_per_tile", "seconds", "ms_per_warp") if x in v}, "e2e", round(v["e2e"]["value"], 1), "cpu", (v.get("cpu_baseline") or {}).get("value"))
r = json.load(open("gpurun_out/r02s35/bench_reference_n1.json")); print("reference", r.get("value"), r.get("cpu_baseline", {}).get("sample", "")[:120])
P
