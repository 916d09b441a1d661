#!/bin/bash
# round-2 GPU session 2: full GPU suite (DCT round trip, split loop, new fixtures), the extended bench line, ncu of the new aggregation
O=gpurun_out/r02s2; mkdir -p $O
export PARITY=0
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > $O/tests.log; tail -6 $O/tests.log
timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench_n1.json 2> $O/bench.err; tail -c 300 $O/bench.err; python - <<'P'
import json
try:
    d = json.load(open("gpurun_out/r02s2/bench_n1.json"))
    print("value %.1f e2e %.1f agg %.3f verified %s" % (d["value"], d["e2e"]["value"], d["roofline"]["kernel_ms"], d.get("outputs_verified")))
    for k, v in d.get("extra_configs", {}).items():
        print(k, {x: (round(v[x], 2) if isinstance(v[x], float) else v[x]) for x in ("value", "error", "ms_per_tile", "seconds", "ms_per_warp", "fusion_ms_per_tile") if x in v},
              "e2e", round(v.get("e2e", {}).get("value", 0), 1), "cpu", round((v.get("cpu_baseline") or {}).get("value", 0), 4))
except Exception as e:
    print("bench line unreadable:", e)
P
# launch list + full capture of the aggregation kernel of one C2 tile (with 5 % no-data so that the DCT kernels show up)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_c2_probe.csv python scripts/c2_probe.py > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:aggregate -s 2 -c 1 -o /tmp/ncu_agg -f python scripts/c2_probe.py > /tmp/ncu_agg.log 2>&1
python scripts/ncu_summary.py /tmp/ncu_agg.ncu-rep $O/ncu_aggregate.txt > /dev/null 2>&1
ncu -i /tmp/ncu_agg.ncu-rep --page source --csv > /tmp/agg_src.csv 2>/dev/null && python scripts/ncu_hot.py /tmp/agg_src.csv 25 > $O/ncu_aggregate_hot_sass.txt 2>&1
NANB=0.05 timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread --clock-control none --csv --log-file $O/launches_c2_nodata.csv python scripts/c2_probe.py > /dev/null 2>&1
timeout 400 python bench.py --impl reference --steps 1 --warmup 0 > $O/bench_reference_n1.json 2> $O/bench_ref.err; tail -c 700 $O/bench_reference_n1.json
head -30 $O/ncu_aggregate.txt
