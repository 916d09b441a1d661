#!/bin/bash
O=gpurun_out/r02s3; mkdir -p $O
export PARITY=0
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > $O/tests.log; tail -6 $O/tests.log
timeout 600 python bench.py --no-cpu --steps 5 --warmup 3 > $O/bench_n1.json 2> $O/bench.err; tail -c 300 $O/bench.err; python - <<'P'
import json
try:
    d = json.load(open("gpurun_out/r02s3/bench_n1.json"))
    print("value %.1f e2e %.1f agg %.3f verified %s" % (d["value"], d["e2e"]["value"], d["roofline"]["kernel_ms"], d.get("outputs_verified")))
    for k, v in d.get("extra_configs", {}).items():
        print(k, {x: (round(v[x], 2) if isinstance(v[x], float) else v[x]) for x in ("value", "error", "ms_per_tile", "seconds", "ms_per_warp", "fusion_ms_per_tile") if x in v},
              "e2e", round(v.get("e2e", {}).get("value", 0), 1), (v.get("roofline") or {}).get("stage_ms"))
except Exception as e:
    print("bench line unreadable:", e)
P
NANB=0.05 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_c2_nodata.csv python scripts/c2_probe.py > /dev/null 2>&1
python - <<'P'
import csv, collections
rows = list(csv.reader(open("gpurun_out/r02s3/launches_c2_nodata.csv")))
hdr = None; agg = collections.OrderedDict()
for r in rows:
    if "Kernel Name" in r: hdr = r; continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r)); agg.setdefault(d["Kernel Name"][:48], []).append(float(d["Metric Value"].replace(",", "")))
for k, v in agg.items(): print("%-50s n=%2d avg %9.1f us" % (k, len(v), sum(v) / len(v) / 1e3))
P
