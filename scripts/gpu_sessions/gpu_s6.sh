#!/bin/bash
O=gpurun_out/r02s6; mkdir -p $O
export PARITY=0
timeout 900 python -m pytest tests -m gpu -q --timeout 120 2>&1 | tail -25 > $O/tests.log; tail -6 $O/tests.log
timeout 600 python bench.py --no-cpu --steps 5 --warmup 3 > $O/bench_n1.json 2> $O/bench.err; tail -c 300 $O/bench.err; python - <<'P'
import json
try:
    d = json.load(open("gpurun_out/r02s6/bench_n1.json"))
    print("value %.1f e2e %.1f agg %.3f verified %s" % (d["value"], d["e2e"]["value"], d["roofline"]["kernel_ms"], d.get("outputs_verified")))
    for k, v in d.get("extra_configs", {}).items():
        print(k, {x: (round(v[x], 2) if isinstance(v[x], float) else v[x]) for x in ("value", "error", "ms_per_tile", "seconds", "ms_per_warp", "fusion_ms_per_tile") if x in v},
              "e2e", round(v.get("e2e", {}).get("value", 0), 1), (v.get("roofline") or {}).get("stage_ms"))
except Exception as e:
    print("bench line unreadable:", e)
P
NANB=0.05 timeout 300 ncu --set full --clock-control none --import-source on -k regex:"rt_inverse|dct_gemm" -s 4 -c 2 -o /tmp/ncu_dct -f python scripts/c2_probe.py > /tmp/ncu_dct.log 2>&1
python scripts/ncu_summary.py /tmp/ncu_dct.ncu-rep $O/ncu_dct.txt > /dev/null 2>&1; cat $O/ncu_dct.txt | head -70
ncu -i /tmp/ncu_dct.ncu-rep --page source --csv > /tmp/dct_src.csv 2>/dev/null && python scripts/ncu_hot.py /tmp/dct_src.csv 14 > $O/ncu_dct_hot_sass.txt 2>&1; head -20 $O/ncu_dct_hot_sass.txt | cut -c1-170
