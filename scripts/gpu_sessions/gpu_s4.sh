#!/bin/bash
O=gpurun_out/r02s4; mkdir -p $O
export PARITY=0
timeout 600 python -m pytest tests -m gpu -q -k "wider or more_than_512 or test_mgm_multi" 2>&1 | tail -12 > $O/tests_wide.log; tail -4 $O/tests_wide.log
summ() { python - "$1" <<'P'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print("value %.1f e2e %.1f agg %.3f verified %s" % (d["value"], d["e2e"]["value"], d["roofline"]["kernel_ms"], d.get("outputs_verified")))
    for k, v in d.get("extra_configs", {}).items():
        print(" ", k, {x: (round(v[x], 2) if isinstance(v[x], float) else v[x]) for x in ("value", "error", "ms_per_tile", "seconds") if x in v}, (v.get("roofline") or {}).get("kernel_ms"))
except Exception as e:
    print("bench line unreadable:", e)
P
}
# TMA staging variant: full GPU suite (bit-exactness, no deadlock: each test runs under the suite timeout) then the headline bench
S2PB200_LIB=$PWD/s2p_b200/libs2pb200_tma.so timeout 900 python -m pytest tests -m gpu -q -x --timeout 120 2>&1 | tail -12 > $O/tests_tma.log; tail -4 $O/tests_tma.log
S2PB200_LIB=$PWD/s2p_b200/libs2pb200_tma.so timeout 400 python bench.py --no-cpu --steps 6 --warmup 3 --only-extra C2_nodata_5pct,C4_tile_queue_1026x1026x192_strong > $O/bench_tma.json 2> $O/bench_tma.err; tail -c 300 $O/bench_tma.err; summ $O/bench_tma.json
timeout 400 python bench.py --no-cpu --steps 6 --warmup 3 --only-extra C2_nodata_5pct,C4_tile_queue_1026x1026x192_strong > $O/bench_default.json 2> $O/bench_default.err; summ $O/bench_default.json
S2PB200_LIB=$PWD/s2p_b200/libs2pb200_wide2.so timeout 400 python bench.py --no-cpu --steps 6 --warmup 3 --only-extra C2_nodata_5pct,C4_tile_queue_1026x1026x192_strong > $O/bench_wide2.json 2> $O/bench_wide2.err; summ $O/bench_wide2.json
# ncu of the TMA variant's aggregation
S2PB200_LIB=$PWD/s2p_b200/libs2pb200_tma.so timeout 300 ncu --set full --clock-control none --import-source on -k regex:aggregate -s 2 -c 1 -o /tmp/ncu_agg_tma -f python scripts/c2_probe.py > /tmp/ncu_agg_tma.log 2>&1
python scripts/ncu_summary.py /tmp/ncu_agg_tma.ncu-rep $O/ncu_aggregate_tma.txt > /dev/null 2>&1
ncu -i /tmp/ncu_agg_tma.ncu-rep --page source --csv > /tmp/agg_src.csv 2>/dev/null && python scripts/ncu_hot.py /tmp/agg_src.csv 25 > $O/ncu_aggregate_tma_hot_sass.txt 2>&1
head -32 $O/ncu_aggregate_tma.txt | tail -22
