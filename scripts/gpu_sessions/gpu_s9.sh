#!/bin/bash
O=gpurun_out/r02s9; mkdir -p $O
export PARITY=0
S2PB_CHUNKED=2 S2PB_CHUNKED_MIN_DP=300 BIG=only COMPARE=/tmp/x timeout 300 ncu --set full --clock-control none --import-source on -k regex:aggregate_chunked -s 1 -c 1 -o /tmp/ncu_ck -f python scripts/chunked_probe.py > /tmp/ncu_ck.log 2>&1; tail -3 /tmp/ncu_ck.log
python scripts/ncu_summary.py /tmp/ncu_ck.ncu-rep $O/ncu_chunked.txt > /dev/null 2>&1
ncu -i /tmp/ncu_ck.ncu-rep --page source --csv > /tmp/ck_src.csv 2>/dev/null && python scripts/ncu_hot.py /tmp/ck_src.csv 40 > $O/ncu_chunked_hot_sass.txt 2>&1
head -34 $O/ncu_chunked.txt | tail -30; head -48 $O/ncu_chunked_hot_sass.txt | cut -c1-190
timeout 300 python bench.py --no-cpu --steps 5 --warmup 3 --only-extra rectification_warp > $O/bench_warp.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r02s9/bench_warp.json')); print(d['value'], {k:(v.get('value'),v.get('ms_per_warp')) for k,v in d.get('extra_configs',{}).items()})" 2>&1 | cut -c1-300
