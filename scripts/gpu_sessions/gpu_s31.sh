#!/bin/bash
# A/B: deeper sharing (74 CTAs when >= N other aggregations are pending), workspaces in flight
O=gpurun_out/r02s31; mkdir -p $O
run() { timeout 300 python bench.py --no-cpu --no-extra --steps 8 --warmup 3 "$@" 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d['roofline']
print('value %.1f e2e %.1f agg %.3f ms verified %s' % (d['value'], d['e2e']['value'], r['kernel_ms'], d['outputs_verified']))"; }
( echo -n "default (8 slots): "; run
  echo -n "deep>=2: "; S2PB_AGG_DEEP=2 run
  echo -n "deep>=4: "; S2PB_AGG_DEEP=4 run
  echo -n "6 slots: "; run --slots 6
  echo -n "12 slots: "; run --slots 12
  echo -n "12 slots deep>=4: "; S2PB_AGG_DEEP=4 run --slots 12 ) | tee $O/ab.txt
