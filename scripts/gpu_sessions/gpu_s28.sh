#!/bin/bash
# A/B: how many tiles' aggregations share the GPU (persistent CTAs per launch: 296 / 148 / 74 / 49)
O=gpurun_out/r02s28; mkdir -p $O
for v in 1 -2 -3 1; do
S2PB_AGG_CTAS=$v timeout 300 python bench.py --no-cpu --no-extra --steps 8 --warmup 3 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d['roofline']
print('agg cap $v: value %.1f e2e %.1f agg %.3f ms verified %s' % (d['value'], d['e2e']['value'], r['kernel_ms'], d['outputs_verified']))"
done | tee $O/ab.txt
S2PB_AGG_CTAS=1 timeout 300 python bench.py --no-cpu --only-extra C2_nodata_5pct --steps 4 --warmup 3 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read())
for k, e in d.get('extra_configs', {}).items(): print('cap 1', k, e.get('value'), 'e2e', e['e2e']['value'])" | tee -a $O/ab.txt
