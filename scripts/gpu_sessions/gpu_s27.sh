#!/bin/bash
# A/B: one persistent aggregation CTA per SM, so that two tiles' aggregations share the SMs (half as many bands deep per pass)
O=gpurun_out/r02s27; mkdir -p $O
for v in 0 1 0 1; do
S2PB_AGG_CTAS=$v timeout 300 python bench.py --no-cpu --no-extra --steps 8 --warmup 3 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d['roofline']
print('agg ctas/SM cap $v: value %.1f e2e %.1f agg %.3f ms verified %s' % (d['value'], d['e2e']['value'], r['kernel_ms'], d['outputs_verified']))"
done | tee $O/ab.txt
