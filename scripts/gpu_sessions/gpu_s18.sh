#!/bin/bash
O=gpurun_out/r02s18; mkdir -p $O
for v in 0 1; do S2PB_WTA_STAGED=$v PARITY=0 timeout 200 python scripts/c2_probe.py 2>&1 | grep "iter 3" | sed "s/^/staged=$v /"; done | tee $O/standalone.txt
for g in 3 6 24; do
S2PB_WTA_STAGED=1 S2PB_WTA_GRID=$g timeout 300 python bench.py --no-cpu --no-extra --steps 8 --warmup 3 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d['roofline']
print('staged grid x$g value %.1f e2e %.1f agg %.3f ms verified %s' % (d['value'], d['e2e']['value'], r['kernel_ms'], d['outputs_verified']))"
done | tee $O/ab.txt
