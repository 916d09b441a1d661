#!/bin/bash
# session 17: WTA with shared-memory prefetch (128-slot slabs) -- suite, then A/B with tiles in flight
O=gpurun_out/r02s17; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x --timeout 180 2>&1 | tail -15 > $O/tests.log; tail -4 $O/tests.log
for v in 0 1 0 1; do
S2PB_WTA_STAGED=$v timeout 300 python bench.py --no-cpu --no-extra --steps 10 --warmup 4 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d['roofline']
print('staged=$v value %.1f e2e %.1f agg %.3f ms verified %s' % (d['value'], d['e2e']['value'], r['kernel_ms'], d['outputs_verified']), d.get('stage_ms'))"
done | tee $O/ab.txt
