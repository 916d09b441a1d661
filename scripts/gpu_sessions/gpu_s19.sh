#!/bin/bash
O=gpurun_out/r02s19; mkdir -p $O
PARITY=0 timeout 200 python scripts/c2_probe.py 2>&1 | grep "iter 3" | tee $O/standalone.txt
for k in 1 2; do
timeout 300 python bench.py --no-cpu --no-extra --steps 10 --warmup 4 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d['roofline']
print('value %.1f e2e %.1f agg %.3f ms verified %s' % (d['value'], d['e2e']['value'], r['kernel_ms'], d['outputs_verified']))"
done | tee $O/ab.txt
timeout 900 python -m pytest tests -m gpu -q -x --timeout 180 2>&1 | tail -15 > $O/tests.log; tail -3 $O/tests.log
