#!/bin/bash
# strip-staged cost kernel: suite, stage times, in-flight A/B
O=gpurun_out/r02s29; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x --timeout 180 2>&1 | tail -8 > $O/tests.log; tail -3 $O/tests.log
for v in 0 1 0 1; do
S2PB_COST_STRIP=$v PARITY=0 timeout 120 python scripts/c2_probe.py 2>&1 | grep "iter 3" | sed "s/^/strip=$v /"
S2PB_COST_STRIP=$v timeout 300 python bench.py --no-cpu --no-extra --steps 8 --warmup 3 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d['roofline']
print('strip=$v value %.1f e2e %.1f agg %.3f ms verified %s' % (d['value'], d['e2e']['value'], r['kernel_ms'], d['outputs_verified']))"
done | tee $O/ab.txt
