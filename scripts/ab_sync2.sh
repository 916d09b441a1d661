#!/bin/bash
# A/B of the barrier-every-second-step aggregation (libs2pb200.so) against the per-step barrier (libs2pb200_base.so)
timeout 500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_distances.py -m gpu -q -x 2>&1 | tail -2
timeout 400 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "test_aggregate or test_tiny" 2>&1 | grep -E "passed|failed|RACECHECK|Race|hazard" | head -8
for lib in s2p_b200/libs2pb200_base.so s2p_b200/libs2pb200.so; do
  export S2PB200_LIB=$PWD/$lib
  timeout 250 python bench.py --no-cpu --steps 6 --warmup 3 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d['roofline']
print('$lib', d['config']['labels'], 'value %.1f e2e %.1f agg %.3f ms frac %.3f stage' % (d['value'], d['e2e']['value'], r['kernel_ms'], r['frac']), {k: round(v, 2) for k, v in r['stage_ms'].items()})"
  timeout 250 python bench.py --no-cpu --steps 3 --warmup 3 --slots 4 --tiles 8 --dmin -128 --dmax 127 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d['roofline']
print('$lib', d['config']['labels'], 'value %.1f e2e %.1f agg %.3f ms frac %.3f' % (d['value'], d['e2e']['value'], r['kernel_ms'], r['frac']))"
done
