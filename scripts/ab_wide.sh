#!/bin/bash
# A/B of the two-CTA configuration for 5..8 labels per lane: bench lines at D = 192 and 256 and the C3-like mgm_multi tile
for lib in s2p_b200/libs2pb200_base.so s2p_b200/libs2pb200.so; do
  export S2PB200_LIB=$PWD/$lib
  for r in "-96 95" "-128 127"; do
    set -- $r
    timeout 250 python bench.py --no-cpu --steps 3 --warmup 3 --slots 4 --tiles 8 --dmin $1 --dmax $2 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d['roofline']
print('$lib', d['config']['labels'], 'value %.1f e2e %.1f agg %.2f ms frac %.3f stage' % (d['value'], d['e2e']['value'], r['kernel_ms'], r['frac']), {k: round(v, 2) for k, v in r['stage_ms'].items()})"
  done
  BIG=1 WATCHDOG=200 timeout 250 python scripts/multi_probe.py 2>&1 | grep -A1 "(532, 768)" | tail -1
done
