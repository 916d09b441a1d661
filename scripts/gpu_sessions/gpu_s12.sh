#!/bin/bash
O=gpurun_out/r02s12; mkdir -p $O
export PARITY=0
for lib in s2p_b200/libs2pb200.so s2p_b200/libs2pb200_sfa.so s2p_b200/libs2pb200_pix.so s2p_b200/libs2pb200_both.so s2p_b200/libs2pb200.so; do
  S2PB200_LIB=$PWD/$lib timeout 250 python bench.py --no-cpu --no-extra --steps 8 --warmup 3 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d['roofline']
print('$lib', 'value %.1f e2e %.1f agg %.3f ms verified %s' % (d['value'], d['e2e']['value'], r['kernel_ms'], d['outputs_verified']))"
done | tee $O/ab.txt
S2PB200_LIB=$PWD/s2p_b200/libs2pb200_both.so timeout 600 python -m pytest tests -m gpu -q --timeout 120 2>&1 | tail -3 > $O/tests_both.log; tail -2 $O/tests_both.log
