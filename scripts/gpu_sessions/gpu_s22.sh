#!/bin/bash
O=gpurun_out/r02s22; mkdir -p $O
for lib in libs2pb200.so libs2pb200_g8.so libs2pb200_t96.so libs2pb200.so; do
S2PB200_LIB=$PWD/s2p_b200/$lib PARITY=0 timeout 200 python scripts/c2_probe.py 2>&1 | grep "iter 3" | sed "s/^/$lib /"
S2PB200_LIB=$PWD/s2p_b200/$lib timeout 300 python bench.py --no-cpu --no-extra --steps 10 --warmup 4 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d['roofline']
print('$lib value %.1f e2e %.1f agg %.3f ms verified %s' % (d['value'], d['e2e']['value'], r['kernel_ms'], d['outputs_verified']))"
done | tee $O/ab.txt
