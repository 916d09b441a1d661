#!/bin/bash
# timing diagnostics (wrong results on purpose): how much of the aggregation is the band hand-off poll / the per-step CTA barrier
O=gpurun_out/r02s25; mkdir -p $O
for lib in libs2pb200.so libs2pb200_np.so libs2pb200_nb.so libs2pb200_nbp.so; do
S2PB200_LIB=$PWD/s2p_b200/$lib PARITY=0 timeout 120 python scripts/c2_probe.py 2>&1 | grep "iter 3" | sed "s/^/$lib /"
done | tee $O/diag.txt
