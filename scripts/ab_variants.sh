#!/bin/bash
# A/B of build-time variants of the aggregation kernel.
#   bash scripts/ab_variants.sh build      (in the CPU container: builds s2p_b200/libs2pb200_<name>.so for every variant)
#   bash scripts/ab_variants.sh run        (on the GPU box: one bench line per library, default first)
# Remove the variant libraries afterwards (they travel with every gpurun snapshot): bash scripts/ab_variants.sh clean
VARIANTS="split:-DS2PB_SPLIT_LOOP=1 pub16:-DS2PB_PUBLISH=16 pub32:-DS2PB_PUBLISH=32 stream:-DS2PB_STREAM_STORES=1"
case "$1" in
build)
  for v in $VARIANTS; do
    name=${v%%:*}; flag=${v#*:}
    make -C s2p_b200/csrc -j10 EXTRA=$flag OBJDIR=build_$name OUT=../libs2pb200_$name.so 2>&1 | grep -E "error|warning"
    ls -la s2p_b200/libs2pb200_$name.so
  done ;;
run)
  for lib in s2p_b200/libs2pb200.so s2p_b200/libs2pb200_*.so; do
    S2PB200_LIB=$PWD/$lib timeout 250 python bench.py --no-cpu --steps 6 --warmup 3 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d['roofline']
print('$lib', 'value %.1f e2e %.1f agg %.3f ms' % (d['value'], d['e2e']['value'], r['kernel_ms']))"
  done ;;
clean)
  rm -rf s2p_b200/csrc/build_* s2p_b200/libs2pb200_*.so ;;
*) echo "usage: $0 build|run|clean" ;;
esac
