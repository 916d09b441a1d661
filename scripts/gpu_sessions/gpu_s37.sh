#!/bin/bash
# final build: one more fuzz pass (new seed) after the FP64 product rewrite
O=gpurun_out/r02s37; mkdir -p $O
timeout 400 python scripts/fuzz_gpu.py 150 11 2>&1 | tail -5 > $O/fuzz.log; cat $O/fuzz.log
