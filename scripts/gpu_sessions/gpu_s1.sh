#!/bin/bash
# round-2 GPU session 1: baseline tests, first contact of the emulator-fixed chunked path, no-data cliff, A/B knobs
O=gpurun_out/r02s1; mkdir -p $O
export PARITY=0
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -4 > $O/tests.log; tail -1 $O/tests.log
for m in 0 1; do
  S2PB_CHUNKED=$m S2PB_CHUNKED_MIN_DP=0 timeout 300 python scripts/chunked_probe.py > $O/chunked_small_$m.txt 2>&1; cat $O/chunked_small_$m.txt
done
for m in 0 1 2 3; do
  S2PB_CHUNKED=$m S2PB_CHUNKED_MIN_DP=0 BIG=only COMPARE=/tmp/x timeout 200 python scripts/chunked_probe.py > $O/chunked_big_$m.txt 2>&1; cat $O/chunked_big_$m.txt
done
S2PB_CHUNKED=1 S2PB_CHUNKED_MIN_DP=0 timeout 400 python scripts/fuzz_gpu.py 60 7 > $O/fuzz_chunked.txt 2>&1; tail -3 $O/fuzz_chunked.txt
timeout 300 python bench.py --no-cpu --steps 6 --warmup 3 --nan-border 0.05 > $O/bench_nan.json 2> $O/bench_nan.err; tail -c 600 $O/bench_nan.json
bash scripts/ab_variants.sh run > $O/ab.txt 2>&1; cat $O/ab.txt
