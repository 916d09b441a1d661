#!/bin/bash
O=gpurun_out/r02s36; mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2.json 2> $O/bench_n2.err; tail -c 400 $O/bench_n2.err
python - <<'P'
import json
try:
    d = json.loads(open("gpurun_out/r02s36/bench_n2.json").read().strip().splitlines()[-1])
    print("N=2 value %.1f e2e %.1f verified %s" % (d["value"], d["e2e"]["value"], d.get("outputs_verified")))
    for k, v in d.get("extra_configs", {}).items():
        print(k, {x: (round(v[x], 2) if isinstance(v[x], float) else v[x]) for x in ("value", "error", "seconds", "tiles_taken_by_rank0", "scaling") if x in v})
except Exception as e:
    print("bench line unreadable:", e)
P
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 --no-extra > $O/bench_ref_n2.json 2> $O/bench_ref_n2.err; tail -c 300 $O/bench_ref_n2.json
