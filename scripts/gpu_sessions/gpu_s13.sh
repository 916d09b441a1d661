#!/bin/bash
# session 13: threaded mgm_multi batch — full GPU suite, C3 with 4 tiles in flight, then the evidence runs
O=gpurun_out/r02s13; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --timeout 180 2>&1 | tail -15 > $O/tests.log; tail -4 $O/tests.log
timeout 400 python bench.py --no-cpu --only-extra C3_mgm_multi_256 --steps 6 --warmup 3 2>$O/c3.err > $O/c3.json; tail -3 $O/c3.err
python - <<'P'
import json
d = json.load(open('gpurun_out/r02s13/c3.json'))
for e in d.get('extra_configs', []):
    print(json.dumps(e)[:900])
P
timeout 900 python bench.py --steps 20 --warmup 5 2>$O/bench.err > $O/bench_n1.json
python - <<'P'
import json
d = json.load(open('gpurun_out/r02s13/bench_n1.json')); r = d['roofline']
print('value %.1f e2e %.1f agg %.3f verified %s launches %s' % (d['value'], d['e2e']['value'], r['kernel_ms'], d['outputs_verified'], d['gpu_launches']))
for e in d.get('extra_configs', []):
    print(e.get('workload'), e.get('value'), 'e2e', (e.get('e2e') or {}).get('value') if isinstance(e.get('e2e'), dict) else e.get('e2e'))
print(d.get('cpu_baseline'), d.get('clocks'))
P
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>$O/ref.err > $O/bench_reference_n1.json; cat $O/bench_reference_n1.json | cut -c1-600
