#!/bin/bash
# FP64 GEMMs of the DCT round trip with 4x4 register tiles: suite, no-data stage times and config, C3 tile
O=gpurun_out/r02s32; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x --timeout 180 2>&1 | tail -8 > $O/tests.log; tail -3 $O/tests.log
NANB=0.05 PARITY=0 timeout 120 python scripts/c2_probe.py 2>&1 | grep "iter 3" | tee $O/nodata_probe.txt
timeout 120 python scripts/c3_probe.py 3 2>&1 | grep "tile 3" | tee -a $O/nodata_probe.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"dct_gemm|rt_inverse" --csv --log-file $O/dct_launches.csv python scripts/c3_probe.py 0 > /dev/null 2>&1
NANB=0.05 PARITY=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"dct_gemm|rt_inverse" -c 8 --csv --log-file $O/dct_launches_nodata.csv python scripts/c2_probe.py > /dev/null 2>&1
python - <<'P'
import csv
for f in ('dct_launches.csv', 'dct_launches_nodata.csv'):
    rows = [r for r in csv.reader(open('gpurun_out/r02s32/' + f)) if len(r) > 10]
    hdr = rows[0]; ik = hdr.index('Kernel Name'); iv = hdr.index('Metric Value')
    for r in rows[1:9]: print(f, r[ik][:50], r[iv])
P
timeout 300 python bench.py --no-cpu --only-extra C2_nodata_5pct --steps 6 --warmup 3 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('C2 value %.1f e2e %.1f' % (d['value'], d['e2e']['value']))
for k, e in d.get('extra_configs', {}).items(): print(k, e.get('value'), 'e2e', e['e2e']['value'])" | tee $O/nodata.txt
