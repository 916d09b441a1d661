#!/bin/bash
# session 14: interleaved WTA for LPL 3/5/6 — suite, no-data config, C3 launch list
O=gpurun_out/r02s14; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --timeout 180 2>&1 | tail -15 > $O/tests.log; tail -4 $O/tests.log
timeout 400 python bench.py --no-cpu --only-extra C2_nodata_5pct --steps 6 --warmup 3 2>$O/nodata.err > $O/nodata.json
python - <<'P'
import json
d = json.load(open('gpurun_out/r02s14/nodata.json'))
for k, e in d.get('extra_configs', {}).items():
    print(k, e.get('value'), 'e2e', e['e2e']['value'], e.get('stage_ms'))
P
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_c3.csv python bench.py --no-cpu --only-extra C3_mgm_multi_256 --steps 1 --warmup 1 > $O/c3_ncu.log 2>&1
python - <<'P'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/r02s14/launches_c3.csv')) if len(r) > 10]
hdr = rows[0]; ik = hdr.index('Kernel Name'); iv = hdr.index('Metric Value'); iu = hdr.index('Metric Unit'); ig = hdr.index('Grid Size')
agg = collections.OrderedDict()
for r in rows[1:]:
    v = float(r[iv].replace(',', '')); u = r[iu]
    v = v / 1000 if u in ('ns', 'nsecond') else v * (1000 if u in ('ms', 'msecond') else 1)
    k = r[ik][:60]
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print('total %.1f us over %d launches' % (tot, sum(a[0] for a in agg.values())))
for k, a in sorted(agg.items(), key=lambda x: -x[1][1])[:30]:
    print('%-62s n=%4d sum %9.1f us  avg %8.1f' % (k, a[0], a[1], a[1] / a[0]))
P
