#!/bin/bash
# session 15: the pyramid of one C3 tile, launch by launch
O=gpurun_out/r02s15; mkdir -p $O
S2PB_TRACE=2 timeout 300 python scripts/c3_probe.py 2 > $O/c3_trace.txt 2>&1; grep -E "level|tile" $O/c3_trace.txt | tail -14
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_c3.csv python scripts/c3_probe.py 0 > $O/c3_ncu.log 2>&1
python - <<'P'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/r02s15/launches_c3.csv')) if len(r) > 10]
hdr = rows[0]; ik = hdr.index('Kernel Name'); iv = hdr.index('Metric Value'); iu = hdr.index('Metric Unit')
agg = collections.OrderedDict()
for r in rows[1:]:
    v = float(r[iv].replace(',', '')); u = r[iu]
    v = v / 1000 if u in ('ns', 'nsecond') else v * (1000 if u in ('ms', 'msecond') else 1)
    k = r[ik].replace('s2pb::', '')[:58]
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print('total %.1f us over %d launches' % (tot, sum(a[0] for a in agg.values())))
for k, a in sorted(agg.items(), key=lambda x: -x[1][1])[:40]:
    print('%-60s n=%4d sum %9.1f us  avg %8.1f' % (k, a[0], a[1], a[1] / a[0]))
P
