#!/bin/bash
O=gpurun_out/r02s34; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 180 -k "nodata or multi or zero or ragged or golden" 2>&1 | tail -3 | tee $O/tests.log
NANB=0.05 PARITY=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"dct_gemm|rt_inverse" -c 4 --csv --log-file $O/dct_launches_nodata.csv python scripts/c2_probe.py > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"dct_gemm<double" -c 4 --csv --log-file $O/dct_launches_c3.csv python scripts/c3_probe.py 0 > /dev/null 2>&1
python - <<'P'
import csv
for f in ('dct_launches_nodata.csv', 'dct_launches_c3.csv'):
    rows = [r for r in csv.reader(open('gpurun_out/r02s34/' + f)) if len(r) > 10]
    hdr = rows[0]; ik = hdr.index('Kernel Name'); iv = hdr.index('Metric Value')
    for r in rows[1:5]: print(f, r[ik][:50], r[iv])
P
