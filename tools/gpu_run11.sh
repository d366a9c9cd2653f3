#!/bin/bash
mkdir -p gpurun_out
python tools/time_setup.py > gpurun_out/time_setup.log 2>&1; tail -2 gpurun_out/time_setup.log
ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 200 --csv --log-file gpurun_out/setup_launches.csv python tools/time_setup.py > gpurun_out/ncu_setup.log 2>&1
python - <<'PY'
import csv
rows = list(csv.reader(open('gpurun_out/setup_launches.csv', errors='ignore')))
hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
kn, mv = rows[hdr].index('Kernel Name'), rows[hdr].index('Metric Value')
out=[(r[kn][:48], float(r[mv].replace(',',''))/1e3) for r in rows[hdr+2:] if len(r)>mv]
for k,v in out[-32:]: print(f'{k:50s} {v:9.1f} us')
PY
python -m pytest tests -m gpu -q -x 2>&1 | tail -2
