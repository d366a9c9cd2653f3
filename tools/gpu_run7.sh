#!/bin/bash
mkdir -p gpurun_out
for v in cur cnt cur cnt; do
  env VANERF_B200_LIB=$PWD/build_variants/$v.so timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-fp32-path --no-reuse-variant --no-secondary > gpurun_out/var_$v.json 2> gpurun_out/var_$v.err
  python - $v <<'PY'
import json, sys
d = json.load(open(f'gpurun_out/var_{sys.argv[1]}.json'))
print(sys.argv[1], 'ms/view', round(d['ms_per_view'], 2), {k: round(v, 2) for k, v in d['kernel_ms_per_step'].items()}, 'e2e ms', round(d['e2e']['ms_per_view'], 2), d['clocks'])
PY
done
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv
python tools/time_setup.py > gpurun_out/time_setup.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/setup_launches.csv python tools/time_setup.py > gpurun_out/ncu_setup.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open('gpurun_out/setup_launches.csv', errors='ignore')))
hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
kn, mv = rows[hdr].index('Kernel Name'), rows[hdr].index('Metric Value')
agg = collections.OrderedDict()
for r in rows[hdr + 2:]:
    if len(r) > mv:
        agg.setdefault(r[kn][:60], []).append(float(r[mv].replace(',', '')))
for k, v in agg.items():
    print(f'{k:60s} n={len(v):3d} last {v[-1] / 1e3:9.1f} us')
PY
