#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; grep -E "passed|failed|FAILED|Error" gpurun_out/pytest_gpu.log | tail -8
python tools/time_setup.py > gpurun_out/time_setup.log 2>&1; tail -3 gpurun_out/time_setup.log
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit $?"
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench_default.json'))
    print('ms/view', d['ms_per_view'], 'rays/s', d['value'], 'mlp frac', d['roofline']['frac'], 'gather frac', d['roofline_gather']['frac'], 'geom', d['geometry']['ms_per_view'])
    print(d['kernel_ms_per_step']); print('e2e', d['e2e']); print(d['clocks'], 'launches', d['gpu_launches'])
    for k in ('fp32_path','workload_C','workload_D','workload_E','coarse_reuse','cpu_baseline'):
        if k in d:
            b=d[k]; print(k, {x: b[x] for x in ('ms_per_view','ms_per_frame','ms_per_step','value','frames_per_s') if x in b}, b.get('roofline',{}).get('frac'), b.get('e2e',{}).get('ms_per_view'), b.get('sample',''))
except Exception as e: print('bench parse failed', e); print(open('gpurun_out/bench_default.err').read()[-3000:])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/setup_launches.csv python tools/time_setup.py > gpurun_out/ncu_setup.log 2>&1
python - <<'PY'
import csv
rows = list(csv.reader(open('gpurun_out/setup_launches.csv', errors='ignore')))
hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
kn, mv = rows[hdr].index('Kernel Name'), rows[hdr].index('Metric Value')
out=[(r[kn][:48], float(r[mv].replace(',',''))/1e3) for r in rows[hdr+2:] if len(r)>mv]
last=[i for i,(k,_) in enumerate(out) if k.startswith('void k_gf_conv3x3<3')][-1]
for k,v in out[last:last+24]: print(f'{k:50s} {v:9.1f} us')
PY
