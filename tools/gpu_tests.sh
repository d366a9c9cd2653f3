#!/bin/bash
# full GPU test-suite (all failures listed) + the default bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; grep -E "passed|failed|FAILED|Error" gpurun_out/pytest_gpu.log | tail -15; grep -E "max-abs errors" gpurun_out/pytest_gpu.log | tail -20
timeout 900 python bench.py ${BENCH_ARGS:-} > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit $?"
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench_default.json'))
    print('ms/view', d['ms_per_view'], 'rays/s', d['value'], 'mlp frac', d['roofline']['frac'], 'gather frac', d['roofline_gather']['frac'], 'geom', d['geometry']['ms_per_view'])
    print(d['kernel_ms_per_step']); print('e2e', d['e2e']); print(d['clocks'], 'launches', d['gpu_launches'])
    for k in ('fp32_path','workload_C','workload_D','coarse_reuse','cpu_baseline'):
        if k in d:
            b=d[k]; print(k, {x: b[x] for x in ('ms_per_view','ms_per_frame','value','frames_per_s') if x in b}, b.get('roofline',{}).get('frac'), b.get('e2e',{}).get('ms_per_view'), b.get('sample',''))
except Exception as e: print('bench parse failed', e); print(open('gpurun_out/bench_default.err').read()[-3000:])
PY
