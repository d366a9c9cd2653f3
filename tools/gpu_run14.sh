#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --precision fp32 --steps 3 --warmup 2 --no-cpu-baseline --no-reuse-variant --no-secondary > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err; echo "bench fp32 exit $?"
python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/bench_fp32.json'))
    print('fp32 split: ms/view', round(d['ms_per_view'], 2), {k: round(v, 2) for k, v in d['kernel_ms_per_step'].items()}, 'e2e ms', round(d['e2e']['ms_per_view'], 2), 'frac', d['roofline']['frac'])
except Exception as e:
    print('parse failed', e); print(open('gpurun_out/bench_fp32.err').read()[-2000:])
PY
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -2
