#!/bin/bash
# full GPU test-suite + bf16 / fp32 bench lines
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --precision bf16 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err; echo "bench exit $?"
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench_bf16.json'))
    print('ms/view', d['ms_per_view'], 'rays/s', d['value'], 'mlp frac', d['roofline']['frac'], 'TF', d['roofline']['achieved'])
    print(d['kernel_ms_per_step']); print('e2e', d['e2e']); print(d['clocks'])
except Exception as e: print('bench parse failed', e); print(open('gpurun_out/bench_bf16.err').read()[-2000:])
PY
