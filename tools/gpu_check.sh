#!/bin/bash
# full GPU parity suite + smoke + workload B/D bench lines (no profiler)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err; echo "bench exit $?"
timeout 600 python bench.py --workload D --frames 8 --steps 2 --warmup 1 > gpurun_out/bench_D.json 2> gpurun_out/bench_D.err; echo "bench D exit $?"; cat gpurun_out/bench_D.json | cut -c1-400; tail -3 gpurun_out/bench_D.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_bf16.json')); print('B ms/view', d['ms_per_view'], d['kernel_ms_per_step'], 'e2e', d['e2e']['ms_per_view'], 'mlp frac', d['roofline']['frac'], 'gather frac', d['roofline_gather']['frac'], d.get('fp32_path'))
PY
