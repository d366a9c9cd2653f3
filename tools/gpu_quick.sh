#!/bin/bash
# quick correctness + timing of the tensor-core path; a failing / hanging correctness pass skips the bench
mkdir -p gpurun_out
timeout 150 python tools/tc_debug.py > gpurun_out/tc_debug.log 2>&1; rc=$?; echo "tc_debug exit $rc"; grep -E "selftest|shade|raw |latent|rgba|FAILED" gpurun_out/tc_debug.log | head -40
if [ $rc -ne 0 ] || grep -q "tc_error=[1-9]" gpurun_out/tc_debug.log; then echo "correctness pass failed: no bench"; exit 1; fi
timeout 300 python bench.py --precision bf16 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err; echo "bench exit $?"
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench_bf16.json'))
    print('ms/view', d['ms_per_view'], 'rays/s', d['value'], 'mlp frac', d['roofline']['frac'], 'TF', d['roofline']['achieved'])
    print(d['kernel_ms_per_step']); print('e2e', d['e2e']); print(d['clocks'])
except Exception as e: print('bench parse failed', e); print(open('gpurun_out/bench_bf16.err').read()[-2000:])
PY
