#!/bin/bash
# First GPU pass: parity tests, smoke, short bench.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
tail -5 gpurun_out/smoke.log
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err; echo "bench exit $?"
cat gpurun_out/bench_fp32.json; tail -5 gpurun_out/bench_fp32.err
