#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tc_gpu.py -m gpu -x -q -s > gpurun_out/pytest_tc.log 2>&1; echo "pytest tc exit $?" >> gpurun_out/pytest_tc.log
grep -E "max-abs|passed|failed|Error|error" gpurun_out/pytest_tc.log | tail -30
timeout 600 python bench.py --precision bf16 --steps 3 --warmup 3 > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err; echo "bench exit $?"
cat gpurun_out/bench_bf16.json; tail -5 gpurun_out/bench_bf16.err
