#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; grep -E "passed|failed|FAILED|Error" gpurun_out/pytest_gpu.log | tail -8
python tools/time_setup.py > gpurun_out/time_setup.log 2>&1; tail -2 gpurun_out/time_setup.log
