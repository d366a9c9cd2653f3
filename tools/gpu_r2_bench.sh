#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_encoders.py -m gpu -x -q 2>&1 | tail -5
S=$(date +%s); timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit $? in $(( $(date +%s) - S )) s"
tail -3 gpurun_out/bench_default.err
