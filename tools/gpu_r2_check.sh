#!/bin/bash
# Round-2 validation pass: parity suite, smoke, the driver's bench lines (ours, reference arm), workload E.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -3 gpurun_out/smoke.log
S=$(date +%s); timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit $? in $(( $(date +%s) - S )) s"
timeout 600 python bench.py --workload E --steps 3 --warmup 3 > gpurun_out/bench_E.json 2> gpurun_out/bench_E.err; echo "bench E exit $?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "bench reference exit $?"
head -c 600 gpurun_out/bench_reference.json; echo
