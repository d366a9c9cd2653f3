#!/bin/bash
# GPU pass 2: full parity suite, smoke, ncu launch list + full captures of the fp32 path.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
timeout 300 python tools/prof_small.py --rays 8192 > gpurun_out/prof_small_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_fp32.csv python tools/prof_small.py --rays 8192 > gpurun_out/ncu_list.log 2>&1
echo "ncu list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_gather|k_mlp_simt|k_geom_query' -s 6 -c 3 -o gpurun_out/prof_fp32 python tools/prof_small.py --rays 8192 > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"
ls -la gpurun_out
