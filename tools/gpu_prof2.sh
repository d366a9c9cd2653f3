#!/bin/bash
# ncu full capture of the geometry and gather kernels on a short bf16 run
mkdir -p gpurun_out
timeout 200 python tools/prof_small.py --rays 8192 --precision bf16 > gpurun_out/prof_small_bf16_plain.log 2>&1 &&
timeout 800 ncu --set full --clock-control none --import-source on -k regex:'k_geom_query|k_gather_tc' -s 2 -c 3 -o gpurun_out/prof_geom_r2 -f python tools/prof_small.py --rays 8192 --precision bf16 > gpurun_out/ncu_geom.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_geom.log
