#!/bin/bash
# ncu full capture of the bf16 path's three dominant kernels on a short run (8192 rays, coarse + fine)
mkdir -p gpurun_out
timeout 200 python tools/prof_small.py --rays 8192 --precision bf16 > gpurun_out/prof_small_bf16_plain.log 2>&1 &&
timeout 800 ncu --set full --clock-control none --import-source on -k regex:'k_mlp_tc|k_gather_tc|k_geom_query' -s 9 -c 6 -o gpurun_out/prof_bf16_r2 -f python tools/prof_small.py --rays 8192 --precision bf16 > gpurun_out/ncu_full_bf16.log 2>&1
echo "ncu full exit $?"; tail -3 gpurun_out/ncu_full_bf16.log
