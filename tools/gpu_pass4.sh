#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/prof_small.py --rays 8192 --precision bf16 > gpurun_out/prof_small_bf16_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_mlp_tc|k_gather_tc' -s 4 -c 2 -o gpurun_out/prof_bf16 python tools/prof_small.py --rays 8192 --precision bf16 > gpurun_out/ncu_full_bf16.log 2>&1
echo "ncu full exit $?"; tail -3 gpurun_out/ncu_full_bf16.log
