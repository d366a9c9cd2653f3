#!/bin/bash
# round 2 experiment: trace of the current build + ablation variants
mkdir -p gpurun_out
VANERF_B200_LIB=$PWD/vanerf_b200/libvanerf_b200_trace.so timeout 200 python tools/tc_trace.py 592 > gpurun_out/tc_trace_r2.log 2>&1; echo "trace exit $?"; head -1 gpurun_out/tc_trace_r2.log; tail -8 gpurun_out/tc_trace_r2.log
SKIP_DEBUG=1 bash tools/gpu_variants.sh base pe4 a32 a4
