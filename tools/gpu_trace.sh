#!/bin/bash
mkdir -p gpurun_out
VANERF_B200_LIB=$PWD/vanerf_b200/libvanerf_b200_trace.so timeout 200 python tools/tc_trace.py 592 > gpurun_out/tc_trace.log 2>&1; echo "trace exit $?"; head -1 gpurun_out/tc_trace.log; tail -12 gpurun_out/tc_trace.log
bash tools/gpu_quick.sh
