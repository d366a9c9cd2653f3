#!/bin/bash
mkdir -p gpurun_out
for v in "$@"; do
VANERF_B200_LIB=$PWD/build_variants/$v.so timeout 200 python tools/tc_trace.py 592 > gpurun_out/tc_$v.log 2>&1; echo "$v exit $?"; head -1 gpurun_out/tc_$v.log
done
