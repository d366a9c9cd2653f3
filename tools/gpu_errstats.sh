#!/bin/bash
for v in "$@"; do echo "=== $v"; for lay in bvv narrow; do VANERF_B200_LIB=$PWD/build_variants/$v.so timeout 300 python tools/tc_errstats.py $lay 12 2>&1 | grep -E "coarse|fine|per pixel|Error"; done; done
