#!/bin/bash
# developer aid: per-case error tables of the bf16 parity suite for several builds (no -x: every case reports)
for v in "$@"; do echo "=== $v"
  VANERF_B200_LIB=$PWD/build_variants/$v.so timeout 600 python -m pytest tests/test_tc_gpu.py -q -m gpu -s -k "shading_matches or full_view" 2>&1 | grep -E "passed|failed|bf16 max-abs|FAILED" | cut -c1-420
done
