# developer aid: memcheck of one bf16 parity case with a variant library
mkdir -p gpurun_out
VANERF_B200_LIB=$PWD/build_variants/$1.so timeout 800 /usr/local/cuda/bin/compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_tc_gpu.py -x -q -m gpu -k "512-334-3-ref" > gpurun_out/sanitizer.log 2>&1
grep -B2 -A12 "Invalid\|Error" gpurun_out/sanitizer.log | head -60 | cut -c1-220
tail -5 gpurun_out/sanitizer.log
