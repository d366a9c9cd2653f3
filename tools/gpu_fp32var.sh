# developer aid: variants of the fp32 path: fp32 parity tests on the first + one timed fp32 view each
VANERF_B200_LIB=$PWD/build_variants/$1.so timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2
for v in "$@"; do echo "=== $v"; VANERF_B200_LIB=$PWD/build_variants/$v.so timeout 300 python bench.py --precision fp32 --steps 1 --warmup 3 --no-cpu-baseline --no-reuse-variant 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_view'],1), {k: round(v,1) for k,v in d['kernel_ms_per_step'].items()})"; done
