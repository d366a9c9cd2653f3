# developer aid: gather-kernel variants: bf16 parity suite on the first, short bench of each
VANERF_B200_LIB=$PWD/build_variants/$1.so timeout 900 python -m pytest tests/test_tc_gpu.py -x -q -m gpu 2>&1 | tail -3
for v in "$@"; do echo "=== $v"; VANERF_B200_LIB=$PWD/build_variants/$v.so timeout 200 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-fp32-path 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_view'],2), {k: round(v,2) for k,v in d['kernel_ms_per_step'].items()}, 'e2e', round(d['e2e']['ms_per_view'],2))"; done
