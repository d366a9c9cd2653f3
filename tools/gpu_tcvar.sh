# developer aid: variants of the tensor-core path: bf16 parity suite (errors printed) + short bench of each
for v in "$@"; do echo "=== $v"
  VANERF_B200_LIB=$PWD/build_variants/$v.so timeout 600 python -m pytest tests/test_tc_gpu.py -x -q -m gpu -s 2>&1 | grep -E "passed|failed|bf16 max-abs" | cut -c1-330 | tail -8
  VANERF_B200_LIB=$PWD/build_variants/$v.so timeout 200 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-fp32-path --no-reuse-variant 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_view'],2), {k: round(v,2) for k,v in d['kernel_ms_per_step'].items()}, 'e2e', round(d['e2e']['ms_per_view'],2))"
done
