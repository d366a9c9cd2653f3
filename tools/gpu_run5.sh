#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_train.py tests/test_stages.py -m gpu -q -s 2>&1 | tail -5
timeout 600 python bench.py --workload E --steps 3 --warmup 2 > gpurun_out/bench_E.json 2> gpurun_out/bench_E.err; echo "bench E exit $?"; cat gpurun_out/bench_E.json | cut -c1-1500; tail -3 gpurun_out/bench_E.err
for v in cur; do
  env VANERF_B200_LIB=$PWD/build_variants/$v.so timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-fp32-path --no-reuse-variant --no-secondary > gpurun_out/var_$v.json 2> gpurun_out/var_$v.err
  python - $v <<'PY'
import json, sys
d = json.load(open(f'gpurun_out/var_{sys.argv[1]}.json'))
print(sys.argv[1], 'ms/view', round(d['ms_per_view'], 2), {k: round(v, 2) for k, v in d['kernel_ms_per_step'].items()}, 'e2e ms', round(d['e2e']['ms_per_view'], 2))
PY
done
