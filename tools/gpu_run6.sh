#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train.py tests/test_gfeat_gpu.py -m gpu -q -s > gpurun_out/pytest6.log 2>&1; echo "pytest exit $?"; grep -E "gfeat|passed|failed|Error|rel [0-9.]+e-0[0-3]" gpurun_out/pytest6.log | tail -30
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "full pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
python tools/time_setup.py 2>&1 | tail -4
env VANERF_B200_LIB=$PWD/build_variants/cur.so timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-fp32-path --no-reuse-variant --no-secondary > gpurun_out/var_cur.json 2> gpurun_out/var_cur.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/var_cur.json'))
print('cur ms/view', round(d['ms_per_view'], 2), {k: round(v, 2) for k, v in d['kernel_ms_per_step'].items()}, 'e2e ms', round(d['e2e']['ms_per_view'], 2))
PY
