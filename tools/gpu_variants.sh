#!/bin/bash
# Developer aid: correctness pass (tools/tc_debug.py) + short bench of several builds of the library (variants of compile-time
# knobs, built by tools/build_variants.py into build_variants/*.so), selected through VANERF_B200_LIB.
# usage: tools/gpu_variants.sh "name[:ENV=VAL,...]" ...
mkdir -p gpurun_out
for spec in "$@"; do
  name=${spec%%:*}; envs=""
  if [[ "$spec" == *:* ]]; then envs=$(echo "${spec#*:}" | tr ',' ' '); fi
  lib=$PWD/build_variants/$name.so
  tag=$(echo "$spec" | tr ':,=' '___')
  echo "=== $spec"
  if [ -z "$SKIP_DEBUG" ] || [[ "$name" != a* ]]; then
  env VANERF_B200_LIB=$lib $envs timeout 150 python tools/tc_debug.py > gpurun_out/var_${tag}_debug.log 2>&1; rc=$?
  grep -E "tc_error|FAILED|max" gpurun_out/var_${tag}_debug.log | tail -4
  if [ $rc -ne 0 ] || grep -q "tc_error=[1-9]" gpurun_out/var_${tag}_debug.log; then echo "correctness pass failed ($rc)"; continue; fi
  fi
  env VANERF_B200_LIB=$lib $envs timeout 300 python bench.py --precision bf16 --steps 2 --warmup 3 --no-cpu-baseline --no-fp32-path --no-reuse-variant --no-secondary > gpurun_out/var_${tag}.json 2> gpurun_out/var_${tag}.err
  python - "$tag" <<'PY'
import json, sys
try:
    d = json.load(open(f'gpurun_out/var_{sys.argv[1]}.json'))
    print('ms/view', round(d['ms_per_view'], 2), {k: round(v, 2) for k, v in d['kernel_ms_per_step'].items()}, 'e2e ms', round(d['e2e']['ms_per_view'], 2), d['clocks'])
except Exception as e:
    print('bench parse failed', e); print(open(f'gpurun_out/var_{sys.argv[1]}.err').read()[-1500:])
PY
done
