#!/bin/bash
# multi-GPU checks: N-rank image == 1-GPU image through NCCL, then the bench at N ranks
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/dist_check.py 2>&1 | grep -E "precision|DIST_CHECK|Error|error" | head
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 3 --warmup 3 ${BENCH_ARGS:-} > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "bench exit $?"
python - $N <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads([l for l in open(f'gpurun_out/bench_{n}gpu.json') if l.startswith('{')][-1])
    print(n, 'GPUs: ms/view', round(d['ms_per_view'], 2), 'rays/s', round(d['value']), 'e2e ms', round(d['e2e']['ms_per_view'], 2), 'e2e rays/s', round(d['e2e']['value']), d['clocks'])
    for k in ('workload_C', 'workload_D', 'workload_E'):
        if k in d:
            b = d[k]; print(k, {x: b[x] for x in ('ms_per_view', 'ms_per_frame', 'ms_per_step', 'value') if x in b}, b.get('e2e', {}).get('ms_per_view'), b.get('e2e', {}).get('ms_per_step'))
except Exception as e:
    print('parse failed', e); print(open(f'gpurun_out/bench_{n}gpu.err').read()[-2500:])
PY
