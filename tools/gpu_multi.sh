# 2-GPU validation of both sharding modes (torchrun, NCCL): workload B (pixels interleaved over ranks) and D (frames round robin)
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "B exit $?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload D --frames 8 --steps 2 --warmup 1 > gpurun_out/bench_D_${N}gpu.json 2> gpurun_out/bench_D_${N}gpu.err; echo "D exit $?"
python - $N <<'PY'
import json, sys
n = sys.argv[1]
for f in (f'gpurun_out/bench_{n}gpu.json', f'gpurun_out/bench_D_{n}gpu.json'):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); print(f, 'n_gpus', d['n_gpus'], 'value', round(d['value']), 'ms_per_step', round(d['ms_per_step'], 2), 'e2e', d['e2e'].get('value') and round(d['e2e']['value']))
    except Exception as e:
        print(f, 'parse failed', e); print(open(f.replace('.json', '.err')).read()[-1500:])
PY
