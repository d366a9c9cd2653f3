#!/bin/bash
# Round-end GPU pass: parity suite, smoke, bench lines, ncu launch list of the bench command, ncu full capture (bf16 path).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err; echo "bench exit $?"
timeout 600 python bench.py --workload D --frames 8 --steps 2 --warmup 1 > gpurun_out/bench_D.json 2> gpurun_out/bench_D.err; echo "bench D exit $?"
timeout 200 python tools/time_setup.py > gpurun_out/time_setup.log 2>&1; tail -2 gpurun_out/time_setup.log
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "bench reference exit $?"
python - <<'PY'
import json
for f in ('gpurun_out/bench_bf16.json','gpurun_out/bench_reference.json'):
    try:
        d=json.load(open(f)); print(f, 'value', d['value'], d.get('ms_per_step'), d.get('kernel_ms_per_step'), d.get('e2e'), d.get('fp32_path'), d.get('cpu_baseline'))
    except Exception as e: print(f, 'parse failed', e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_bf16.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-fp32-path > gpurun_out/ncu_list_bf16.log 2>&1; echo "ncu list exit $?"
timeout 200 python tools/prof_small.py --rays 8192 --precision bf16 > gpurun_out/prof_small_bf16_plain.log 2>&1 &&
timeout 800 ncu --set full --clock-control none --import-source on -k regex:'k_mlp_tc|k_gather_tc|k_geom_query' -s 3 -c 6 -o gpurun_out/prof_bf16_r7 -f python tools/prof_small.py --rays 8192 --precision bf16 > gpurun_out/ncu_full_bf16.log 2>&1
echo "ncu full exit $?"
