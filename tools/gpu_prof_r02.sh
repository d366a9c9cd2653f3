#!/bin/bash
# Round-2 profile pass: plain bench (must exit 0), ncu launch list of the same command, ncu --set full of the hot kernels
# (bf16 path and split-precision fp32 path) on a small ray batch.
mkdir -p gpurun_out
timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-fp32-path --no-reuse-variant --no-secondary > gpurun_out/prof_bench_plain.json 2> gpurun_out/prof_bench_plain.err || { echo "plain bench failed"; exit 1; }
cat gpurun_out/prof_bench_plain.json | cut -c1-400
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches_bf16.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-fp32-path --no-reuse-variant --no-secondary > gpurun_out/ncu_list_bf16.log 2>&1; echo "ncu list exit $?"
timeout 200 python tools/prof_small.py --rays 16384 --precision bf16 > gpurun_out/prof_small_bf16_plain.log 2>&1 || { echo "prof_small bf16 failed"; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_mlp_tc|k_gather_tc|k_geom_query' -s 3 -c 6 -o gpurun_out/r02_prof_bf16 -f python tools/prof_small.py --rays 16384 --precision bf16 > gpurun_out/ncu_full_bf16.log 2>&1
echo "ncu full bf16 exit $?"
timeout 200 python tools/prof_small.py --rays 16384 --precision fp32 > gpurun_out/prof_small_fp32_plain.log 2>&1 || { echo "prof_small fp32 failed"; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_mlp_tc|k_rec_split|k_gather' -s 4 -c 4 -o gpurun_out/r02_prof_fp32 -f python tools/prof_small.py --rays 16384 --precision fp32 > gpurun_out/ncu_full_fp32.log 2>&1
echo "ncu full fp32 exit $?"
ls -la gpurun_out/*.ncu-rep
