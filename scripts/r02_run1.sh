#!/bin/bash
# round 2, GPU call 1 (one B200):  gpurun --timeout 1500 -- 'bash scripts/r02_run1.sh'
# parity suite, the default bench line (c5 384^3 + configs[1] 256^3 + drop-in e2e), configs[2]/[3] on one GPU (300 steps,
# clocks in the line), launch list of the default bench command
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/r02_gpu.txt
free -g >> gpurun_out/r02_gpu.txt; nproc >> gpurun_out/r02_gpu.txt
timeout 1100 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/r02_tests_gpu_n1.log
tail -5 gpurun_out/r02_tests_gpu_n1.log
timeout 400 python bench.py --steps 200 --warmup 20 > gpurun_out/r02_bench_c5_n1.json 2> gpurun_out/r02_bench_c5_n1.err
cut -c1-600 gpurun_out/r02_bench_c5_n1.json; tail -3 gpurun_out/r02_bench_c5_n1.err
for w in c2 c3 c4; do
  timeout 300 python bench.py --workload $w --steps 300 --warmup 20 --no-dropin > gpurun_out/r02_bench_${w}_n1.json 2> gpurun_out/r02_bench_${w}_n1.err
  cut -c1-300 gpurun_out/r02_bench_${w}_n1.json; tail -2 gpurun_out/r02_bench_${w}_n1.err
done
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_c5_n1_driver_cmd.json 2> gpurun_out/r02_bench_c5_n1_driver_cmd.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_c5_n1.csv python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu --no-parity --no-configs1 > gpurun_out/r02_ncu_launches.log 2>&1
tail -3 gpurun_out/r02_ncu_launches.log
