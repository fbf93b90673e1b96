#!/bin/bash
# round 2, GPU call 1j (one B200): FINAL build (Smagorinsky kernel evaluates the equilibrium twice, 6 CTAs/SM): parity suite, c4 and c5 lines
set -x
mkdir -p gpurun_out
timeout 1100 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r02_tests_gpu_n1.log
tail -4 gpurun_out/r02_tests_gpu_n1.log
timeout 300 python bench.py --workload c4 --steps 300 --warmup 20 --no-dropin > gpurun_out/r02_bench_c4_n1.json 2> gpurun_out/r02_bench_c4_n1.err
cut -c1-200 gpurun_out/r02_bench_c4_n1.json
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_c5_n1_driver_cmd.json 2> gpurun_out/r02_bench_c5_n1_driver_cmd.err
cut -c1-200 gpurun_out/r02_bench_c5_n1_driver_cmd.json
