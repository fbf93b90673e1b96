#!/bin/bash
# round 2, GPU call 2b (two B200): slab parity with the FINAL kernel (all three transports) and the c5 / c4 lines on 2 GPUs
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "slabs_match" 2>&1 | tail -6 > gpurun_out/r02_tests_multi_n2_final.log
cat gpurun_out/r02_tests_multi_n2_final.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 2 --steps 200 --warmup 20 --no-cpu > gpurun_out/r02_bench_c5_n2_final.json 2> gpurun_out/r02_bench_c5_n2_final.err
cut -c1-200 gpurun_out/r02_bench_c5_n2_final.json; tail -2 gpurun_out/r02_bench_c5_n2_final.err
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu --no-dropin --no-configs1 --no-e2e > gpurun_out/r02_bench_c5_n1_final_box2.json 2> gpurun_out/r02_bench_c5_n1_final_box2.err
cut -c1-200 gpurun_out/r02_bench_c5_n1_final_box2.json
