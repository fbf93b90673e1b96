#!/bin/bash
# round 2, GPU call 2c (two B200): device-side exchange numbers + CUDA-graph batches on slabs: slab parity (all transports; the small
# cases run through the batches), then latency-bound slabs 64^3 / 128^3 per GPU with and without the batches, and the 1-GPU rate at 64^3
set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "slabs_match" 2>&1 | tail -6 > gpurun_out/r02_tests_multi_n2_graphs.log
cat gpurun_out/r02_tests_multi_n2_graphs.log
b2() { # b2 <res> <tag> [env]
  env $3 timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2953$(($1/64)) bench.py --gpus 2 --workload c2 --steps 4000 --warmup 40 --res $1 --no-e2e --no-cpu --no-parity > gpurun_out/r02_slab_$2.json 2> gpurun_out/r02_slab_$2.err
}
b2 64 64_graphs X=1 & CUDA_VISIBLE_DEVICES=0 true; wait
b2 64 64_live LUMA_B200_GRAPH_SLABS=0
b2 128 128_graphs X=1
CUDA_VISIBLE_DEVICES=0 timeout 100 python bench.py --workload c2 --steps 4000 --warmup 40 --res 64 --no-e2e --no-cpu --no-parity --no-dropin > gpurun_out/r02_slab_64_n1.json 2> gpurun_out/r02_slab_64_n1.err
for f in gpurun_out/r02_slab_*.json; do echo $f; python -c "
import json
d=json.load(open('$f')); print(d['n_gpus'], round(d['value']), round(1e3*d['ms_per_step'],2), 'us/step', d['gpu_launches'])"; done
