#!/bin/bash
# round 2, GPU call 2 (two B200):  gpurun --gpus 2 --timeout 900 -- 'bash scripts/r02_run2.sh'
# slab parity on 2 ranks with all three transports (peer-store copy kernel, NCCL, peer stores fused into the face kernels),
# then copy-kernel vs fused vs NCCL at three slab sizes, and compute-sanitizer memcheck on a 2-rank peer-store run
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "slabs_match" 2>&1 | tail -15 > gpurun_out/r02_tests_multi_n2.log
cat gpurun_out/r02_tests_multi_n2.log
for halo in fused p2p nccl; do
for res in 384 128 64; do
  st=300; [ $res -lt 200 ] && st=2000
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2953$((res/64)) bench.py --gpus 2 --steps $st --warmup 20 --halo $halo --res $res --no-e2e --no-cpu --no-parity > gpurun_out/r02_halo_${halo}_${res}_n2.json 2> gpurun_out/r02_halo_${halo}_${res}_n2.err
done; done
for f in gpurun_out/r02_halo_*_n2.json; do echo $f; python -c "
import json
d=json.load(open('$f')); print(round(d['value']), d['ms_per_step'], d['gpu_launches'], d['roofline']['frac'])"; done
# memcheck of a 2-rank peer-store run (small grid: the sanitizer serialises everything)
LUMA_TEST_NO_FUSED=1 timeout 300 compute-sanitizer --tool memcheck --target-processes all --log-file gpurun_out/r02_sanitize_n2_%p.log \
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 tests/mgpu_worker.py chan3d,cyl3d > gpurun_out/r02_sanitize_n2.out 2>&1
tail -3 gpurun_out/r02_sanitize_n2.out; grep -h "ERROR SUMMARY" gpurun_out/r02_sanitize_n2_*.log | sort | uniq -c
