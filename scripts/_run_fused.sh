# round 2, first multi-GPU call:  gpurun --gpus 2 --timeout 600 -- 'bash scripts/_run_fused.sh'
# parity of the experimental fused halo exchange (LUMA_B200_FUSED_HALO) on 2 GPUs, then copy-kernel vs fused vs NCCL at three slab sizes
set -x
mkdir -p gpurun_out
LUMA_TEST_FUSED=1 timeout 500 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "slabs_match" 2>&1 | tail -15 > gpurun_out/fused_tests.log
cat gpurun_out/fused_tests.log
for halo in fused p2p nccl; do
for res in 256 128 64; do
  st=300; [ $res -lt 200 ] && st=2000
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2953$((res/64)) bench.py --gpus 2 --steps $st --warmup 20 --halo $halo --res $res --no-e2e --no-cpu > gpurun_out/fused_${halo}_${res}.json 2> gpurun_out/fused_${halo}_${res}.err
done; done
for f in gpurun_out/fused_*.json; do echo $f; python -c "
import json
d=json.load(open('$f')); print(round(d['value']), d['ms_per_step'], d['gpu_launches'])"; done
