set -x
timeout 1500 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "slabs_match" 2>&1 | tail -25 > gpurun_out/s3_tests8.log
for halo in p2p nccl; do
for res in 256 128 64; do
  st=300; [ $res -lt 200 ] && st=2000
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$((res/64)) bench.py --gpus 2 --steps $st --warmup 20 --halo $halo --res $res --no-e2e --no-cpu > gpurun_out/s3_halo_${halo}_${res}.json 2> gpurun_out/s3_halo_${halo}_${res}.err
done; done
cat gpurun_out/s3_tests8.log
for f in gpurun_out/s3_halo_*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f')); print(round(d['value']), d['ms_per_step'], d['gpu_launches'])"; done
tail -5 gpurun_out/s3_halo_p2p_256.err
