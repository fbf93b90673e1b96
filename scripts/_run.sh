set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader | head -8
timeout 700 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "slabs_match" 2>&1 | tail -25 > gpurun_out/s5_tests_multi.log
cat gpurun_out/s5_tests_multi.log
for halo in p2p nccl; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 4 --steps 300 --warmup 20 --halo $halo --no-e2e --no-cpu > gpurun_out/s5_c2_n4_${halo}.json 2> gpurun_out/s5_c2_n4_${halo}.err
done
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 4 --steps 40 --warmup 5 --workload k27 --no-e2e --no-cpu > gpurun_out/s5_k27_n4.json 2> gpurun_out/s5_k27_n4.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 4 --steps 2000 --warmup 20 --res 64 --no-e2e --no-cpu > gpurun_out/s5_c2_64_n4_p2p.json 2> gpurun_out/s5_c2_64_n4_p2p.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 4 --steps 2000 --warmup 20 --res 64 --halo nccl --no-e2e --no-cpu > gpurun_out/s5_c2_64_n4_nccl.json 2> gpurun_out/s5_c2_64_n4_nccl.err
for f in gpurun_out/s5_*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f')); print(round(d['value']), d['ms_per_step'], d['gpu_launches'], d['config']['parallelism'][:60])"; done
tail -3 gpurun_out/s5_c2_n4_p2p.err
