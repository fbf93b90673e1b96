# peer-store halo at 8 GPUs:  gpurun --gpus 8 --timeout 300 -- 'bash scripts/_run8.sh'
set -x
mkdir -p gpurun_out
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 100 --warmup 5 --no-e2e --no-cpu > gpurun_out/s7_c2_n8_p2p.json 2> gpurun_out/s7_c2_n8_p2p.err
cut -c1-200 gpurun_out/s7_c2_n8_p2p.json; tail -5 gpurun_out/s7_c2_n8_p2p.err
