set -x
nvidia-smi -L | wc -l
timeout 1200 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/s3_mgpu_tests.log
for n in 2 4 8; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2950$n bench.py --gpus $n --steps 300 --warmup 20 > gpurun_out/s3_bench_n$n.json 2> gpurun_out/s3_bench_n$n.err
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 100 --warmup 10 --workload c5 > gpurun_out/s3_bench_c5_n8.json 2> gpurun_out/s3_bench_c5_n8.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 8 --steps 20 --warmup 3 --impl reference > gpurun_out/s3_bench_ref_n8.json 2> gpurun_out/s3_bench_ref_n8.err
cat gpurun_out/s3_mgpu_tests.log; cat gpurun_out/s3_bench_n*.json gpurun_out/s3_bench_c5_n8.json gpurun_out/s3_bench_ref_n8.json | cut -c1-400
tail -3 gpurun_out/s3_bench_n8.err
