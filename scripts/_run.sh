set -x
timeout 1700 python -m pytest tests/test_gpu_multi.py -m gpu -x -q --durations=5 2>&1 | tail -25 > gpurun_out/s3_tests7.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 bench.py --gpus 2 --steps 300 --warmup 20 > gpurun_out/s3_bench3_n2.json 2> gpurun_out/s3_bench3_n2.err
cat gpurun_out/s3_tests7.log; cut -c1-400 gpurun_out/s3_bench3_n2.json
