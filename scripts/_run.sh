set -x
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/s3_tests4.log
timeout 300 python bench.py > gpurun_out/s3_bench2_n1.json 2> gpurun_out/s3_bench2_n1.err
cat gpurun_out/s3_tests4.log gpurun_out/s3_bench2_n1.json; tail -3 gpurun_out/s3_bench2_n1.err
