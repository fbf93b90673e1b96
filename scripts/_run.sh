set -x
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/s3_tests2.log
timeout 300 python scripts/quick_perf.py > gpurun_out/s3_quick2.log 2>&1
cat gpurun_out/s3_tests2.log gpurun_out/s3_quick2.log
