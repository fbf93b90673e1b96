set -x
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/s3_tests6.log
timeout 300 python bench.py > gpurun_out/s3_bench3_n1.json 2> gpurun_out/s3_bench3_n1.err
python - > gpurun_out/s3_small.log 2>&1 <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "scripts"))
from quick_perf import run
run(256, dims=2, steps=4000)
run(512, dims=2, steps=2000)
run(64, dims=3, steps=2000)
run(128, dims=3, steps=500)
run(256, steps=200, smag=True)
os.environ["LUMA_B200_GRAPH_STEPS"] = "0"
run(256, dims=2, steps=4000)
run(512, dims=2, steps=2000)
run(64, dims=3, steps=2000)
run(128, dims=3, steps=500)
PY
cat gpurun_out/s3_tests6.log gpurun_out/s3_bench3_n1.json gpurun_out/s3_small.log; tail -3 gpurun_out/s3_bench3_n1.err
