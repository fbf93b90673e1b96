#!/bin/bash
# round 2, GPU call 1b (one B200): full parity suite again, the (case x size x knob) probe incl. the TMA-staged variant, compute-sanitizer
set -x
mkdir -p gpurun_out
timeout 1100 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r02_tests_gpu_n1.log
tail -5 gpurun_out/r02_tests_gpu_n1.log
timeout 600 python scripts/r02_probe.py > gpurun_out/r02_probe.txt 2>&1
cat gpurun_out/r02_probe.txt
bash scripts/_sanitize.sh
