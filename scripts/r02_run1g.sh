#!/bin/bash
# round 2, GPU call 1g (one B200): warp-uniform select path for links / pass-through lanes: probe (incl. thick z-walls), parity suite
set -x
mkdir -p gpurun_out
timeout 400 python scripts/r02_probe.py walls2 > gpurun_out/r02_probe_walls_after.txt 2>&1; cat gpurun_out/r02_probe_walls_after.txt
timeout 1100 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/r02_tests_gpu_n1.log
tail -5 gpurun_out/r02_tests_gpu_n1.log
