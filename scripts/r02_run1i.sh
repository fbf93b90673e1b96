#!/bin/bash
# round 2, GPU call 1i (one B200): Smagorinsky "evaluate the equilibrium twice" variants (parity + bandwidth), and BASELINE configs[0]
# (256^2 D2Q9 cavity, launch-bound) through the drop-in binary vs the batched Python call
set -x
mkdir -p gpurun_out
: > gpurun_out/r02_probe_smag.txt
for lib in luma_b200/libluma_b200.so luma_b200/libluma_b200_smagrc6.so luma_b200/libluma_b200_smagrc5.so; do
  [ -f $lib ] || continue
  echo "== $lib" >> gpurun_out/r02_probe_smag.txt
  LUMA_B200_LIB=$PWD/$lib timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "(upload_path_bitwise or device_init_path) and (cyl or duct3d or thin3d)" 2>&1 | tail -2 >> gpurun_out/r02_probe_smag.txt
  LUMA_B200_LIB=$PWD/$lib timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "parity_at_size and c4_r64" 2>&1 | tail -2 >> gpurun_out/r02_probe_smag.txt
  LUMA_B200_LIB=$PWD/$lib timeout 200 python scripts/r02_probe.py smag >> gpurun_out/r02_probe_smag.txt 2>&1
done
cat gpurun_out/r02_probe_smag.txt
oracle/_ref/luma_dropin_cav2d_c1 bench 40 8000 > gpurun_out/r02_c1_dropin.json 2> gpurun_out/r02_c1_dropin.err; cat gpurun_out/r02_c1_dropin.json
timeout 120 python scripts/r02_probe.py c1 > gpurun_out/r02_c1_python.txt 2>&1; cat gpurun_out/r02_c1_python.txt
