#!/bin/bash
# round 2, GPU call 1c (one B200): parity of the two kernel variants, variant/tuning matrix, ncu --set full captures
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "stats_window or without_populations or parity_at_size" 2>&1 | tail -5 > gpurun_out/r02_tests_fixed.log; cat gpurun_out/r02_tests_fixed.log
for v in V2 TMA FILL; do
  env LUMA_B200_$v=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "upload_path_bitwise or device_init_path" 2>&1 | tail -4 > gpurun_out/r02_tests_variant_$v.log; cat gpurun_out/r02_tests_variant_$v.log
done
timeout 900 python scripts/r02_probe.py v2 > gpurun_out/r02_probe_v2.txt 2>&1; cat gpurun_out/r02_probe_v2.txt
: > gpurun_out/r02_probe_libs.txt
for lib in luma_b200/libluma_b200.so luma_b200/libluma_b200_*.so; do
  LUMA_B200_LIB=$PWD/$lib timeout 200 python scripts/r02_probe.py lib >> gpurun_out/r02_probe_libs.txt 2>&1
done
cat gpurun_out/r02_probe_libs.txt
ncu1() {  # ncu1 <tag> <kernel regex> <case> <res> [ENV=VALUE]
  local tag=$1 k=$2; shift 2
  timeout 300 ncu --set full --clock-control none --import-source on -k "regex:$k" -s 14 -c 1 -f -o gpurun_out/r02_ncu_$tag python scripts/r02_probe.py one "$@" > gpurun_out/r02_ncu_$tag.log 2>&1
  ncu -i gpurun_out/r02_ncu_$tag.ncu-rep --page raw --csv > gpurun_out/r02_ncu_${tag}_raw.csv 2>/dev/null
  tail -2 gpurun_out/r02_ncu_$tag.log
}
ncu1 box256 "^k_step<" box 256
ncu1 channel256 "^k_step<" channel 256
ncu1 cavity384 "^k_step<" cavity 384
ncu1 channelf512 "^k_step<" channel_f 512
ncu1 channels256 "^k_step<" channel_s 256
ncu1 box256_fill "^k_step<" box 256 LUMA_B200_FILL=1
ncu1 v2_channel256 "^k_step_v2<" channel 256 LUMA_B200_V2=1
ncu1 tma_channel256 "^k_step_tma<" channel 256 LUMA_B200_TMA=1
ls -la gpurun_out/*.ncu-rep
