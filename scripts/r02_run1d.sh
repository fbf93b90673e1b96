#!/bin/bash
# round 2, GPU call 1d (one B200): wall-orientation probe, ncu --set full captures of k_step on the cases that differ, and of the two variants
set -x
mkdir -p gpurun_out
timeout 300 python scripts/r02_probe.py walls > gpurun_out/r02_probe_walls.txt 2>&1; cat gpurun_out/r02_probe_walls.txt
ncu1() {  # ncu1 <tag> <kernel regex> <case> <res> [ENV=VALUE]
  local tag=$1 k=$2; shift 2
  timeout 300 ncu --set full --clock-control none --import-source on -k "regex:^$k\$" -s 14 -c 1 -f -o /tmp/r02_ncu_$tag python scripts/r02_probe.py one "$@" > gpurun_out/r02_ncu_$tag.log 2>&1
  ncu -i /tmp/r02_ncu_$tag.ncu-rep --page raw --csv > gpurun_out/r02_ncu_${tag}_raw.csv 2>/dev/null
  tail -2 gpurun_out/r02_ncu_$tag.log
}
ncu1 box256 k_step box 256
ncu1 channel256 k_step channel 256
ncu1 cavity384 k_step cavity 384
ncu1 channelf512 k_step channel_f 512
ncu1 channels256 k_step channel_s 256
ncu1 v2_channel256 k_step_v2 channel 256 LUMA_B200_V2=1
ncu1 tma_channel256 k_step_tma channel 256 LUMA_B200_TMA=1
cp /tmp/r02_ncu_cavity384.ncu-rep gpurun_out/; ls -la /tmp/r02_*.ncu-rep; du -sh gpurun_out
