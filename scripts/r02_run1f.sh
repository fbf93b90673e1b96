#!/bin/bash
# round 2, GPU call 1f (one B200): ncu of the same kernel on the fully periodic box vs walls in z only (sector-completing stores on)
set -x
mkdir -p gpurun_out
for c in walls_ walls_z; do
  timeout 300 ncu --set full --clock-control none -k 'regex:^k_step$' -s 14 -c 1 -f -o /tmp/r02_ncu_$c python scripts/r02_probe.py one $c 256 > gpurun_out/r02_ncu_$c.log 2>&1
  ncu -i /tmp/r02_ncu_$c.ncu-rep --page raw --csv > gpurun_out/r02_ncu_${c}256_raw.csv 2>/dev/null
done
ls -la gpurun_out/*walls*
