#!/bin/bash
# development aid: time k_step for every tuning variant library present, two interleaved passes
for pass in 1 2; do
for lib in luma_b200/libluma_b200.so luma_b200/libluma_b200_*.so; do
  echo "== pass $pass $lib"
  LUMA_B200_LIB=$PWD/$lib python - <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
sys.path.insert(0, os.path.join(os.getcwd(), "scripts"))
from quick_perf import run
run(256, steps=300)
run(384, steps=80)
PY
done
done
