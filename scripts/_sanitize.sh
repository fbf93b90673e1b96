#!/bin/bash
# compute-sanitizer on one B200:  gpurun --timeout 900 -- 'bash scripts/_sanitize.sh'
# memcheck + initcheck on single-GPU cases that cover every kernel family (list kernel with regularised BCs, per-link path,
# time averages, KBC on D3Q27, two-plane periodic slab), racecheck on the shared-memory kernels (layout conversion, momentum exchange)
set -x
mkdir -p gpurun_out
cat > /tmp/san_case.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
import luma_b200
from oracle import port
from oracle.cases import CASES
from util import defs_from_case
for name in sys.argv[1].split(","):
    case = CASES[name]
    ref = port.PortGrid(case)
    for path in ("upload", "upload_nof", "init"):
        if path == "upload_nof" and not case.no_flow:
            continue
        g = luma_b200.GridObj(defs_from_case(case))
        if path == "init":
            g.LBM_initGrid()
        else:
            g.upload(ref.f if path == "upload" else None, ref.rho, ref.u, ref.lattyp, ref.uin(0), ref.uin(1), ref.uin(2))
        for _ in range(3):
            g.LBM_multi_opt()
        g.LBM_multi_opt(4)
        got = g.download()
        if case.ld_out:
            g.computeLiftDrag()
        if case.time_averaged:
            g.download_timeav()
        g.close()
    ref.step(7)
    assert np.array_equal(got["f"], ref.f), name
    ref.close()
    print("sanitize case ok:", name, flush=True)
PY
CASES=cyl3d,fevel2d_tav,kbc3d_chan,thin3d,cav2d_64,slipchan3d
for tool in memcheck initcheck; do
  timeout 800 compute-sanitizer --tool $tool --log-file gpurun_out/r02_sanitize_${tool}.log python /tmp/san_case.py $CASES > gpurun_out/r02_sanitize_${tool}.out 2>&1
  tail -2 gpurun_out/r02_sanitize_${tool}.out; grep "ERROR SUMMARY" gpurun_out/r02_sanitize_${tool}.log
done
timeout 600 compute-sanitizer --tool racecheck --log-file gpurun_out/r02_sanitize_racecheck.log python /tmp/san_case.py cyl3d,cav2d_64 > gpurun_out/r02_sanitize_racecheck.out 2>&1
tail -2 gpurun_out/r02_sanitize_racecheck.out; grep "RACECHECK SUMMARY\|ERROR SUMMARY" gpurun_out/r02_sanitize_racecheck.log
