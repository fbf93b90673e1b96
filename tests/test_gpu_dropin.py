"""The drop-in, end to end: the UNMODIFIED LUMA host code (GridManager, GridObj::LBM_initGrid,
ObjectManager body labelling -- compiled from /root/reference by `make -C oracle dropin`) linked with
luma_b200/host/GridObj_ops_lbm_b200.cpp in place of its CPU LBM_multi_opt, stepping on the GPU through
the C ABI.  Its dumps must carry the digests the reference's own CPU run produced (tests/golden)."""
import hashlib
import json
import os
import subprocess
import tempfile

import numpy as np
import pytest

from oracle.cases import CASES

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(os.path.dirname(HERE), "oracle", "_ref")


def _digest(path):
    return hashlib.sha256(np.fromfile(path, dtype=np.float64).tobytes()).hexdigest()


@pytest.mark.parametrize("name", ["cav2d_64", "cyl3d", "chan3d", "tunnel2d", "cav2d_reramp", "sliptunnel2d", "fevel2d_tav", "pleft3d_tav",
                                  "kbc2d_cyl", "kbc3d_chan"])
def test_luma_host_with_gpu_time_step_reproduces_reference_digests(name):
    exe = os.path.join(REF, "luma_dropin_" + name)
    if not os.path.exists(exe):
        pytest.skip("drop-in binary not built (make -C oracle dropin; needs the reference sources)")
    case = CASES[name]
    steps = [s for s in case.steps if s <= 1000]
    gold = json.load(open(os.path.join(HERE, "golden", name + ".json")))
    with tempfile.TemporaryDirectory(prefix="luma_dropin_") as out:
        r = subprocess.run([exe, "dump", out, ",".join(map(str, steps))], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600)
        assert r.returncode == 0, r.stdout.decode()[-2000:] + open(os.path.join(out, "luma_ref_log.out")).read()[-2000:]
        for tag in ["init"] + ["t%d" % s for s in steps]:
            snap = gold["snapshots"][tag]
            for nm in ("f", "rho", "u") + (("rho_timeav", "ui_timeav", "uiuj_timeav") if case.time_averaged else ()):
                assert _digest(os.path.join(out, "%s.%s.f64" % (tag, nm))) == snap[nm], (name, tag, nm)
            if tag != "init":
                sc = dict(l.strip().split("=", 1) for l in open(os.path.join(out, tag + ".scalars.txt")) if "=" in l)
                assert int(sc["t"]) == int(tag[1:])
                assert float(sc["omega"]) == float(snap["scalars"]["omega"])
