"""Pins the oracle: the C restatement (oracle/luma_oracle.c) must reproduce, bit-for-bit, the
outputs of the UNMODIFIED reference sources compiled as oracle/_ref/luma_ref_<case>.

Two legs:
* against the committed digests in tests/golden/*.json (made by tests/golden/make_golden.py from
  the compiled reference) -- runs anywhere, needs nothing from /root/reference;
* against the compiled reference itself when oracle/_ref holds the binary (build container, or
  the GPU box, where the prebuilt binaries travel) -- full arrays, so a mismatch is localised.
"""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import port
from oracle.cases import CASES

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
# the CPU suite must stay within minutes: cap the number of steps replayed per case here
MAX_STEPS = {"cyl3d": 100, "cav3d_32": 1000, "chan3d": 1000, "cav2d_c1": 1000}


def _digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _golden(name):
    with open(os.path.join(GOLDEN, name + ".json")) as fh:
        return json.load(fh)


@pytest.mark.parametrize("name", list(CASES))
def test_port_matches_golden_digests(name):
    case = CASES[name]
    gold = _golden(name)
    g = port.PortGrid(case)
    assert float(gold["meta"]["omega"]) == g.omega
    assert float(gold["meta"]["nu"]) == g.nu
    assert float(gold["meta"]["gravity"]) == g.gravity
    assert float(gold["meta"]["rho_out"]) == g.rho_out
    assert gold["init"]["lattyp"] == _digest(g.lattyp)
    assert gold["init"]["ux_in"] == _digest(g.uin(0))
    assert gold["init"]["uy_in"] == _digest(g.uin(1))
    assert gold["init"]["xpos"] == _digest(g.pos(0))
    assert gold["init"]["ypos"] == _digest(g.pos(1))
    # wall descriptors: the reference only evaluates them on BC sites; compare there
    for tag in ["init"] + ["t%d" % s for s in case.steps]:
        if tag != "init":
            s = int(tag[1:])
            if s > MAX_STEPS.get(name, 1000):
                break
            g.step(s - g.t)
        snap = gold["snapshots"][tag]
        assert snap["f"] == _digest(g.f), (name, tag, "f")
        assert snap["rho"] == _digest(g.rho), (name, tag, "rho")
        assert snap["u"] == _digest(g.u), (name, tag, "u")
        if case.time_averaged:
            for nm in ("rho_timeav", "ui_timeav", "uiuj_timeav"):
                assert snap[nm] == _digest(getattr(g, nm)), (name, tag, nm)
        if tag != "init":
            assert float(snap["scalars"]["omega"]) == g.omega
            if case.ld_out:
                F = g.force
                assert float(snap["scalars"]["Fx"]) == F[0]
                assert float(snap["scalars"]["Fy"]) == F[1]
                assert float(snap["scalars"]["Fz"]) == F[2]
    g.close()


@pytest.mark.parametrize("name", ["cav2d_64", "chan3d_gz", "tunnel3d", "cyl2d", "sliptunnel3d", "fevel2d", "fevel2d_tav", "pleft3d_tav",
                                  "kbc2d_cyl", "kbc3d_chan"])
def test_port_matches_compiled_reference_arrays(name):
    if port.ref_binary(name) is None:
        pytest.skip("oracle/_ref/luma_ref_%s not built here" % name)
    case = CASES[name]
    steps = [s for s in case.steps if s <= 100]
    ref = port.run_ref_dump(name, steps)
    g = port.PortGrid(case)
    init = ref["init"]
    assert np.array_equal(init["lattyp"], g.lattyp)
    bc = np.isin(g.lattyp, (6, 7, 8))
    assert np.array_equal(init["wall"].reshape(-1, 5)[bc], g.wall.reshape(-1, 5)[bc])
    assert np.array_equal(init["f"], g.f)
    for s in steps:
        g.step(s - g.t)
        d = ref["t%d" % s]
        for nm in ("f", "rho", "u") + (("rho_timeav", "ui_timeav", "uiuj_timeav") if case.time_averaged else ()):
            a, b = d[nm], getattr(g, nm)
            assert np.array_equal(a, b), "%s t=%d %s: %d sites differ, first at %d" % (
                name, s, nm, int((a != b).sum()), int(np.flatnonzero(a != b)[0]))
    g.close()


def test_ramp_helpers():
    case = CASES["cav2d_64"]
    g = port.PortGrid(case)
    assert g.velocity_ramp(0.0) == 0.0
    assert g.velocity_ramp(1.0) == 1.0
    assert 0.0 < g.velocity_ramp(0.1) < 1.0
    g.close()
