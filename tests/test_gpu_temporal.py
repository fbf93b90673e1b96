"""Two-steps-per-sweep path (k_step2, temporal blocking through L2): must stay bit-identical to the
oracle for every strip / lag / ring geometry, including multi-strip sweeps, the periodic closure in
x, lid boundary tiles, per-step omega (Reynolds ramp) and Smagorinsky/forcing variants."""
import numpy as np
import pytest

import luma_b200
from oracle import port
from oracle.cases import CASES
from util import defs_from_case, first_diff

pytestmark = pytest.mark.gpu

ELIGIBLE = ["chan3d", "chan3d_gz", "cav3d_32", "cav3d_64", "cav2d_64", "cav2d_reramp", "chan2d"]
GEOMS = [(0, 0, 0), (8, 1, 4), (5, 2, 4), (16, 4, 7), (3, 3, 6)]     # rows_per_strip, lag, ring_slots (0 = default)


@pytest.mark.parametrize("name", ELIGIBLE)
@pytest.mark.parametrize("geom", GEOMS)
def test_fused_sweeps_bitwise_vs_oracle(name, geom):
    case = CASES[name]
    ref = port.PortGrid(case)
    g = luma_b200.GridObj(defs_from_case(case)).LBM_initGrid()
    g.set_temporal_blocking(True, *geom)
    assert g.temporal_blocking_status().startswith("on"), g.temporal_blocking_status()
    done = 0
    for s in (1, 2, 3, 4, 11, 40, 101):
        g.LBM_multi_opt(s - done)
        ref.step(s - done)
        done = s
        got = g.download()
        for nm in ("f", "rho", "u"):
            assert np.array_equal(got[nm], getattr(ref, nm)), "%s %r t=%d %s: %s" % (name, geom, s, nm, first_diff(got[nm], getattr(ref, nm)))
        assert g.t == ref.t and g.omega == ref.omega
    assert g.stats()["fused_steps"] >= 90
    g.close(); ref.close()


def test_ineligible_cases_fall_back_to_the_one_step_path():
    case = CASES["cyl3d"]                      # x-normal inlet/outlet planes
    ref = port.PortGrid(case)
    g = luma_b200.GridObj(defs_from_case(case)).LBM_initGrid()
    g.set_temporal_blocking(True)
    assert not g.temporal_blocking_status().startswith("on")
    g.LBM_multi_opt(20); ref.step(20)
    assert np.array_equal(g.download()["f"], ref.f)
    assert g.stats()["fused_steps"] == 0
    g.close(); ref.close()


def test_forces_after_fused_sweeps_use_the_last_step():
    """the last step of a call is a plain step, so the momentum-exchange force still refers to it"""
    case = CASES["cav3d_32"]
    a = luma_b200.GridObj(defs_from_case(case)).LBM_initGrid()
    b = luma_b200.GridObj(defs_from_case(case)).LBM_initGrid()
    a.set_temporal_blocking(True, 8, 2, 5)
    a.LBM_multi_opt(25); b.LBM_multi_opt(25)
    assert np.array_equal(a.computeLiftDrag(), b.computeLiftDrag())
    a.close(); b.close()
