"""BASELINE.json configs[2] and configs[3] at FULL size on one B200.  The oracle cannot replay 134 M /
67 M cells in seconds, so these check size-independent properties the cases offer (the reduced shapes
chan3d / cyl3d carry the bit-for-bit parity):

* periodic channel 512^3 (configs[2]): every site of a (y = const) line runs the same operations on the
  same values, so the fields must be bitwise independent of x and z; bounce-back + periodic wrap + Guo
  forcing conserve mass, so sum(rho) stays at its initial value to round-off; the profile is symmetric in y.
* cylinder 1024x256x256 (configs[3]): the body spans the periodic z direction, so the fields must be
  bitwise independent of z; two runs are bit-identical (no atomics, no order dependence).

1 GPU <-> N GPU equality at these sizes is in tests/test_gpu_multi.py (needs >= 2 GPUs).
"""
import os
import sys

import numpy as np
import pytest

import luma_b200
from luma_b200 import capi

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402  (only for the workload table shared with bench.py)

pytestmark = pytest.mark.gpu


def _free_gb():
    import torch
    free, _ = torch.cuda.mem_get_info(0)
    return free / 2 ** 30


def test_full_size_c3_channel_invariance_and_mass():
    if _free_gb() < 60:
        pytest.skip("needs ~50 GB of device memory")
    d = bench.workload_defs("c3", 1)
    N, M, K = d.L_N, d.L_M, d.L_K
    assert (N, M, K) == (512, 512, 512)
    g = luma_b200.GridObj(d).LBM_initGrid()
    lt = g.LatTyp.reshape(N, M, K)
    fluid = lt == 1
    assert fluid[:, 1:-1, :].all() and not fluid[:, 0, :].any() and not fluid[:, -1, :].any()
    mass0 = float(fluid.sum()) * d.L_RHOIN
    g.LBM_multi_opt(40)
    out = g.download(capi.RHO | capi.U)
    rho = out["rho"].reshape(N, M, K)
    u = out["u"].reshape(N, M, K, 3)
    # bitwise independence of x and z
    assert np.array_equal(rho, np.broadcast_to(rho[:1, :, :1], rho.shape))
    assert np.array_equal(u, np.broadcast_to(u[:1, :, :1, :], u.shape))
    # mass
    mass = float(rho[fluid].sum(dtype=np.longdouble))
    assert abs(mass - mass0) <= 1e-12 * mass0, (mass, mass0)
    # the force drives +x only, symmetric in y (direction numbering is not mirror symmetric: round-off allowed)
    prof = u[0, :, 0, 0]
    assert (prof[1:-1] > 0).all() and np.max(np.abs(prof[1:-1] - prof[-2:0:-1])) < 1e-15
    assert np.max(np.abs(u[0, :, 0, 2])) == 0.0
    assert g.stats()["mlups_last_call"] > 1000
    g.close()


def test_full_size_c4_cylinder_z_invariance_and_determinism():
    if _free_gb() < 30:
        pytest.skip("needs ~25 GB of device memory")
    d = bench.workload_defs("c4", 1)
    N, M, K = d.L_N, d.L_M, d.L_K
    assert (N, M, K) == (1024, 256, 256)
    g = luma_b200.GridObj(d).LBM_initGrid()
    lt = g.LatTyp.reshape(N, M, K)
    assert (lt[256:288, 112:144, :] == 0).all() and (lt[0, 1:-1, :] == 6).all() and (lt[-1, 1:-1, :] == 7).all()
    g.LBM_multi_opt(30)
    out = g.download(capi.RHO | capi.U)
    F = g.computeLiftDrag()
    rho = out["rho"].reshape(N, M, K)
    u = out["u"].reshape(N, M, K, 3)
    assert np.isfinite(rho).all()
    assert np.array_equal(rho, np.broadcast_to(rho[:, :, :1], rho.shape))
    assert np.array_equal(u, np.broadcast_to(u[:, :, :1, :], u.shape))
    assert np.max(np.abs(u[..., 2])) == 0.0
    assert F[0] > 0.0 and abs(F[2]) == 0.0          # drag along +x, nothing along the span
    g.close()
    h = luma_b200.GridObj(d).LBM_initGrid()
    h.LBM_multi_opt(30)
    again = h.download(capi.RHO | capi.U)
    assert np.array_equal(again["rho"], out["rho"]) and np.array_equal(again["u"], out["u"])
    assert np.array_equal(h.computeLiftDrag(), F)
    h.close()


def test_beyond_int32_element_offsets():
    """640^3 cells: 4.98e9 population elements per lattice -- element offsets exceed 2^32, which the reference's
    `int` indexing cannot address.  Same invariance / mass properties as the 512^3 channel."""
    if _free_gb() < 110:
        pytest.skip("needs ~95 GB of device memory")
    res = 640
    d = luma_b200.Definitions(
        L_DIMS=3, L_RESOLUTION=res, L_TIMESTEP=0.05 / res, L_RE=None, L_NU=1.0 / res, L_NO_FLOW=True,
        L_WALL_LEFT=luma_b200.eFluid, L_WALL_RIGHT=luma_b200.eFluid, L_WALL_FRONT=luma_b200.eFluid,
        L_WALL_BACK=luma_b200.eFluid, L_WALL_THICKNESS_CELLS=(0, 0, 1, 1, 0, 0),
        L_GRAVITY_ON=True, L_GRAVITY_FORCE=0.0158, L_GRAVITY_DIRECTION=0)
    assert d.L_N * d.L_M * d.L_K * 19 > 2 ** 32
    g = luma_b200.GridObj(d).LBM_initGrid()
    g.LBM_multi_opt(12)
    rho = g.download(capi.RHO)["rho"].reshape(res, res, res)
    assert np.array_equal(rho, np.broadcast_to(rho[:1, :, :1], rho.shape))
    line = rho[0, 1:-1, 0]
    assert abs(float(line.sum(dtype=np.longdouble)) - (res - 2)) <= 1e-12 * res
    del rho
    ux = g.download(capi.U)["u"].reshape(res, res, res, 3)[..., 0]
    assert np.array_equal(ux, np.broadcast_to(ux[:1, :, :1], ux.shape))
    assert (ux[0, 1:-1, 0] > 0).all()
    g.close()
