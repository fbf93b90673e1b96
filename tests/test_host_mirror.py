"""CPU tests of the host-side mirror (luma_b200.Definitions) and of the C-ABI surface.
No compute calls: there is no GPU here and the product has no CPU path."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import luma_b200
from luma_b200 import capi
from oracle import port
from oracle.cases import CASES
from util import defs_from_case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    L = capi.load()
    hdr = open(os.path.join(ROOT, "include", "luma_b200.h")).read()
    names = sorted(set(re.findall(r"\b(luma_b200_[a-z_0-9]+)\s*\(", hdr)))
    assert len(names) >= 18
    for nm in names:
        assert hasattr(L, nm), "symbol %s declared in include/luma_b200.h is not exported" % nm
    assert L.luma_b200_abi_version() == 4


def test_struct_layout_matches_header():
    p = capi.default_params()
    assert p.struct_size == C.sizeof(capi.LumaCaseParams)
    assert (p.dims, p.num_vels, p.nranks, p.regularised) == (3, 19, 1, 1)
    assert C.sizeof(capi.LumaSiteBC) == 16


def test_strerror_texts_follow_the_reference_messages():
    L = capi.load()
    assert L.luma_b200_strerror(0) == b"ok"
    # src/GridObj_ops_lbm_optimised.cpp:335, :358
    assert b"not within a wall" in L.luma_b200_strerror(capi.EBC_NOT_WALL)
    assert b"corner or an edge" in L.luma_b200_strerror(capi.EBC_PRESSURE_EDGE)


@pytest.mark.parametrize("N,G", [(512, 8), (512, 2), (100, 8), (1024, 8), (384 * 4, 4), (7, 2), (9, 4)])
def test_slab_follows_mpi_uniform_decompose(N, G):
    # MpiManager::mpi_uniformDecompose, src/MpiManager.cpp:1220-1240
    per = -(-N // G)
    last = per - (per * G - N)
    if last <= 0:
        per = N // G
        last = per - (per * G - N)
    total = 0
    for r in range(G):
        off, cnt = capi.slab(N, G, r)
        assert off == per * r
        assert cnt == (last if r == G - 1 else per)
        total += cnt
    assert total == N


def test_slab_rejects_bad_arguments():
    with pytest.raises(capi.LumaB200Error):
        capi.slab(4, 8, 0)     # last core would have <= 0 cells -> the reference L_ERRORs
    with pytest.raises(capi.LumaB200Error):
        capi.slab(16, 2, 2)


def test_create_rejects_inconsistent_cases_before_touching_cuda():
    L = capi.load()
    h = C.c_void_p()
    p = capi.default_params()
    p.N, p.M, p.K, p.x_count = 8, 8, 8, 8
    p.num_vels = 27                      # D3Q27 only together with L_USE_KBC_COLLISION (definitions.h:299-310) ...
    assert L.luma_b200_create(C.byref(h), C.byref(p)) == capi.EINVAL
    p.kbc = 1                            # ... and never with regularised boundaries (init_grids.cpp:266-270)
    assert p.regularised == 1
    assert L.luma_b200_create(C.byref(h), C.byref(p)) == capi.EINVAL
    p.num_vels = 19                      # KBC in 3-D means D3Q27
    assert L.luma_b200_create(C.byref(h), C.byref(p)) == capi.EINVAL
    p.kbc = 0
    p.omega = 2.5                        # init_grids.cpp:353-356
    assert L.luma_b200_create(C.byref(h), C.byref(p)) == capi.EINVAL
    p.omega = 1.0
    p.struct_size = 4
    assert L.luma_b200_create(C.byref(h), C.byref(p)) == capi.EINVAL


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.LumaB200Error) as e:
        luma_b200.GridObj(luma_b200.Definitions(L_RESOLUTION=8))
    assert e.value.code == capi.ECUDA


@pytest.mark.parametrize("name", list(CASES))
def test_definitions_derive_the_reference_scalars(name):
    """omega, nu, gravity, rho_out, inlet profiles, positions and wall descriptors as LUMA derives them
    (checked against the oracle, which is pinned to the compiled reference)."""
    case = CASES[name]
    d = defs_from_case(case)
    g = port.PortGrid(case)
    assert (d.L_N, d.L_M, d.L_K, d.L_NUM_VELS) == (case.N, case.M, case.K, case.Q)
    assert d.omega == g.omega and d.nu == g.nu
    assert d.gravity == g.gravity and d.rho_out == g.rho_out
    for a, b in zip(d.inlet_profiles(), (g.uin(0), g.uin(1), g.uin(2))):
        assert np.array_equal(a, b)
    x, y, z = d.positions()
    assert np.array_equal(x, g.pos(0)) and np.array_equal(y, g.pos(1))
    if case.dims == 3:
        assert np.array_equal(z, g.pos(2))
    lt = g.lattyp
    wall = g.wall.reshape(-1, 5)
    desc = d.boundary_site_descriptors(lt)
    assert len(desc) == int(np.isin(lt, (6, 7, 8)).sum())
    for site, ec, nd, n in desc[:: max(1, len(desc) // 500)]:
        assert (ec, nd, *n) == tuple(int(v) for v in wall[site]), (name, site)
    for s in (0, 1, 7, 100):
        assert d.velocity_ramp_coefficient(s * d.dt) == g.velocity_ramp(s * d.dt)
    g.close()


def test_mpi_branch_of_the_dropin_shim_compiles():
    """The L_BUILD_FOR_MPI branch of luma_b200/host/GridObj_ops_lbm_b200.cpp (one rank per GPU: MPI_Bcast of the NCCL id,
    node-local device choice, collective choice of the halo transport) compiled against the unmodified LUMA headers and the
    serial <mpi.h> stand-in.  oracle/shim/mpi.h explains how L_BUILD_FOR_MPI is raised at the point where a real LUMA build
    sees it (inc/stdafx.h:203-206).  Needs the reference sources; the image has no MPI to link and run it."""
    import subprocess
    if not os.path.exists("/root/reference/LUMA/inc/stdafx.h"):
        pytest.skip("reference sources not present")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [cxx, "-O0", "-std=c++0x", "-w", "-include", "limits", "-I" + os.path.join(ROOT, "oracle", "shim"), "-I/root/reference",
           "-include", os.path.join(ROOT, "oracle", "cases", "chan3d.h"), "-I" + os.path.join(ROOT, "include"),
           "-DLUMA_SHIM_MPI_SYNTAX", "-c", os.path.join(ROOT, "luma_b200", "host", "GridObj_ops_lbm_b200.cpp"), "-o", "/dev/null"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    assert r.returncode == 0, r.stdout.decode()[-3000:]
    pre = subprocess.run(cmd[:-4] + ["-E", cmd[-3]], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout.decode()
    assert "MPI_Allgather" in pre and "MPI_Comm_split_type" in pre and "luma_b200_p2p_attach" in pre      # the branch was really compiled
