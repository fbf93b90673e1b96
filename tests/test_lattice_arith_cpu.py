"""The product's per-site arithmetic (luma_b200/csrc/lattice.cuh: macroscopic, equilibrium_all, kbc_collide,
guo_force, the D2Q9 / D3Q19 / D3Q27 tables) compiled for the host by a test-only harness
(tests/harness/lattice_host.cpp, g++ -ffp-contract=off) and checked bit-for-bit against the oracle on every
fluid site whose neighbours are all fluid.  This is a CPU-side early warning for the arithmetic the GPU
parity tests (-m gpu) pin end to end; it never runs in the product (there is no CPU path)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import port
from oracle.cases import CASES

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "harness", "lattice_host.cpp")
LIB = os.path.join(HERE, "harness", "liblattice_host.so")
HDR = os.path.join(os.path.dirname(HERE), "luma_b200", "csrc", "lattice.cuh")


@pytest.fixture(scope="module")
def harness():
    if (not os.path.exists(LIB)) or os.path.getmtime(LIB) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.run([cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-o", LIB, SRC], check=True)
    L = C.CDLL(LIB)
    dp = C.POINTER(C.c_double)
    L.lattice_host_sites.argtypes = [C.c_int, C.c_longlong, dp, dp, C.c_int, C.c_int, C.c_double, dp, dp, dp, dp]
    L.lattice_host_sites.restype = C.c_int
    L.lattice_host_c.restype = C.c_int
    return L


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


@pytest.mark.parametrize("name,steps", [("kbc2d", 40), ("kbc2d_cyl", 25), ("kbc3d_chan", 12), ("kbc3d", 8), ("chan3d", 20),
                                        ("chan3d_gz", 10), ("torus2d", 30)])
def test_site_arithmetic_matches_oracle(harness, name, steps):
    case = CASES[name]
    g = port.PortGrid(case)
    N, M, K, Q, D = case.N, case.M, case.K, case.Q, case.dims
    c = np.array([[harness.lattice_host_c(Q, v, d) for d in range(3)] for v in range(Q)])
    lt = g.lattyp.reshape(N, M, K)
    fluid = lt == 1
    # sites whose every source site (periodic wrap included) is eFluid: plain pull, macro, collide
    ok = fluid.copy()
    for v in range(Q):
        ok &= np.roll(fluid, shift=tuple(c[v]), axis=(0, 1, 2))
    assert ok.sum() > 50
    F = np.zeros(3)
    if case.gravity_on:
        F[case.gravity_dir] = 1.0 * g.gravity * 1.0          # rho_init * gravity * refinement ratio
    checked = 0
    for s in range(steps):
        f0 = g.f.reshape(N, M, K, Q)
        fp = np.stack([np.roll(f0[..., v], shift=tuple(c[v]), axis=(0, 1, 2)) for v in range(Q)], axis=-1)
        omega = g.omega
        g.step(1)
        pulled = np.ascontiguousarray(fp[ok])
        own = np.ascontiguousarray(f0[ok])
        n = pulled.shape[0]
        out = np.empty((n, Q)); rho = np.empty(n); u = np.empty((n, D))
        assert harness.lattice_host_sites(Q, n, _p(pulled), _p(own), int(case.kbc), int(case.gravity_on), omega, _p(F), _p(out), _p(rho), _p(u)) == 0
        f1 = g.f.reshape(N, M, K, Q)[ok]
        assert np.array_equal(rho, g.rho.reshape(N, M, K)[ok]), (name, s, "rho")
        assert np.array_equal(u, g.u.reshape(N, M, K, D)[ok]), (name, s, "u")
        bad = np.flatnonzero((out != f1).any(axis=1))
        assert bad.size == 0, (name, s, "f", bad.size, out[bad[0]], f1[bad[0]])
        checked += n
    assert checked > 0
    g.close()
