"""Logic check of the device code without a GPU: the kernel bodies of luma_b200/csrc/kernels_impl.cuh
(k_cell_words, k_step, k_bc with the regularised / general / KBC paths, k_velsrc) are compiled for the host by a
test-only harness (tests/harness/kernels_host.cpp) that runs them one "thread" at a time, and every parity case
is replayed against the oracle, bit for bit.  The boundary-site lists are built here the way
finalize_geometry() (luma_b200/csrc/api.cu) builds them.

This is an early warning for kernel edits made where no GPU is at hand; the parity tests proper are the
`-m gpu` ones, which go through the C ABI on a B200.  The harness is not a CPU implementation of the product
(single slab, no ABI, no streams, no exchange) and nothing in luma_b200/ can reach it.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import port
from oracle.cases import CASES

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "harness", "kernels_host.cpp")
LIB = os.path.join(HERE, "harness", "libkernels_host.so")
CSRC = os.path.join(os.path.dirname(HERE), "luma_b200", "csrc")
CUDA_INC = "/usr/local/cuda/include"
SITE_STEP_BUDGET = 7.0e6        # emulated site updates per case (keeps the CPU suite within minutes)


class EmuCase(C.Structure):
    _fields_ = [("Q", C.c_int32), ("D", C.c_int32), ("P", C.c_int32), ("M", C.c_int32), ("K", C.c_int32),
                ("regularised", C.c_int32), ("coll", C.c_int32), ("force", C.c_int32), ("gravity_dir", C.c_int32),
                ("velramp_on", C.c_int32), ("general", C.c_int32),
                ("omega", C.c_double), ("rhoin", C.c_double), ("rho_out", C.c_double), ("gravity", C.c_double), ("csmag", C.c_double),
                ("ramp", C.c_double), ("ramp_t", C.c_double), ("t_now", C.c_double), ("t_next", C.c_double)]


@pytest.fixture(scope="module")
def emu():
    if not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")):
        pytest.skip("CUDA headers not installed (the harness only needs their host-side type declarations)")
    deps = [SRC] + [os.path.join(CSRC, n) for n in ("kernels_impl.cuh", "kernels.cuh", "lattice.cuh")]
    if (not os.path.exists(LIB)) or os.path.getmtime(LIB) < max(os.path.getmtime(d) for d in deps):
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.run([cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-w",
                        "-I", CUDA_INC, "-o", LIB, SRC], check=True)
    L = C.CDLL(LIB)
    for nm in ("emu_cell_words", "emu_step", "emu_velsrc", "emu_lattice_c", "emu_class_shift"):
        getattr(L, nm).restype = C.c_int
    return L


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _pack_desc(ec, nd, n):
    return ((ec & 3) << 30) | ((nd & 3) << 22) | (((n[0] + 1) | ((n[1] + 1) << 2) | ((n[2] + 1) << 4)) << 24)


def _lists(case, types, desc, c):
    """finalize_geometry() of api.cu for one rank: sites k_bc handles, class-4 fluid sites, extra advances of
    the time averages, forced-equilibrium inlet sites."""
    N, M, K, Q, D = case.N, case.M, case.K, case.Q, case.dims
    reg = case.regularised
    sid = lambda i, j, k: (i * M + j) * K + k
    never = lambda t: t in (0, 2) or (t == 6 and not reg)
    lst, forced, extra, vel, general = [], set(), {}, [], False
    for i, j, k in np.argwhere(types > 1):
        i, j, k = int(i), int(j), int(k)
        t = int(types[i, j, k])
        me = sid(i, j, k)
        if t == 9 or (t == 6 and not reg):
            general = True
            for v in range(Q):
                di, dj, dk = (i + c[v][0]) % N, (j + c[v][1]) % M, (k + c[v][2]) % K
                dt = int(types[di, dj, dk])
                if never(dt):
                    continue
                if dt == 1:
                    forced.add(sid(di, dj, dk))
        if t == 8:
            general = True
            lst.append(me)
            continue
        if t == 9:
            lst.append(me)
            continue
        if not reg:
            if t == 7:
                general = True
                lst.append(me)
            else:
                vel.append(me)
            continue
        d = int(desc[i, j, k])
        ec = d >> 30
        assert ec >= 1
        if ec > 1 or t == 7:
            n = [((d >> (24 + 2 * a)) & 3) - 1 for a in range(3)]
            ncalls = (D - 1) if t == 7 else 1
            for m in (1, 2):
                idn = sid(i + m * n[0], j + m * n[1], k + m * n[2])
                if idn > me and case.time_averaged:
                    extra[idn] = extra.get(idn, 0) + ncalls
                    forced.add(idn)
        lst.append(me)
    full = sorted(lst + sorted(forced))
    extra_by_entry = None
    if extra:
        extra_by_entry = np.zeros(len(full), dtype=np.int32)
        pos = {s: a for a, s in enumerate(full)}
        for s, e in extra.items():
            extra_by_entry[pos[s]] += e
    return (np.array(full, dtype=np.int64), np.array(sorted(forced), dtype=np.int64), extra_by_entry,
            np.array(vel, dtype=np.int64), general or bool(forced))


@pytest.mark.parametrize("name", list(CASES))
def test_kernel_bodies_replay_the_oracle(emu, name):
    case = CASES[name]
    ref = port.PortGrid(case)
    N, M, K, Q, D = case.N, case.M, case.K, case.Q, case.dims
    cells = N * M * K
    stride = (cells + 15) // 16 * 16
    c = [[emu.emu_lattice_c(Q, v, d) for d in range(3)] for v in range(Q)]

    types = ref.lattyp.astype(np.uint8)
    wall = ref.wall.reshape(cells, 5)
    desc = np.zeros(cells, dtype=np.uint32)
    for s in np.flatnonzero(np.isin(types, (6, 7, 8)) & (wall[:, 0] > 0)):      # what upload receives: LumaSiteBC of V/P/slip sites
        desc[s] = _pack_desc(int(wall[s, 0]), int(wall[s, 1]), [int(x) for x in wall[s, 2:5]])
    bc_list, forced, bc_extra, vel, general = _lists(case, types.reshape(N, M, K), desc.reshape(N, M, K), c)

    p = EmuCase()
    p.Q, p.D, p.P, p.M, p.K = Q, D, N, M, K
    p.regularised = int(case.regularised)
    p.coll = 2 if case.kbc else (1 if case.bgksmag else 0)
    p.force, p.gravity_dir = int(case.gravity_on), case.gravity_dir
    p.velramp_on = int(case.velocity_ramp is not None)
    p.general = int(general)
    p.rhoin, p.rho_out, p.gravity, p.csmag = 1.0, ref.rho_out, ref.gravity, case.csmag

    cw = np.zeros(cells, dtype=np.uint32)
    assert emu.emu_cell_words(C.byref(p), _ptr(types), _ptr(desc), _ptr(cw)) == 0
    shift = emu.emu_class_shift(Q)
    for s in forced:                                   # k_force_general
        if (int(cw[s]) >> shift) & 7 == 1:
            cw[s] = (int(cw[s]) & ~(7 << shift)) | (4 << shift)
    # every listed site is of class 2, 3 or 4 and every such site is listed; fluid sites are class 1
    cls = (cw >> np.uint32(shift)) & np.uint32(7)
    assert np.array_equal(np.flatnonzero(cls >= 2), bc_list)
    assert ((cls == 1) | (cls == 4))[types == 1].all() and (cls[types == 0] == 0).all()

    f = [np.zeros((Q, stride)), None]
    f[0][:, :cells] = ref.f.reshape(cells, Q).T
    f[1] = f[0].copy()
    rho = np.zeros(stride); rho[:cells] = ref.rho
    u = np.zeros((D, stride)); u[:, :cells] = ref.u.reshape(cells, D).T
    uin = np.ascontiguousarray(np.stack([ref.uin(0), ref.uin(1), ref.uin(2)]))
    tav = np.zeros((1 + D + 3 * D - 3, stride)) if case.time_averaged else None

    snaps = [s for s in case.steps if s * cells <= SITE_STEP_BUDGET] or [case.steps[0]]
    cur, t = 0, 0
    for snap in snaps:
        while t < snap:
            ref.step(1)
            p.omega = ref.omega                              # the omega this step ran with (Reynolds ramp)
            p.ramp, p.ramp_t = ref.velocity_ramp((t + 1) * case.dt), ref.velocity_ramp(t * case.dt)
            p.t_now, p.t_next = float(t), float(t + 1)
            rc = emu.emu_step(C.byref(p), _ptr(f[cur]), _ptr(f[cur ^ 1]), _ptr(cw), _ptr(rho), _ptr(u), C.c_longlong(stride),
                              _ptr(bc_list), _ptr(bc_extra), C.c_int(len(bc_list)), _ptr(uin), _ptr(types), _ptr(desc), _ptr(tav))
            assert rc == 0
            cur ^= 1
            t += 1
        if (not case.regularised) and case.velocity_ramp is not None and len(vel):
            emu.emu_velsrc(C.byref(p), _ptr(vel), C.c_int(len(vel)), _ptr(types), _ptr(desc), _ptr(u), C.c_longlong(stride), _ptr(uin), C.c_int(N))
        got_f = np.ascontiguousarray(f[cur][:, :cells].T).reshape(-1)
        for nm, a, b in (("f", got_f, ref.f), ("rho", rho[:cells], ref.rho), ("u", np.ascontiguousarray(u[:, :cells].T).reshape(-1), ref.u)):
            bad = np.flatnonzero(a != b)
            assert bad.size == 0, "%s t=%d %s: %d differ, first at %d: %r vs %r" % (name, t, nm, bad.size, bad[0], a[bad[0]], b[bad[0]])
        if tav is not None:
            got = {"rho_timeav": tav[0, :cells], "ui_timeav": np.ascontiguousarray(tav[1:1 + D, :cells].T).reshape(-1),
                   "uiuj_timeav": np.ascontiguousarray(tav[1 + D:, :cells].T).reshape(-1)}
            for nm, a in got.items():
                assert np.array_equal(a, getattr(ref, nm)), (name, t, nm)
    ref.close()
