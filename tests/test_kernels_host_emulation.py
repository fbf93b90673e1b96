"""Logic check of the device code without a GPU: the kernel bodies of luma_b200/csrc/kernels_impl.cuh
(k_cell_words, k_synthetic, k_step, k_bc with the regularised / general / KBC paths, k_velsrc) are compiled for
the host by a test-only harness (tests/harness/kernels_host.cpp, driven by tests/emu.py) that runs them one
"thread" at a time, and the parity cases are replayed against the oracle, bit for bit:

* every case as one slab (upload path), and from the device-side initialisation (k_synthetic);
* the multi-GPU cases as 2 and 3 x-slabs with ghost planes, face / interior launches as luma_b200_step issues them,
  and the exchange driven by the PRODUCT's own halo plan (luma_b200_halo_plan through the C ABI -- pure host logic).

This is an early warning for kernel edits made where no GPU is at hand, and it covers the slab logic on boxes with a
single GPU; the parity tests proper are the `-m gpu` ones, which go through the C ABI on a B200.  The harness is not a
CPU implementation of the product (no ABI, no streams, no transport) and nothing in luma_b200/ can reach it.
"""
import numpy as np
import pytest

import emu
from luma_b200 import ring
from oracle import port
from oracle.cases import CASES
from util import defs_from_case

pytestmark = pytest.mark.skipif(not emu.available(), reason="CUDA headers not installed (the harness needs their host-side type declarations)")
SITE_STEP_BUDGET = 7.0e6        # emulated site updates per case (keeps the CPU suite within minutes)


def _compare(name, tag, got, ref, sl=slice(None)):
    case = ref.case
    Q, D = case.Q, case.dims
    for nm, w in (("f", Q), ("rho", 1), ("u", D)) + ((("rho_timeav", 1), ("ui_timeav", D), ("uiuj_timeav", 3 * D - 3)) if case.time_averaged else ()):
        a = got[nm]
        b = getattr(ref, nm).reshape(-1, w)[sl].reshape(-1)
        bad = np.flatnonzero(a != b)
        assert bad.size == 0, "%s %s %s: %d differ, first at %d: %r vs %r" % (name, tag, nm, bad.size, bad[0], a[bad[0]], b[bad[0]])


def _snaps(case, cells, cap=None):
    s = [x for x in case.steps if x * cells <= SITE_STEP_BUDGET and (cap is None or x <= cap)]
    return s or [case.steps[0]]


@pytest.mark.parametrize("name", [n for n in CASES if n != "cyl3d_ld"])   # cyl3d_ld is cyl3d with another output cadence
def test_one_slab_upload_path(name):
    case = CASES[name]
    ref = port.PortGrid(case)
    s = emu.Slab(case, ref).upload_from(ref).finalize()
    for snap in _snaps(case, s.cells):
        while s.t < snap:
            ref.step(1)
            s.set_scalars(ref, ref.omega)          # the omega this step ran with (Reynolds ramp)
            s.step_all()
            s.advance()
        s.velsrc()
        _compare(name, "t%d" % s.t, s.owned(), ref)
    ref.close()


@pytest.mark.parametrize("name", [n for n in CASES if n != "cyl3d_ld"])   # cyl3d_ld is cyl3d with another output cadence
def test_device_init_path(name):
    """k_synthetic = LBM_initGrid + LBM_initBoundLab + body labelling in index space"""
    case = CASES[name]
    ref = port.PortGrid(case)
    s = emu.Slab(case, ref).init_synthetic(ref)
    got = s.owned()
    assert np.array_equal(got["types"], ref.lattyp.astype(np.uint8))
    wall = ref.wall.reshape(-1, 5)
    bc = np.flatnonzero(np.isin(got["types"], (6, 7, 8)) & (wall[:, 0] > 0))
    want = np.array([emu.pack_desc(int(wall[i, 0]), int(wall[i, 1]), [int(x) for x in wall[i, 2:5]]) for i in bc], dtype=np.uint32)
    assert np.array_equal(got["desc"][bc], want)
    assert np.array_equal(s.f[0], s.f[1])
    _compare(name, "init", got, ref)
    s.finalize()
    for snap in _snaps(case, s.cells, cap=10):
        while s.t < snap:
            ref.step(1)
            s.set_scalars(ref, ref.omega)
            s.step_all()
            s.advance()
        s.velsrc()
        _compare(name, "t%d" % s.t, s.owned(), ref)
    ref.close()


MULTI = ["chan3d", "cyl3d", "chan2d", "cyl2d", "slipchan3d", "sliptunnel2d", "sliptunnel3d", "fevel2d", "fevel3d", "fevel2d_tav",
         "tunnel2d_tav", "cav3d_32", "cav3d_tav", "felid3d", "kbc2d_cyl", "kbc3d_chan"]      # = tests/test_gpu_multi.py


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("name", MULTI)
def test_slabs_with_ghost_planes_and_the_products_halo_plan(name, world):
    case = CASES[name]
    defs = defs_from_case(case)
    ref = port.PortGrid(case)
    plans = [ring.halo_plan(defs, r, world) for r in range(world)]
    slabs = [emu.Slab(case, ref, r, world) for r in range(world)]
    for s in slabs:
        (s.upload_from(ref) if world == 2 else s.init_synthetic(ref)).finalize()
    emu.exchange(slabs, plans, 0)                   # luma_b200_upload ends with an exchange of the current lattice
    MK = case.M * case.K
    steps = 12 if not case.kbc else 8
    if case.N * MK * steps > SITE_STEP_BUDGET:
        steps = 4
    for t in range(steps):
        ref.step(1)
        for s in slabs:
            s.set_scalars(ref, ref.omega)
            s.step_faces()
        emu.exchange(slabs, plans, 1)               # the lattice just written
        for s in slabs:
            s.step_interior()
            s.advance()
        if t in (0, 1, steps - 1):
            for s in slabs:
                s.velsrc()
                _compare(name, "t%d rank %d/%d" % (t + 1, s.rank, world), s.owned(), ref, slice(s.x0 * MK, (s.x0 + s.cnt) * MK))
    ref.close()


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("name", MULTI)
def test_fused_halo_stores_equal_the_plan_driven_exchange(name, world):
    """LUMA_B200_FUSED_HALO: k_step_faces / k_bc store the outgoing populations into the neighbours' ghost planes in their
    epilogue; no copy afterwards.  Same bits as the oracle, hence as the plan-driven exchange."""
    case = CASES[name]
    defs = defs_from_case(case)
    ref = port.PortGrid(case)
    plans = [ring.halo_plan(defs, r, world) for r in range(world)]
    slabs = [emu.Slab(case, ref, r, world) for r in range(world)]
    for s in slabs:
        (s.upload_from(ref) if world == 3 else s.init_synthetic(ref)).finalize()
    emu.exchange(slabs, plans, 0)                   # the exchange that ends upload / init is the plane copy in both modes
    MK = case.M * case.K
    steps = 10 if not case.kbc else 8
    if case.N * MK * steps > SITE_STEP_BUDGET:
        steps = 4
    for t in range(steps):
        ref.step(1)
        for s in slabs:
            s.set_scalars(ref, ref.omega)
        for r, s in enumerate(slabs):
            s.step_faces(left=slabs[(r - 1) % world], right=slabs[(r + 1) % world])
        for s in slabs:
            s.step_interior()
            s.advance()
        if t in (0, 1, steps - 1):
            for s in slabs:
                s.velsrc()
                _compare(name, "fused t%d rank %d/%d" % (t + 1, s.rank, world), s.owned(), ref, slice(s.x0 * MK, (s.x0 + s.cnt) * MK))
    ref.close()
