"""Seeded fuzzing of the device code's logic on the CPU: random (but valid) LUMA cases -- ragged grid sizes, wall
type patterns taken from the menus the fixed cases use, random bodies, forcing, ramps, collision operators, time
averages -- are run through the oracle and through the host-emulated kernel bodies (tests/emu.py), one slab and
2-3 slabs, and must agree bit for bit.  Test infrastructure only; see tests/test_kernels_host_emulation.py."""
import random

import numpy as np
import pytest

import emu
from luma_b200 import ring
from oracle import port
from oracle.cases import Case, E_SOLID as S, E_FLUID as F, E_VELOCITY as V, E_PRESSURE as P, E_SLIP as SL, E_EXTRAPOLATE_RIGHT as X
from util import defs_from_case

pytestmark = pytest.mark.skipif(not emu.available(), reason="CUDA headers not installed")

# (walls L,R,B,T,Fr,Bk ; thickness ; regularised) patterns that are well defined in the reference (no boundary site
# extrapolating from another boundary site, no pressure edges): the fixed cases' patterns
PATTERNS_2D = [
    ((S, S, S, V, S, S), (1, 1, 1, 1, 1, 1), True),
    ((F, F, S, S, F, F), (0, 0, 1, 1, 0, 0), True),
    ((F, F, S, SL, F, F), (0, 0, 1, 1, 0, 0), True),
    ((V, P, S, S, F, F), (1, 1, 1, 1, 0, 0), True),
    ((V, P, SL, SL, F, F), (1, 1, 1, 1, 0, 0), True),
    ((V, P, V, V, F, F), (1, 1, 1, 1, 1, 1), True),
    ((V, X, S, S, F, F), (1, 1, 1, 1, 0, 0), False),
    ((V, X, SL, S, F, F), (1, 1, 1, 1, 0, 0), False),
    ((V, P, S, S, F, F), (1, 1, 1, 1, 0, 0), False),
    ((F, F, F, F, F, F), (0, 0, 0, 0, 0, 0), True),
    ((F, F, S, S, F, F), (0, 0, 2, 1, 0, 0), True),
    ((P, V, S, S, F, F), (1, 1, 1, 1, 0, 0), True),
    ((S, S, V, S, S, S), (1, 1, 1, 1, 1, 1), True),
    ((V, S, S, S, S, S), (1, 1, 1, 1, 1, 1), True),
    ((S, V, S, S, S, S), (1, 1, 1, 1, 1, 1), False),
    ((V, P, S, SL, F, F), (1, 1, 3, 1, 0, 0), True),
    ((S, S, S, S, S, S), (2, 1, 1, 2, 1, 1), True),
]
PATTERNS_3D = [
    ((S, S, S, V, S, S), (1, 1, 1, 1, 1, 1), True),
    ((F, F, S, S, F, F), (0, 0, 1, 1, 0, 0), True),
    ((F, F, SL, S, F, F), (0, 0, 1, 1, 0, 0), True),
    ((V, P, S, S, F, F), (1, 1, 1, 1, 0, 0), True),
    ((V, P, S, S, S, S), (1, 1, 1, 1, 1, 1), True),
    ((V, P, SL, SL, SL, SL), (1, 1, 1, 1, 1, 1), True),
    ((V, P, V, V, V, V), (1, 1, 1, 1, 1, 1), True),
    ((V, X, S, S, F, F), (1, 1, 1, 1, 0, 0), False),
    ((V, X, SL, SL, F, F), (1, 1, 1, 1, 0, 0), False),
    ((S, S, S, V, S, S), (1, 1, 1, 1, 1, 1), False),
    ((V, P, S, S, S, S), (1, 1, 1, 1, 1, 1), False),
    ((F, F, S, S, F, F), (0, 0, 2, 2, 0, 0), True),
    ((P, V, S, S, F, F), (1, 1, 1, 1, 0, 0), True),
    ((S, S, S, S, V, S), (1, 1, 1, 1, 1, 1), True),
    ((S, S, S, S, S, V), (1, 1, 1, 1, 1, 1), False),
    ((V, S, S, S, S, S), (1, 1, 1, 1, 1, 1), True),
    ((F, F, S, S, SL, SL), (0, 0, 1, 2, 1, 1), True),
    ((V, P, S, S, SL, S), (1, 1, 2, 1, 1, 3), True),
    ((F, F, F, F, F, F), (0, 0, 0, 0, 0, 0), True),
]


def random_case(seed):
    rnd = random.Random(seed)
    dims = rnd.choice((2, 3))
    walls, thick, reg = rnd.choice(PATTERNS_3D if dims == 3 else PATTERNS_2D)
    res = rnd.choice((6, 7, 8, 9, 10, 12)) if dims == 3 else rnd.choice((8, 10, 11, 13, 16))
    # extents in cells: x long enough for 3 slabs of >= 4 planes and for two-plane extrapolation
    nx = rnd.randint(13, 20)
    ny = rnd.randint(6, 12)
    nz = rnd.randint(5, 10) if dims == 3 else 1
    kbc = rnd.random() < 0.2
    if kbc and dims == 3:
        reg = False                                    # the reference refuses regularised boundaries on D3Q27
        if not any(p[0] == walls and p[2] is False for p in PATTERNS_3D):
            walls, thick, reg = rnd.choice([p for p in PATTERNS_3D if p[2] is False])
    bgksmag = (not kbc) and rnd.random() < 0.4
    gravity = rnd.random() < 0.5
    has_inlet = V in walls
    box = None
    if rnd.random() < 0.6:
        if rnd.random() < 0.7:
            i0 = rnd.randint(4, nx - 7)
            j0 = rnd.randint(2, ny - 4)
            k0 = rnd.randint(0, max(nz - 3, 0)) if dims == 3 else 0
        else:                                          # anywhere on the grid: touching walls and the periodic faces
            i0 = rnd.randint(0, nx - 1)
            j0 = rnd.randint(0, ny - 1)
            k0 = rnd.randint(0, nz - 1) if dims == 3 else 0
        box = (i0, min(i0 + rnd.randint(1, 3), nx), j0, min(j0 + rnd.randint(1, 2), ny), k0, min(k0 + rnd.randint(1, 3), nz) if dims == 3 else 1)
    return Case("fuzz%d" % seed, dims, res, (nx + 0.5) / res, (ny + 0.5) / res, (nz + 0.5) / res if dims == 3 else 1.0,
                timestep="0.05/%d.0" % res, walls=walls, thick=thick,
                ux0=rnd.choice((1.0, 0.7, -1.0)) if walls[0] != V else 1.0, uy0=rnd.choice((0.0, 0.0, 0.3)) if not has_inlet else 0.0,
                uz0=(rnd.choice((0.0, 0.2)) if dims == 3 else 0.0), re=rnd.choice((8.0, 20.0, 50.0)), regularised=reg,
                no_flow=rnd.random() < 0.5, bgksmag=bgksmag, kbc=kbc, csmag=rnd.choice((0.17, 0.3)),
                gravity_on=gravity, gravity_force=rnd.choice((0.1, 0.3)), gravity_dir=rnd.randrange(dims),
                velocity_ramp=rnd.choice((None, 0.05, 0.2)) if has_inlet else None,
                parabolic_inlet=has_inlet and walls[0] == V and rnd.random() < 0.3,
                pressure_delta=rnd.choice((0.0, 0.0, 0.3)), time_averaged=rnd.random() < 0.4, box=box, ld_out=box is not None,
                steps=(1, 3, 6), doc="seeded random case")


def _cmp(tag, got, ref, sl=slice(None)):
    case = ref.case
    Q, D = case.Q, case.dims
    for nm, w in (("f", Q), ("rho", 1), ("u", D)) + ((("rho_timeav", 1), ("ui_timeav", D), ("uiuj_timeav", 3 * D - 3)) if case.time_averaged else ()):
        a, b = got[nm], getattr(ref, nm).reshape(-1, w)[sl].reshape(-1)
        same = (a == b) | (np.isnan(a) & np.isnan(b))
        assert same.all(), "%s %s: %d differ, first at %d: %r vs %r" % (tag, nm, int((~same).sum()), int(np.flatnonzero(~same)[0]),
                                                                       a[np.flatnonzero(~same)[0]], b[np.flatnonzero(~same)[0]])


@pytest.mark.parametrize("seed", range(int(__import__("os").environ.get("LUMA_FUZZ_SEEDS", "100"))))
def test_random_case_one_slab_and_slabs(seed):
    case = random_case(seed)
    assert int(case.bx * case.resolution) == case.N >= 13
    try:
        ref = port.PortGrid(case)
    except RuntimeError:
        pytest.skip("the reference rejects this case (omega >= 2)")
    world = 2 + seed % 2
    defs = defs_from_case(case)
    plans = [ring.halo_plan(defs, r, world) for r in range(world)]
    try:
        one = emu.Slab(case, ref).upload_from(ref).finalize()
        slabs = [emu.Slab(case, ref, r, world) for r in range(world)]
        for s in slabs:
            (s.init_synthetic(ref) if seed % 3 else s.upload_from(ref)).finalize()
    except emu.Rejected as e:
        pytest.skip("the product refuses this case, as the reference is ill defined on it: %s" % e)
    emu.exchange(slabs, plans, 0)
    MK = case.M * case.K
    for t in range(6):
        try:
            ref.step(1)
        except RuntimeError as e:
            pytest.skip("the reference stops on this case: %s" % e)
        one.set_scalars(ref, ref.omega)
        one.step_all()
        one.advance()
        for s in slabs:
            s.set_scalars(ref, ref.omega)
        if seed % 4 < 2:
            for s in slabs:
                s.step_faces()
            emu.exchange(slabs, plans, 1)
        else:                                           # fused exchange: the face kernels store into the neighbours' ghost planes
            for r, s in enumerate(slabs):
                s.step_faces(left=slabs[(r - 1) % world], right=slabs[(r + 1) % world])
        for s in slabs:
            s.step_interior()
            s.advance()
        if t in (0, 2, 5):
            one.velsrc()
            _cmp("%s one slab t%d" % (case.name, t + 1), one.owned(), ref)
            for s in slabs:
                s.velsrc()
                _cmp("%s rank %d/%d t%d" % (case.name, s.rank, world, t + 1), s.owned(), ref, slice(s.x0 * MK, (s.x0 + s.cnt) * MK))
    ref.close()


@pytest.mark.parametrize("seed", range(0, 120, 3))
def test_random_case_host_mirror_scalars(seed):
    """luma_b200.Definitions (the host-side image of definitions.h the ABI is fed from) derives what the reference derives"""
    case = random_case(seed)
    d = defs_from_case(case)
    try:
        g = port.PortGrid(case)
    except RuntimeError:
        pytest.skip("the reference rejects this case (omega >= 2)")
    assert (d.L_N, d.L_M, d.L_K, d.L_NUM_VELS) == (case.N, case.M, case.K, case.Q)
    assert d.omega == g.omega and d.nu == g.nu and d.gravity == g.gravity and d.rho_out == g.rho_out
    for a, b in zip(d.inlet_profiles(), (g.uin(0), g.uin(1), g.uin(2))):
        assert np.array_equal(a, b)
    lt = g.lattyp
    wall = g.wall.reshape(-1, 5)
    desc = d.boundary_site_descriptors(lt)
    assert len(desc) == int(np.isin(lt, (6, 7, 8)).sum())
    for site, ec, nd, n in desc:
        assert (ec, nd, *n) == tuple(int(v) for v in wall[site]), (case.name, site)
    g.close()
