"""Shared helpers for the test-suite (test infrastructure; may import oracle/)."""
import numpy as np

from luma_b200 import Definitions
from oracle.cases import Case


def defs_from_case(case: Case) -> Definitions:
    """The oracle's case table and the product's Definitions describe the same definitions.h."""
    return Definitions(
        L_DIMS=case.dims, L_RESOLUTION=case.resolution, L_TIMESTEP=case.dt,
        L_BX=case.bx, L_BY=case.by, L_BZ=case.bz,
        L_UX0=case.ux0, L_UY0=case.uy0, L_UZ0=case.uz0,
        L_RE=case.re, L_NU=case.nu,
        L_USE_BGKSMAG=case.bgksmag, L_USE_KBC_COLLISION=case.kbc, L_CSMAG=case.csmag,
        L_GRAVITY_ON=case.gravity_on, L_GRAVITY_FORCE=case.gravity_force, L_GRAVITY_DIRECTION=case.gravity_dir,
        L_NO_FLOW=case.no_flow, L_PARABOLIC_INLET=case.parabolic_inlet,
        L_WALL_LEFT=case.walls[0], L_WALL_RIGHT=case.walls[1], L_WALL_BOTTOM=case.walls[2],
        L_WALL_TOP=case.walls[3], L_WALL_FRONT=case.walls[4], L_WALL_BACK=case.walls[5],
        L_WALL_THICKNESS_CELLS=tuple(case.thick),
        L_REGULARISED_BOUNDARIES=case.regularised,
        L_VELOCITY_RAMP=case.velocity_ramp, L_REYNOLDS_RAMP=case.reynolds_ramp,
        L_PRESSURE_DELTA=case.pressure_delta,
        L_COMPUTE_TIME_AVERAGED_QUANTITIES=case.time_averaged,
        body_box=case.box,
    )


def first_diff(a, b):
    a = np.asarray(a); b = np.asarray(b)
    bad = np.flatnonzero(~((a == b) | (np.isnan(a) & np.isnan(b))))
    if bad.size == 0:
        return "equal"
    i = int(bad[0])
    return "%d/%d differ, first at %d: %r vs %r (max abs diff %.3e)" % (
        bad.size, a.size, i, a[i], b[i], float(np.nanmax(np.abs(a - b))))


def max_rel_err(a, b):
    """north_star tolerance metric: max |a-b| / max|b| (u passes through 0, so normalise by the field max)."""
    a = np.asarray(a); b = np.asarray(b)
    scale = float(np.max(np.abs(b)))
    return float(np.max(np.abs(a - b))) / scale if scale > 0 else float(np.max(np.abs(a - b)))
