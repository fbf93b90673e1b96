"""Host-side mirror of LUMA's compile-time case description (inc/definitions.h) and of the scalars
and small tables the host derives from it before the first time step.

LUMA fixes a case by editing macros in inc/definitions.h; `Definitions` carries the same macros as
attributes with the same names, and derives what GridObj::LBM_initGrid
(src/GridObj_init_grids.cpp:155-384), GridUnits (inc/GridUnits.h) and GridUtils::isWithinDomainWall
(src/GridUtils.cpp:1369-1430) derive, with the same double-precision expressions in the same order,
so that the numbers handed to the C ABI are the ones an unmodified LUMA host would hand over.
Only scalars, O(M) profiles and per-boundary-site descriptors are computed here -- no lattice
arithmetic (that lives in the CUDA library; there is no CPU path).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Optional, Tuple

import numpy as np

# eType, inc/Enumerations.h:84-96
eSolid, eFluid, eRefined, eVelocity, ePressure, eSlip, eExtrapolateRight = 0, 1, 2, 6, 7, 8, 9
# eCartesianDirection
eXDirection, eYDirection, eZDirection = 0, 1, 2

L_PI = 3.14159265358979323846      # inc/stdafx.h:114


def linspace(lo: float, hi: float, n: int) -> np.ndarray:
    """GridUtils::linspace (inc/GridUtils.h:289-309): lo + spacing*i, spacing = (hi-lo)/(n-1)."""
    n = max(int(n), 2)
    spacing = (hi - lo) / float(n - 1)
    return np.array([lo + spacing * i for i in range(n)], dtype=np.float64)


@dataclass
class Definitions:
    # --- lattice / domain (definitions.h:170-177) ---
    L_DIMS: int = 3
    L_RESOLUTION: int = 64
    L_TIMESTEP: float = 0.05 / 64.0
    L_BX: float = 1.0
    L_BY: float = 1.0
    L_BZ: float = 1.0
    # --- fluid (definitions.h:196-204) ---
    L_UX0: float = 1.0
    L_UY0: float = 0.0
    L_UZ0: float = 0.0
    L_RHOIN: float = 1.0
    L_PHYSICAL_RHO: float = 1000.0
    L_RE: Optional[float] = 100.0
    L_NU: Optional[float] = None              # if set, overrides L_RE (init_grids.cpp:336-340)
    # --- models (definitions.h:112-128) ---
    L_USE_BGKSMAG: bool = False
    L_USE_KBC_COLLISION: bool = False         # KBC-D on D2Q9 / KBC-N4 on D3Q27 instead of LBGK (definitions.h:125)
    L_CSMAG: float = 0.3
    L_GRAVITY_ON: bool = False
    L_GRAVITY_FORCE: float = 0.0
    L_GRAVITY_DIRECTION: int = eXDirection
    L_NO_FLOW: bool = True
    L_PARABOLIC_INLET: bool = False
    # --- boundaries (definitions.h:230-253) ---
    L_WALL_LEFT: int = eSolid
    L_WALL_RIGHT: int = eSolid
    L_WALL_BOTTOM: int = eSolid
    L_WALL_TOP: int = eSolid
    L_WALL_FRONT: int = eSolid
    L_WALL_BACK: int = eSolid
    # thickness in coarse cells; the macros are (n * L_COARSE_SITE_WIDTH), 0 -> 0.0
    L_WALL_THICKNESS_CELLS: Tuple[int, int, int, int, int, int] = (1, 1, 1, 1, 1, 1)   # left,right,bottom,top,front,back
    L_REGULARISED_BOUNDARIES: bool = True
    L_VELOCITY_RAMP: Optional[float] = None
    L_REYNOLDS_RAMP: Optional[float] = None
    L_PRESSURE_DELTA: float = 0.0
    # --- output options (definitions.h:130) ---
    L_COMPUTE_TIME_AVERAGED_QUANTITIES: bool = False
    # --- bounce-back body given as an index box (stands in for input/geometry.config + point cloud) ---
    body_box: Optional[Tuple[int, int, int, int, int, int]] = None

    # ---- derived exactly as the reference derives them ----
    @property
    def L_N(self) -> int:
        return int(self.L_BX * self.L_RESOLUTION)          # definitions.h:46

    @property
    def L_M(self) -> int:
        return int(self.L_BY * self.L_RESOLUTION)

    @property
    def L_K(self) -> int:
        return int(self.L_BZ * self.L_RESOLUTION) if self.L_DIMS == 3 else 1   # :48, :320-321

    @property
    def L_NUM_VELS(self) -> int:
        if self.L_DIMS == 3:
            return 27 if self.L_USE_KBC_COLLISION else 19  # definitions.h:299-310
        return 9

    @property
    def dh(self) -> float:
        return 1.0 / float(self.L_RESOLUTION)              # L_COARSE_SITE_WIDTH, definitions.h:50

    @property
    def dt(self) -> float:
        return float(self.L_TIMESTEP)

    @property
    def cs(self) -> float:
        return 1.0 / math.sqrt(3.0)                        # src/stdafx.cpp:153

    @property
    def walls(self):
        return (self.L_WALL_LEFT, self.L_WALL_RIGHT, self.L_WALL_BOTTOM, self.L_WALL_TOP, self.L_WALL_FRONT, self.L_WALL_BACK)

    @property
    def wall_thickness(self):
        """L_WALL_THICKNESS_* (dimensionless), same order as `walls`."""
        return tuple(0.0 if n == 0 else float(n) * self.dh for n in self.L_WALL_THICKNESS_CELLS)

    @property
    def nu(self) -> float:
        """GridUnits::nud2nulbm (inc/GridUnits.h:128) of L_NU or 1/L_RE (init_grids.cpp:336-340)."""
        nud = float(self.L_NU) if self.L_NU is not None else 1.0 / float(self.L_RE)
        return (nud * self.dt) / (self.dh * self.dh)

    @property
    def omega(self) -> float:
        cs = self.cs
        return 1.0 / ((self.nu / (cs * cs)) + 0.5)         # init_grids.cpp:344

    @property
    def gravity(self) -> float:
        return (self.L_GRAVITY_FORCE * (self.dt * self.dt)) / self.dh      # fd2flbm, GridUnits.h:140

    @property
    def rho_out(self) -> float:
        """L_RHOIN + pd2dlbm(L_PRESSURE_DELTA) (optimised.cpp:343-345, GridUnits.h:178)."""
        cs = self.cs
        dm = (self.L_PHYSICAL_RHO / self.L_RHOIN) * self.dh * self.dh * self.dh   # init_grids.cpp:182
        return self.L_RHOIN + (self.L_PRESSURE_DELTA * self.dh * (self.dt * self.dt) / dm) / (cs * cs)

    def positions(self):
        """XPos, YPos, ZPos: cell centres (init_grids.cpp:207-240)."""
        dh = self.dh
        Lx, Ly = dh * self.L_N, dh * self.L_M
        x = linspace(dh / 2.0, Lx - dh / 2.0, self.L_N)
        y = linspace(dh / 2.0, Ly - dh / 2.0, self.L_M)
        if self.L_DIMS == 3:
            z = linspace(dh / 2.0, dh * self.L_K - dh / 2.0, self.L_K)
        else:
            z = np.zeros(2)
        return x, y, z

    def inlet_profiles(self):
        """ux_in, uy_in, uz_in [M] in lattice units (_LBM_initSetInletProfile, init_grids.cpp:1322-1360)."""
        M = self.L_M
        dt, dh = self.dt, self.dh
        if self.L_PARABOLIC_INLET:
            th = self.wall_thickness
            y = self.positions()[1]
            b = dh * M - th[3]
            p = (b + th[2]) / 2.0
            q = b - p
            ux = np.array([((1.5 * self.L_UX0 * dt) / dh) * (1.0 - math.pow((y[j] - p) / q, 2.0)) for j in range(M)])
            return ux, np.zeros(M), np.zeros(M)
        uz0 = self.L_UZ0 if self.L_DIMS == 3 else 0.0
        mk = lambda v: np.full(M, (v * dt) / dh, dtype=np.float64)     # ud2ulbm, GridUnits.h:68
        return mk(self.L_UX0), mk(self.L_UY0), mk(uz0)

    def velocity_ramp_coefficient(self, t: float) -> float:
        """GridUtils::getVelocityRampCoefficient (src/GridUtils.cpp:1808-1816)."""
        if self.L_VELOCITY_RAMP is not None and t <= self.L_VELOCITY_RAMP:
            return (1.0 - math.cos(L_PI * t / self.L_VELOCITY_RAMP)) / 2.0
        return 1.0

    def wall_descriptor(self, i: int, j: int, k: int):
        """GridUtils::isWithinDomainWall for cell (i,j,k): (edgeCount, normalDirection, (nx,ny,nz))."""
        x, y, z = self.positions()
        return self._wall(x[i], y[j], z[k] if self.L_DIMS == 3 else 0.0)

    def _wall(self, x, y, z):
        th = self.wall_thickness
        dh = self.dh
        Lx, Ly, Lz = dh * self.L_N, dh * self.L_M, dh * self.L_K
        n = [0, 0, 0]
        nd, ec = 3, 0
        if x > 0.0 and x < th[0]:
            nd, n[0], ec = 0, 1, ec + 1
        if x < Lx and x > Lx - th[1]:
            nd, n[0], ec = 0, -1, ec + 1
        if y > 0.0 and y < th[2]:
            nd, n[1], ec = 1, 1, ec + 1
        if y < Ly and y > Ly - th[3]:
            nd, n[1], ec = 1, -1, ec + 1
        if self.L_DIMS == 3:
            if z > 0.0 and z < th[4]:
                nd, n[2], ec = 2, 1, ec + 1
            if z < Lz and z > Lz - th[5]:
                nd, n[2], ec = 2, -1, ec + 1
        return ec, nd, tuple(n)

    def boundary_site_descriptors(self, lattyp: np.ndarray, x_offset: int = 0) -> np.ndarray:
        """Descriptors for every eVelocity/ePressure/eSlip site of a (slab of a) LatTyp array laid out
        k + K*(j + M*i), as GridUtils::isWithinDomainWall gives them: a record array with the layout of
        LumaSiteBC (site, edge_count, normal_dir, normal[3]); iterating it yields those four per site."""
        from .capi import SITE_BC_DTYPE
        M, K = self.L_M, self.L_K
        lattyp = np.asarray(lattyp).ravel()
        sites = np.flatnonzero((lattyp == eVelocity) | (lattyp == ePressure) | (lattyp == eSlip)).astype(np.int64)
        x, y, z = self.positions()
        i, rem = np.divmod(sites, M * K)
        j, k = np.divmod(rem, K)
        px, py = x[(i + x_offset) % self.L_N], y[j]
        pz = z[k] if self.L_DIMS == 3 else np.zeros(sites.size)
        th = self.wall_thickness
        dh = self.dh
        Lx, Ly, Lz = dh * self.L_N, dh * self.L_M, dh * self.L_K
        n = np.zeros((sites.size, 3), dtype=np.int8)
        nd = np.zeros(sites.size, dtype=np.int8)
        ec = np.zeros(sites.size, dtype=np.int8)
        tests = [(0, 1, (px > 0.0) & (px < th[0])), (0, -1, (px < Lx) & (px > Lx - th[1])),
                 (1, 1, (py > 0.0) & (py < th[2])), (1, -1, (py < Ly) & (py > Ly - th[3]))]
        if self.L_DIMS == 3:
            tests += [(2, 1, (pz > 0.0) & (pz < th[4])), (2, -1, (pz < Lz) & (pz > Lz - th[5]))]
        for d, sign, hit in tests:            # same order as the reference: the last wall hit names normal_dir
            n[hit, d] = sign
            nd[hit] = d
            ec[hit] += 1
        out = np.zeros(sites.size, dtype=SITE_BC_DTYPE)
        out["site"], out["edge_count"], out["normal_dir"], out["normal"] = sites, ec, nd, n
        return out
