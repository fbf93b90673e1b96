"""Host-side mirror of LUMA's level-0 GridObj (inc/GridObj.h) whose time step runs on the B200.

`GridObj.LBM_multi_opt()` has the meaning of the reference's member of the same name
(src/GridObj_ops_lbm_optimised.cpp:36-193): one complete level-0 time step including the halo
exchange.  State lives on the device between calls; `f`, `rho`, `u`, `LatTyp` give host copies in
the reference's AoS layout when the host needs to look (the IO points of src/main_lbm.cpp:449-561).
Everything goes through the C ABI of include/luma_b200.h -- the same calls INTEGRATION.md's C++
shim makes -- so these Python classes are the reference-facing API used by tests and bench.py.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

from . import capi
from .definitions import Definitions, eFluid, ePressure, eSolid, eVelocity


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class GridObj:
    """Level-0 grid (or this rank's x-slab of it) with its populations resident on one B200."""

    def __init__(self, defs: Definitions, *, rank: int = 0, nranks: int = 1, device: int = 0, t: int = 0,
                 unique_id: Optional[bytes] = None):
        self.defs = defs
        self._L = capi.load()
        p = capi.default_params()
        p.dims, p.num_vels = defs.L_DIMS, defs.L_NUM_VELS
        p.N, p.M, p.K = defs.L_N, defs.L_M, defs.L_K
        p.rank, p.nranks = rank, nranks
        p.x_offset, p.x_count = capi.slab(defs.L_N, nranks, rank)
        p.device = device
        p.regularised = int(defs.L_REGULARISED_BOUNDARIES)
        p.bgksmag, p.csmag = int(defs.L_USE_BGKSMAG), defs.L_CSMAG
        p.gravity_on, p.gravity_dir, p.gravity = int(defs.L_GRAVITY_ON), int(defs.L_GRAVITY_DIRECTION), defs.gravity
        p.rhoin, p.rho_out = defs.L_RHOIN, defs.rho_out
        p.dt, p.dh = defs.dt, defs.dh
        p.omega = defs.omega
        p.velocity_ramp_on = int(defs.L_VELOCITY_RAMP is not None)
        p.velocity_ramp = defs.L_VELOCITY_RAMP or 0.0
        p.reynolds_ramp_on = int(defs.L_REYNOLDS_RAMP is not None)
        p.reynolds_ramp = defs.L_REYNOLDS_RAMP or 0.0
        p.re = float(defs.L_RE) if defs.L_RE is not None else 1.0
        p.t = t
        p.time_averaged = int(defs.L_COMPUTE_TIME_AVERAGED_QUANTITIES)
        p.kbc = int(defs.L_USE_KBC_COLLISION)
        self.params = p
        self.rank, self.nranks = rank, nranks
        self.x_offset, self.x_count = p.x_offset, p.x_count
        self.N_lim, self.M_lim, self.K_lim = p.x_count, p.M, p.K      # local sizes, inc/GridObj.h:122-124 (no halo)
        self.Q, self.D = p.num_vels, p.dims
        self._h = C.c_void_p()
        rc = self._L.luma_b200_create(C.byref(self._h), C.byref(p))
        if rc:
            h = self._h
            try:
                capi.check(rc, h if h else None)
            finally:
                if h:
                    self._L.luma_b200_destroy(h)
                self._h = C.c_void_p()
        if nranks > 1:
            if unique_id is None or len(unique_id) != 128:
                raise ValueError("nranks > 1 needs the 128-byte NCCL unique id broadcast from rank 0")
            buf = C.create_string_buffer(unique_id, 128)
            capi.check(self._L.luma_b200_comm_init(self._h, buf), self._h)

    # ---- device-initiated halo exchange (NVLink peer stores instead of NCCL send/recv) ----
    def p2p_export(self) -> bytes:
        buf = C.create_string_buffer(256)
        capi.check(self._L.luma_b200_p2p_export(self._h, buf), self._h)
        return buf.raw

    def p2p_attach(self, left_blob: bytes, right_blob: bytes):
        capi.check(self._L.luma_b200_p2p_attach(self._h, C.create_string_buffer(left_blob, 256),
                                                C.create_string_buffer(right_blob, 256)), self._h)
        return self

    # ---- construction of the state ----
    def LBM_initGrid(self):
        """Device-side equivalent of GridObj::LBM_initGrid + body labelling for `defs`."""
        d = self.defs
        c = capi.LumaSyntheticCase()
        ux, uy, uz = d.inlet_profiles()
        self._keep = (np.ascontiguousarray(ux), np.ascontiguousarray(uy), np.ascontiguousarray(uz))
        for a in range(6):
            c.wall_type[a] = d.walls[a]
            c.wall_cells[a] = d.L_WALL_THICKNESS_CELLS[a]
        dp = C.POINTER(C.c_double)
        c.ux_in, c.uy_in, c.uz_in = (x.ctypes.data_as(dp) for x in self._keep)
        c.no_flow = int(d.L_NO_FLOW)
        c.has_box = int(d.body_box is not None)
        if d.body_box is not None:
            for a in range(6):
                c.box[a] = d.body_box[a]
        capi.check(self._L.luma_b200_init_synthetic(self._h, C.byref(c)), self._h)
        return self

    def upload(self, f, rho, u, LatTyp, ux_in=None, uy_in=None, uz_in=None, bc_sites=None, halo: int = 0):
        """Hand over the host state of an existing LUMA GridObj (AoS arrays covering this rank's planes,
        plus `halo` planes each side).  `bc_sites`: iterable of (site, edge_count, normal_dir, (nx,ny,nz));
        by default computed from the case's wall thicknesses like GridUtils::isWithinDomainWall."""
        # f=None: the host declares f = feq(rho,u) everywhere (LBM_initGrid's state at t = 0); nothing is uploaded for it
        f = None if f is None else np.ascontiguousarray(f, dtype=np.float64)
        rho = np.ascontiguousarray(rho, dtype=np.float64)
        u = np.ascontiguousarray(u, dtype=np.float64)
        lt = np.ascontiguousarray(LatTyp, dtype=np.int32)
        ncell = (self.x_count + 2 * halo) * self.M_lim * self.K_lim
        if (f is not None and f.size != ncell * self.Q) or rho.size != ncell or u.size != ncell * self.D or lt.size != ncell:
            raise ValueError("upload: array sizes do not match the local grid")
        if bc_sites is None:
            bc_sites = self.defs.boundary_site_descriptors(lt, x_offset=self.x_offset - halo)
        if not (isinstance(bc_sites, np.ndarray) and bc_sites.dtype == capi.SITE_BC_DTYPE):
            rec = np.zeros(len(bc_sites), dtype=capi.SITE_BC_DTYPE)
            for n, (site, ec, nd, nv) in enumerate(bc_sites):
                rec[n] = (site, ec, nd, nv)
            bc_sites = rec
        bc_sites = np.ascontiguousarray(bc_sites)
        arr = bc_sites.ctypes.data_as(C.POINTER(capi.LumaSiteBC)) if len(bc_sites) else None
        prof = [None if a is None else np.ascontiguousarray(a, dtype=np.float64) for a in (ux_in, uy_in, uz_in)]
        capi.check(self._L.luma_b200_upload(self._h, halo, _ptr(f), _ptr(rho), _ptr(u), _ptr(lt), arr, len(bc_sites),
                                            _ptr(prof[0]), _ptr(prof[1]), _ptr(prof[2])), self._h)
        return self

    # ---- the time step ----
    def LBM_multi_opt(self, nsteps: int = 1):
        """Accepts `nsteps` time steps; never waits for the GPU (see luma_b200_step in include/luma_b200.h)."""
        capi.check(self._L.luma_b200_step(self._h, int(nsteps)), self._h)

    def flush(self):
        """Submit every accepted step to the GPU without reading anything back or waiting."""
        capi.check(self._L.luma_b200_flush(self._h), self._h)

    # ---- scalars the reference keeps on the object ----
    def _time(self):
        t, om, nu = C.c_int32(), C.c_double(), C.c_double()
        capi.check(self._L.luma_b200_get_time(self._h, C.byref(t), C.byref(om), C.byref(nu)), self._h)
        return t.value, om.value, nu.value

    @property
    def t(self):
        return self._time()[0]

    @property
    def omega(self):
        return self._time()[1]

    @property
    def nu(self):
        return self._time()[2]

    # ---- host views (AoS, reference layout), owned planes only ----
    def download(self, what=capi.F | capi.RHO | capi.U, out=None):
        n = self.x_count * self.M_lim * self.K_lim
        out = out or {}
        f = out.get("f") if what & capi.F else None
        rho = out.get("rho") if what & capi.RHO else None
        u = out.get("u") if what & capi.U else None
        if what & capi.F and f is None:
            f = np.empty(n * self.Q)
        if what & capi.RHO and rho is None:
            rho = np.empty(n)
        if what & capi.U and u is None:
            u = np.empty(n * self.D)
        capi.check(self._L.luma_b200_download(self._h, 0, what, _ptr(f), _ptr(rho), _ptr(u)), self._h)
        return {"f": f, "rho": rho, "u": u}

    def download_async(self, what, out):
        """Start a download of this step's fields into the (pinned) arrays of `out` while later steps run;
        read them after download_wait().  src/main_lbm.cpp:449-561 with the IO off the critical path."""
        capi.check(self._L.luma_b200_download_async(self._h, 0, what, _ptr(out.get("f")) if what & capi.F else None,
                                                    _ptr(out.get("rho")) if what & capi.RHO else None,
                                                    _ptr(out.get("u")) if what & capi.U else None), self._h)

    def download_wait(self):
        capi.check(self._L.luma_b200_download_wait(self._h), self._h)

    @property
    def f(self):
        return self.download(capi.F)["f"]

    @property
    def rho(self):
        return self.download(capi.RHO)["rho"]

    @property
    def u(self):
        return self.download(capi.U)["u"]

    @property
    def LatTyp(self):
        lt = np.empty(self.x_count * self.M_lim * self.K_lim, dtype=np.int32)
        capi.check(self._L.luma_b200_download_lattyp(self._h, 0, _ptr(lt)), self._h)
        return lt

    def download_timeav(self):
        """rho_timeav, ui_timeav, uiuj_timeav (inc/GridObj.h:93-95) of the owned planes."""
        n = self.x_count * self.M_lim * self.K_lim
        out = {"rho_timeav": np.empty(n), "ui_timeav": np.empty(n * self.D), "uiuj_timeav": np.empty(n * (3 * self.D - 3))}
        capi.check(self._L.luma_b200_download_timeav(self._h, 0, _ptr(out["rho_timeav"]), _ptr(out["ui_timeav"]),
                                                     _ptr(out["uiuj_timeav"])), self._h)
        return out

    def upload_timeav(self, rho_timeav=None, ui_timeav=None, uiuj_timeav=None):
        arrs = [None if a is None else np.ascontiguousarray(a, dtype=np.float64) for a in (rho_timeav, ui_timeav, uiuj_timeav)]
        capi.check(self._L.luma_b200_upload_timeav(self._h, 0, _ptr(arrs[0]), _ptr(arrs[1]), _ptr(arrs[2])), self._h)

    def io_restart_write(self, path: str):
        """Binary restart file of this rank's state (GridObj::io_restart(eWrite), src/GridObj_ops_io.cpp:406)."""
        capi.check(self._L.luma_b200_restart_write(self._h, os.fsencode(path)), self._h)

    def io_restart_read(self, path: str):
        """Replace t, f, rho, u (and the time averages) by a restart file's (GridObj::io_restart(eRead), :519)."""
        capi.check(self._L.luma_b200_restart_read(self._h, os.fsencode(path)), self._h)

    def computeLiftDrag(self):
        """Momentum-exchange force on bounce-back bodies of the last step (this rank's share)."""
        F = (C.c_double * 3)()
        capi.check(self._L.luma_b200_forces(self._h, F), self._h)
        return np.array([F[0], F[1], F[2]])

    def stats(self) -> dict:
        s = capi.LumaStats()
        capi.check(self._L.luma_b200_stats(self._h, C.byref(s)), self._h)
        return {k: getattr(s, k) for k, _ in s._fields_}

    def set_profiling(self, on: bool = True):
        capi.check(self._L.luma_b200_set_profiling(self._h, int(on)), self._h)

    def sync(self):
        capi.check(self._L.luma_b200_sync(self._h), self._h)

    def close(self):
        if getattr(self, "_h", None):
            self._L.luma_b200_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def comm_unique_id() -> bytes:
    """128 bytes of a fresh ncclUniqueId (rank 0 creates it, the host broadcasts it)."""
    buf = C.create_string_buffer(128)
    capi.check(capi.load().luma_b200_comm_unique_id(buf))
    return buf.raw
