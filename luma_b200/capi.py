"""ctypes binding of include/luma_b200.h -- the same C ABI a LUMA build links against.

This module is plumbing only: structures, prototypes, error mapping.  It never computes anything
and there is no fallback: if the CUDA library is missing or fails, the call raises.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

# status codes (include/luma_b200.h)
OK, EINVAL, ECUDA, ENCCL, ENOMEM, EUNSUPPORTED, ESTATE, EBC_NOT_WALL, EBC_PRESSURE_EDGE, EBC_OFFGRID = range(10)
# eType (inc/Enumerations.h:84-96)
E_SOLID, E_FLUID, E_REFINED, E_VELOCITY, E_PRESSURE, E_SLIP, E_EXTRAPOLATE_RIGHT = 0, 1, 2, 6, 7, 8, 9
F, RHO, U = 1, 2, 4


class LumaCaseParams(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("dims", C.c_int32), ("num_vels", C.c_int32),
        ("N", C.c_int32), ("M", C.c_int32), ("K", C.c_int32),
        ("rank", C.c_int32), ("nranks", C.c_int32),
        ("x_offset", C.c_int32), ("x_count", C.c_int32), ("device", C.c_int32),
        ("regularised", C.c_int32), ("bgksmag", C.c_int32), ("csmag", C.c_double),
        ("gravity_on", C.c_int32), ("gravity_dir", C.c_int32), ("gravity", C.c_double),
        ("rhoin", C.c_double), ("rho_out", C.c_double), ("dt", C.c_double), ("dh", C.c_double),
        ("omega", C.c_double),
        ("velocity_ramp_on", C.c_int32), ("velocity_ramp", C.c_double),
        ("reynolds_ramp_on", C.c_int32), ("reynolds_ramp", C.c_double), ("re", C.c_double),
        ("t", C.c_int32), ("time_averaged", C.c_int32), ("kbc", C.c_int32),
    ]


class LumaSiteBC(C.Structure):
    _fields_ = [("site", C.c_int64), ("edge_count", C.c_int8), ("normal_dir", C.c_int8),
                ("normal", C.c_int8 * 3), ("pad_", C.c_int8 * 3)]


# the same 16-byte record as a numpy dtype (arrays of descriptors are handed over without a Python loop)
import numpy as _np  # noqa: E402
SITE_BC_DTYPE = _np.dtype({"names": ["site", "edge_count", "normal_dir", "normal"],
                           "formats": ["<i8", "i1", "i1", ("i1", (3,))], "offsets": [0, 8, 9, 10], "itemsize": 16})


class LumaSyntheticCase(C.Structure):
    _fields_ = [("wall_type", C.c_int32 * 6), ("wall_cells", C.c_int32 * 6), ("u_in", C.c_double * 3),
                ("ux_in", C.POINTER(C.c_double)), ("uy_in", C.POINTER(C.c_double)), ("uz_in", C.POINTER(C.c_double)),
                ("no_flow", C.c_int32), ("has_box", C.c_int32), ("box", C.c_int32 * 6)]


class LumaStats(C.Structure):
    _fields_ = [("steps", C.c_int64), ("ms_last_call", C.c_double), ("ms_per_step", C.c_double),
                ("mlups_last_call", C.c_double), ("kernel_launches", C.c_int64),
                ("halo_bytes_per_step", C.c_int64), ("cells", C.c_int64),
                ("step_kernel_launches", C.c_int64), ("step_kernel_ms", C.c_double), ("step_kernel_cells", C.c_int64),
                ("graph_launches", C.c_int64)]


class LumaHaloMsg(C.Structure):
    _fields_ = [("is_send", C.c_int32), ("peer", C.c_int32), ("pop", C.c_int32), ("plane", C.c_int32)]


class LumaB200Error(RuntimeError):
    def __init__(self, code, text, detail=""):
        super().__init__("luma_b200 error %d: %s%s" % (code, text, (" [" + detail + "]") if detail else ""))
        self.code = code


_lib = None
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


def lib_path() -> str:
    return _build.LIB


def load(build_if_missing: bool = True):
    """Load libluma_b200.so (building it in-tree first when it is missing or stale and nvcc exists)."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    if build_if_missing and not _build.is_current():
        try:
            _build.build()
        except Exception:
            if not os.path.exists(path):
                raise
    if not os.path.exists(path):
        raise ImportError("libluma_b200.so is not built (python -m luma_b200.build); there is no CPU fallback")
    L = C.CDLL(path)
    H = C.c_void_p
    L.luma_b200_abi_version.restype = C.c_int
    L.luma_b200_strerror.restype = C.c_char_p
    L.luma_b200_strerror.argtypes = [C.c_int]
    L.luma_b200_last_error.restype = C.c_char_p
    L.luma_b200_last_error.argtypes = [H]
    L.luma_b200_default_params.restype = None
    L.luma_b200_default_params.argtypes = [C.POINTER(LumaCaseParams)]
    L.luma_b200_create.argtypes = [C.POINTER(H), C.POINTER(LumaCaseParams)]
    L.luma_b200_destroy.restype = None
    L.luma_b200_destroy.argtypes = [H]
    L.luma_b200_slab.argtypes = [C.c_int32, C.c_int32, C.c_int32, _ip, _ip]
    L.luma_b200_comm_unique_id.argtypes = [C.c_void_p]
    L.luma_b200_comm_init.argtypes = [H, C.c_void_p]
    L.luma_b200_p2p_export.argtypes = [H, C.c_void_p]
    L.luma_b200_p2p_attach.argtypes = [H, C.c_void_p, C.c_void_p]
    L.luma_b200_upload.argtypes = [H, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.POINTER(LumaSiteBC), C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]
    L.luma_b200_init_synthetic.argtypes = [H, C.POINTER(LumaSyntheticCase)]
    L.luma_b200_step.argtypes = [H, C.c_int32]
    L.luma_b200_flush.argtypes = [H]
    L.luma_b200_restart_write.argtypes = [H, C.c_char_p]
    L.luma_b200_restart_read.argtypes = [H, C.c_char_p]
    L.luma_b200_download.argtypes = [H, C.c_int32, C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p]
    L.luma_b200_download_lattyp.argtypes = [H, C.c_int32, C.c_void_p]
    L.luma_b200_download_async.argtypes = [H, C.c_int32, C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p]
    L.luma_b200_download_wait.argtypes = [H]
    L.luma_b200_download_timeav.argtypes = [H, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    L.luma_b200_upload_timeav.argtypes = [H, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    L.luma_b200_get_time.argtypes = [H, _ip, _dp, _dp]
    L.luma_b200_forces.argtypes = [H, _dp]
    L.luma_b200_stats.argtypes = [H, C.POINTER(LumaStats)]
    L.luma_b200_sync.argtypes = [H]
    L.luma_b200_set_profiling.argtypes = [H, C.c_int32]
    L.luma_b200_halo_plan.argtypes = [C.POINTER(LumaCaseParams), C.POINTER(LumaHaloMsg), C.c_int32, _ip]
    L.luma_b200_selftest_div_const.argtypes = [C.c_int32, C.c_int64, C.c_uint64, C.POINTER(C.c_int64)]
    for nm in ("restart_write", "restart_read", "flush", "create", "slab", "comm_unique_id", "comm_init", "p2p_export", "p2p_attach", "upload", "init_synthetic", "step", "download",
               "download_lattyp", "download_async", "download_wait", "download_timeav", "upload_timeav", "get_time", "forces", "stats", "sync", "set_profiling", "selftest_div_const", "halo_plan"):
        getattr(L, "luma_b200_" + nm).restype = C.c_int
    if L.luma_b200_abi_version() != 4:
        raise ImportError("libluma_b200.so ABI version mismatch")
    _lib = L
    return L


def check(rc: int, handle=None):
    if rc != OK:
        L = load()
        detail = L.luma_b200_last_error(handle).decode() if handle else ""
        raise LumaB200Error(rc, L.luma_b200_strerror(rc).decode(), detail)


def default_params() -> LumaCaseParams:
    p = LumaCaseParams()
    load().luma_b200_default_params(C.byref(p))
    return p


def slab(N: int, nranks: int, rank: int):
    """(x_offset, x_count) of `rank` under the reference's uniform decomposition."""
    off, cnt = C.c_int32(), C.c_int32()
    check(load().luma_b200_slab(N, nranks, rank, C.byref(off), C.byref(cnt)))
    return off.value, cnt.value
