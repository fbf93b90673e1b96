"""TEST INFRASTRUCTURE ONLY -- ctypes binding of the C restatement (oracle/libluma_oracle.so)
and a runner for the compiled reference binaries (oracle/_ref/luma_ref_<case>).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package (luma_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

try:  # imported as oracle.port (package) or with oracle/ on sys.path
    from .cases import Case, CASES, BENCH_CASES
except ImportError:  # pragma: no cover
    from cases import Case, CASES, BENCH_CASES

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libluma_oracle.so")
REF_DIR = os.path.join(HERE, "_ref")


class OracleCase(C.Structure):
    _fields_ = [
        ("dims", C.c_int32), ("resolution", C.c_int32),
        ("N", C.c_int32), ("M", C.c_int32), ("K", C.c_int32),
        ("dt", C.c_double),
        ("wall_type", C.c_int32 * 6), ("wall_thick", C.c_double * 6),
        ("u0", C.c_double * 3), ("rhoin", C.c_double),
        ("use_nu", C.c_int32), ("nu", C.c_double), ("re", C.c_double),
        ("regularised", C.c_int32), ("no_flow", C.c_int32),
        ("bgksmag", C.c_int32), ("csmag", C.c_double),
        ("gravity_on", C.c_int32), ("gravity_force", C.c_double), ("gravity_dir", C.c_int32),
        ("velocity_ramp_on", C.c_int32), ("velocity_ramp", C.c_double),
        ("reynolds_ramp_on", C.c_int32), ("reynolds_ramp", C.c_double),
        ("parabolic_inlet", C.c_int32), ("pressure_delta", C.c_double),
        ("ld_out", C.c_int32), ("has_box", C.c_int32), ("box", C.c_int32 * 6),
        ("time_averaged", C.c_int32), ("kbc", C.c_int32),
    ]


def case_struct(case: Case, struct_type=OracleCase):
    """Fill the run-time image of the definitions.h macros from a Case (same values the generated
    header gives the compiled reference)."""
    s = struct_type()
    s.dims, s.resolution = case.dims, case.resolution
    s.N, s.M, s.K = case.N, case.M, case.K
    s.dt = case.dt
    dh = 1.0 / float(case.resolution)          # L_COARSE_SITE_WIDTH
    for a in range(6):
        s.wall_type[a] = case.walls[a]
        s.wall_thick[a] = 0.0 if case.thick[a] == 0 else float(case.thick[a]) * dh
    s.u0[0], s.u0[1], s.u0[2] = case.ux0, case.uy0, (case.uz0 if case.dims == 3 else 0.0)
    s.rhoin = 1.0
    s.use_nu = int(case.nu is not None)
    s.nu = case.nu if case.nu is not None else 0.0
    s.re = case.re if case.re is not None else 1.0
    s.regularised, s.no_flow = int(case.regularised), int(case.no_flow)
    s.bgksmag, s.csmag = int(case.bgksmag), case.csmag
    s.gravity_on, s.gravity_force, s.gravity_dir = int(case.gravity_on), case.gravity_force, case.gravity_dir
    s.velocity_ramp_on = int(case.velocity_ramp is not None)
    s.velocity_ramp = case.velocity_ramp or 0.0
    s.reynolds_ramp_on = int(case.reynolds_ramp is not None)
    s.reynolds_ramp = case.reynolds_ramp or 0.0
    s.parabolic_inlet = int(case.parabolic_inlet)
    s.pressure_delta = case.pressure_delta
    s.ld_out = int(case.ld_out)
    s.has_box = int(case.box is not None)
    if case.box is not None:
        for a in range(6):
            s.box[a] = case.box[a]
    s.time_averaged = int(case.time_averaged)
    if hasattr(s, "kbc"):
        s.kbc = int(case.kbc)
    return s


_lib = None


def build_port() -> str:
    subprocess.run(["make", "-C", HERE, "port"], check=True, stdout=subprocess.DEVNULL)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(HERE, "luma_oracle.c")
        if (not os.path.exists(LIB_PATH)) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
            build_port()
        L = C.CDLL(LIB_PATH)
        L.luma_oracle_create.restype = C.c_void_p
        L.luma_oracle_create.argtypes = [C.POINTER(OracleCase)]
        L.luma_oracle_destroy.argtypes = [C.c_void_p]
        L.luma_oracle_step.restype = C.c_int
        L.luma_oracle_step.argtypes = [C.c_void_p, C.c_int]
        for nm in ("f", "fnew", "rho", "u", "rho_timeav", "ui_timeav", "uiuj_timeav"):
            fn = getattr(L, "luma_oracle_" + nm)
            fn.restype = C.POINTER(C.c_double)
            fn.argtypes = [C.c_void_p]
        for nm in ("lattyp", "wall"):
            fn = getattr(L, "luma_oracle_" + nm)
            fn.restype = C.POINTER(C.c_int32)
            fn.argtypes = [C.c_void_p]
        for nm in ("uin", "pos"):
            fn = getattr(L, "luma_oracle_" + nm)
            fn.restype = C.POINTER(C.c_double)
            fn.argtypes = [C.c_void_p, C.c_int]
        for nm in ("omega", "nu", "gravity", "rho_out"):
            fn = getattr(L, "luma_oracle_" + nm)
            fn.restype = C.c_double
            fn.argtypes = [C.c_void_p]
        L.luma_oracle_t.restype = C.c_int
        L.luma_oracle_t.argtypes = [C.c_void_p]
        L.luma_oracle_force.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        L.luma_oracle_velocity_ramp.restype = C.c_double
        L.luma_oracle_velocity_ramp.argtypes = [C.POINTER(OracleCase), C.c_double]
        L.luma_oracle_reynolds_ramp.restype = C.c_double
        L.luma_oracle_reynolds_ramp.argtypes = [C.POINTER(OracleCase), C.c_double]
        _lib = L
    return _lib


class PortGrid:
    """The C restatement holding one case (mirrors GridObj for level 0)."""

    def __init__(self, case: Case):
        self.case = case
        self._L = lib()
        self._cs = case_struct(case)
        self._h = self._L.luma_oracle_create(C.byref(self._cs))
        if not self._h:
            raise RuntimeError("luma_oracle_create failed for case %s" % case.name)
        self.N, self.M, self.K, self.Q, self.D = case.N, case.M, case.K, case.Q, case.dims
        self.nsites = self.N * self.M * self.K

    def close(self):
        if self._h:
            self._L.luma_oracle_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def step(self, n=1):
        rc = self._L.luma_oracle_step(self._h, int(n))
        if rc:
            raise RuntimeError("oracle fatal condition %d" % rc)

    def _arr(self, fn, n, dtype, *extra):
        p = fn(self._h, *extra)
        return np.ctypeslib.as_array(p, shape=(n,)).astype(dtype, copy=True)

    @property
    def f(self):
        return self._arr(self._L.luma_oracle_f, self.nsites * self.Q, np.float64)

    @property
    def rho(self):
        return self._arr(self._L.luma_oracle_rho, self.nsites, np.float64)

    @property
    def u(self):
        return self._arr(self._L.luma_oracle_u, self.nsites * self.D, np.float64)

    # time-averaged statistics (zeros unless the case sets time_averaged)
    @property
    def rho_timeav(self):
        return self._arr(self._L.luma_oracle_rho_timeav, self.nsites, np.float64)

    @property
    def ui_timeav(self):
        return self._arr(self._L.luma_oracle_ui_timeav, self.nsites * self.D, np.float64)

    @property
    def uiuj_timeav(self):
        return self._arr(self._L.luma_oracle_uiuj_timeav, self.nsites * (3 * self.D - 3), np.float64)

    @property
    def lattyp(self):
        return self._arr(self._L.luma_oracle_lattyp, self.nsites, np.int32)

    @property
    def wall(self):
        return self._arr(self._L.luma_oracle_wall, self.nsites * 5, np.int32)

    def uin(self, d):
        return self._arr(self._L.luma_oracle_uin, self.M, np.float64, d)

    def pos(self, d):
        n = (self.N, self.M, self.K)[d]
        return self._arr(self._L.luma_oracle_pos, n, np.float64, d)

    @property
    def omega(self):
        return self._L.luma_oracle_omega(self._h)

    @property
    def nu(self):
        return self._L.luma_oracle_nu(self._h)

    @property
    def gravity(self):
        return self._L.luma_oracle_gravity(self._h)

    @property
    def rho_out(self):
        return self._L.luma_oracle_rho_out(self._h)

    @property
    def t(self):
        return self._L.luma_oracle_t(self._h)

    @property
    def force(self):
        F = (C.c_double * 3)()
        self._L.luma_oracle_force(self._h, F)
        return np.array(list(F))

    def velocity_ramp(self, t_dimless):
        return self._L.luma_oracle_velocity_ramp(C.byref(self._cs), float(t_dimless))

    def reynolds_ramp(self, t_dimless):
        return self._L.luma_oracle_reynolds_ramp(C.byref(self._cs), float(t_dimless))


# ---------------------------------------------------------------------------------------------
# compiled reference (oracle/_ref)
# ---------------------------------------------------------------------------------------------

def ref_binary(name: str, omp: bool = False):
    p = os.path.join(REF_DIR, "luma_ref_%s%s" % (name, "_omp" if omp else ""))
    return p if os.path.exists(p) else None


def _read_kv(path):
    out = {}
    with open(path) as fh:
        for line in fh:
            if "=" in line:
                k, v = line.strip().split("=", 1)
                out[k] = v
    return out


def run_ref_dump(name: str, steps, outdir=None, omp: bool = False, threads=None):
    """Run the compiled reference for `steps` (ascending) and return {tag: {array name: ndarray}}.
    omp=True uses the OpenMP build (bit-identical results: every site is updated from the old lattice only)."""
    exe = ref_binary(name, omp=omp)
    if exe is None:
        raise FileNotFoundError("oracle/_ref/luma_ref_%s%s not built (make -C oracle ref)" % (name, "_omp" if omp else ""))
    own = outdir is None
    if own:
        tmp = tempfile.TemporaryDirectory(prefix="luma_ref_")
        outdir = tmp.name
    subprocess.run([exe, "dump", outdir, ",".join(str(int(s)) for s in steps)], check=True,
                   env=dict(os.environ, OMP_NUM_THREADS=str(int(threads or os.cpu_count() or 1)) if omp else "1"))
    res = {}
    for tag in ["init"] + ["t%d" % s for s in steps]:
        d = {}
        for arr, dt in (("f", np.float64), ("rho", np.float64), ("u", np.float64)):
            d[arr] = np.fromfile(os.path.join(outdir, "%s.%s.f64" % (tag, arr)), dtype=dt)
        for arr in ("rho_timeav", "ui_timeav", "uiuj_timeav"):
            path = os.path.join(outdir, "%s.%s.f64" % (tag, arr))
            if os.path.exists(path):
                d[arr] = np.fromfile(path, dtype=np.float64)
        d["scalars"] = _read_kv(os.path.join(outdir, tag + ".scalars.txt"))
        if tag == "init":
            d["lattyp"] = np.fromfile(os.path.join(outdir, "init.lattyp.i32"), dtype=np.int32)
            d["wall"] = np.fromfile(os.path.join(outdir, "init.wall.i32"), dtype=np.int32)
            for nm in ("xpos", "ypos", "zpos", "ux_in", "uy_in", "uz_in"):
                d[nm] = np.fromfile(os.path.join(outdir, "init.%s.f64" % nm), dtype=np.float64)
            d["meta"] = _read_kv(os.path.join(outdir, "init.meta.txt"))
        res[tag] = d
    if own:
        tmp.cleanup()
    return res


def run_ref_bench(name: str, warmup: int, steps: int, threads=None):
    """Time the compiled reference's own LBM_multi_opt loop (OpenMP build); returns its JSON dict."""
    import json
    exe = ref_binary(name, omp=True)
    if exe is None:
        raise FileNotFoundError("oracle/_ref/luma_ref_%s_omp not built" % name)
    env = dict(os.environ)
    if threads:
        env["OMP_NUM_THREADS"] = str(int(threads))
    out = subprocess.run([exe, "bench", str(int(warmup)), str(int(steps))], check=True, env=env,
                         stdout=subprocess.PIPE).stdout.decode()
    return json.loads(out.strip().splitlines()[-1])
