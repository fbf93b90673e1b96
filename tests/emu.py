"""TEST INFRASTRUCTURE ONLY -- driver of tests/harness/kernels_host.cpp: the product's kernel bodies
(luma_b200/csrc/kernels_impl.cuh) compiled for the host and run one "thread" at a time, so that the device code's
logic can be replayed against the oracle where no GPU is at hand.  Not a CPU implementation of the product: no C
ABI, no streams, no transport; nothing under luma_b200/ can reach it.

`Slab` mirrors what one handle of luma_b200/csrc/api.cu holds for one rank: SoA lattices (with ghost planes when
there is more than one slab), eType, wall descriptors, cell words and the boundary-site lists, built the way
finalize_geometry() builds them.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "harness", "kernels_host.cpp")
LIB = os.path.join(HERE, "harness", "libkernels_host.so")
CSRC = os.path.join(os.path.dirname(HERE), "luma_b200", "csrc")
CUDA_INC = "/usr/local/cuda/include"


class EmuCase(C.Structure):
    _fields_ = [("Q", C.c_int32), ("D", C.c_int32), ("P", C.c_int32), ("M", C.c_int32), ("K", C.c_int32),
                ("regularised", C.c_int32), ("coll", C.c_int32), ("force", C.c_int32), ("gravity_dir", C.c_int32),
                ("velramp_on", C.c_int32), ("general", C.c_int32),
                ("omega", C.c_double), ("rhoin", C.c_double), ("rho_out", C.c_double), ("gravity", C.c_double), ("csmag", C.c_double),
                ("ramp", C.c_double), ("ramp_t", C.c_double), ("t_now", C.c_double), ("t_next", C.c_double),
                ("wrap_x", C.c_int32), ("p0", C.c_int32), ("pstep", C.c_int32), ("nplanes", C.c_int32), ("run_bc", C.c_int32),
                ("N", C.c_int32), ("x_first", C.c_int32),
                ("faces", C.c_int32), ("peer_P", C.c_int32 * 2), ("peer_stride", C.c_longlong * 2), ("peer_f", C.c_void_p * 2)]


_lib = None


def available():
    return os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h"))


def load():
    global _lib
    if _lib is None:
        deps = [SRC] + [os.path.join(CSRC, n) for n in ("kernels_impl.cuh", "kernels.cuh", "lattice.cuh")]
        if (not os.path.exists(LIB)) or os.path.getmtime(LIB) < max(os.path.getmtime(d) for d in deps):
            cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
            subprocess.run([cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-w",
                            "-I", CUDA_INC, "-o", LIB, SRC], check=True)
        L = C.CDLL(LIB)
        for nm in ("emu_cell_words", "emu_step", "emu_velsrc", "emu_synthetic", "emu_lattice_c", "emu_class_shift"):
            getattr(L, nm).restype = C.c_int
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def pack_desc(ec, nd, n):
    """cw_pack_bc() of lattice.cuh"""
    return ((ec & 3) << 30) | ((nd & 3) << 22) | (((n[0] + 1) | ((n[1] + 1) << 2) | ((n[2] + 1) << 4)) << 24)


def slab_of(N, world, rank):
    """luma_b200_slab (MpiManager::mpi_uniformDecompose)"""
    per = -(-N // world)
    last = per - (per * world - N)
    if last <= 0:
        per = N // world
        last = per - (per * world - N)
    return per * rank, (last if rank == world - 1 else per)


class Rejected(Exception):
    """a case luma_b200_upload / init_synthetic refuses (api.cu finalize_geometry): what the reference L_ERRORs on, or
    what is loop-order dependent in the reference"""


class Slab:
    def __init__(self, case, ref, rank=0, world=1):
        self.L = load()
        self.case, self.rank, self.world = case, rank, world
        N, M, K, Q, D = case.N, case.M, case.K, case.Q, case.dims
        self.N, self.M, self.K, self.Q, self.D, self.MK = N, M, K, Q, D, M * K
        self.ghost = 1 if world > 1 else 0
        self.x0, self.cnt = slab_of(N, world, rank)
        self.P = self.cnt + 2 * self.ghost
        self.cells = self.P * self.MK
        self.stride = (self.cells + 15) // 16 * 16
        self.c = [[self.L.emu_lattice_c(Q, v, d) for d in range(3)] for v in range(Q)]
        # global x of every local plane (ring)
        self.gx = [(self.x0 - self.ghost + p) % N for p in range(self.P)]
        p = EmuCase()
        p.Q, p.D, p.P, p.M, p.K = Q, D, self.P, M, K
        p.regularised = int(case.regularised)
        p.coll = 2 if case.kbc else (1 if case.bgksmag else 0)
        p.force, p.gravity_dir = int(case.gravity_on), case.gravity_dir
        p.velramp_on = int(case.velocity_ramp is not None)
        p.rhoin, p.rho_out, p.gravity, p.csmag = 1.0, ref.rho_out, ref.gravity, case.csmag
        p.wrap_x = 0 if self.ghost else 1
        p.N, p.x_first = N, self.x0 - self.ghost
        self.p = p
        self.f = [np.full((Q, self.stride), np.nan), np.full((Q, self.stride), np.nan)]
        self.rho = np.zeros(self.stride)
        self.u = np.zeros((D, self.stride))
        self.types = np.zeros(self.cells, dtype=np.uint8)
        self.desc = np.zeros(self.cells, dtype=np.uint32)
        self.cw = np.zeros(self.cells, dtype=np.uint32)
        self.uin = np.ascontiguousarray(np.stack([ref.uin(0), ref.uin(1), ref.uin(2)]))
        self.tav = np.zeros((1 + D + 3 * D - 3, self.stride)) if case.time_averaged else None
        self.cur, self.t = 0, 0

    # ---- state ----
    def upload_from(self, ref):
        """what luma_b200_upload + exchange_ghost_planes leave on the device: owned planes from the host arrays, the
        ghost planes' eType / descriptors / rho / u from the ring neighbours (their populations come with exchange())"""
        N, MK, Q, D = self.N, self.MK, self.Q, self.D
        f = ref.f.reshape(N, MK, Q)
        rho = ref.rho.reshape(N, MK)
        u = ref.u.reshape(N, MK, D)
        lt = ref.lattyp.reshape(N, MK).astype(np.uint8)
        wall = ref.wall.reshape(N, MK, 5)
        for pl, gx in enumerate(self.gx):
            sl = slice(pl * MK, (pl + 1) * MK)
            owned = self.ghost <= pl < self.P - self.ghost
            if owned:
                self.f[0][:, sl] = f[gx].T
            self.rho[sl] = rho[gx]
            self.u[:, sl] = u[gx].T
            self.types[sl] = lt[gx]
            for s in np.flatnonzero(np.isin(lt[gx], (6, 7, 8)) & (wall[gx, :, 0] > 0)):
                self.desc[pl * MK + s] = pack_desc(int(wall[gx, s, 0]), int(wall[gx, s, 1]), [int(x) for x in wall[gx, s, 2:5]])
        self.f[1][:] = self.f[0]
        return self

    def init_synthetic(self, ref):
        """luma_b200_init_synthetic: k_synthetic on all local planes"""
        case = self.case
        wt = np.array(case.walls, dtype=np.int32)
        wc = np.array(case.thick, dtype=np.int32)
        box = np.array(case.box if case.box is not None else (0,) * 6, dtype=np.int32)
        self.f[0][:] = 0.0
        self.f[1][:] = 0.0
        rc = self.L.emu_synthetic(C.byref(self.p), _ptr(wt), _ptr(wc), _ptr(self.uin), C.c_double(ref.velocity_ramp(0.0)),
                                  C.c_int(int(case.no_flow)), C.c_int(int(case.box is not None)), _ptr(box), _ptr(self.types), _ptr(self.desc),
                                  _ptr(self.f[0]), _ptr(self.f[1]), _ptr(self.rho), _ptr(self.u), C.c_longlong(self.stride))
        assert rc == 0
        return self

    # ---- finalize_geometry() of api.cu ----
    def finalize(self):
        case = self.case
        P, M, K, Q, D = self.P, self.M, self.K, self.Q, self.D
        pb, pe = self.ghost, P - self.ghost
        wrap = self.ghost == 0
        reg = case.regularised
        c = self.c
        types = self.types.reshape(P, M, K)
        desc = self.desc.reshape(P, M, K)
        sid = lambda i, j, k: (i * M + j) * K + k
        never = lambda t: t in (0, 2) or (t == 6 and not reg)
        lst, forced, extra, vel, general = [], set(), {}, [], False
        for i, j, k in np.argwhere(types > 1):
            i, j, k = int(i), int(j), int(k)
            t = int(types[i, j, k])
            me = sid(i, j, k)
            owned = pb <= i < pe
            if t == 9 or (t == 6 and not reg):
                general = True
                for v in range(Q):
                    di, dj, dk = i + c[v][0], (j + c[v][1]) % M, (k + c[v][2]) % K
                    if wrap:
                        di %= P
                    if di < pb or di >= pe:
                        continue
                    dt = int(types[di, dj, dk])
                    if never(dt):
                        continue
                    if t == 9 and i - 2 < pb:
                        raise Rejected("eExtrapolateRight site within two planes of the low x end of the slab")
                    if dt == 1:
                        forced.add(sid(di, dj, dk))
            if not owned:
                continue
            if t == 8:
                general = True
                if (int(desc[i, j, k]) >> 30) == 0:
                    raise Rejected("slip site outside every wall region")
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
            if ec == 0:
                raise Rejected("regularised BC site not within a wall")
            if ec > 1 and t == 7:
                raise Rejected("pressure BC on an edge or corner")
            if ec > 1 or t == 7:
                n = [((d >> (24 + 2 * a)) & 3) - 1 for a in range(3)]
                ncalls = (D - 1) if t == 7 else 1
                for m in (1, 2):
                    pp, jj, kk = i + m * n[0], j + m * n[1], k + m * n[2]
                    gi = self.x0 + (i - self.ghost) + m * n[0]
                    if gi < 0 or gi >= self.N or jj < 0 or jj >= M or kk < 0 or kk >= K:
                        raise Rejected("extrapolation site off grid")
                    if pp < pb or pp >= pe:
                        raise Rejected("slab too thin: a boundary site extrapolates from a plane owned by another rank")
                    idn = sid(pp, jj, kk)
                    tn = int(types[pp, jj, kk])
                    if tn not in (0, 1):
                        raise Rejected("a boundary site extrapolates from another boundary site")
                    if idn > me:
                        if tn == 0:
                            raise Rejected("a boundary site extrapolates from an eSolid site with a larger index")
                        if case.time_averaged:
                            extra[idn] = extra.get(idn, 0) + ncalls
                            forced.add(idn)
            lst.append(me)
        full = sorted(lst + sorted(forced))
        self.bc_extra = None
        if extra:
            self.bc_extra = np.zeros(len(full), dtype=np.int32)
            pos = {s: a for a, s in enumerate(full)}
            for s, e in extra.items():
                self.bc_extra[pos[s]] += e
        self.bc_list = np.array(full, dtype=np.int64)
        self.vel = np.array(vel, dtype=np.int64)
        self.p.general = int(general or bool(forced))
        assert self.L.emu_cell_words(C.byref(self.p), _ptr(self.types), _ptr(self.desc), _ptr(self.cw)) == 0
        shift = self.L.emu_class_shift(Q)
        for s in sorted(forced):                                   # k_force_general
            if (int(self.cw[s]) >> shift) & 7 == 1:
                self.cw[s] = (int(self.cw[s]) & ~(7 << shift)) | (4 << shift)
        cls = (self.cw >> np.uint32(shift)) & np.uint32(7)
        assert np.array_equal(np.flatnonzero(cls >= 2), self.bc_list)
        return self

    # ---- one time step (the scalars are the ones luma_b200_step derives; omega is handed in) ----
    def set_scalars(self, ref, omega):
        t, dt = self.t, self.case.dt
        self.p.omega = omega
        self.p.ramp, self.p.ramp_t = ref.velocity_ramp((t + 1) * dt), ref.velocity_ramp(t * dt)
        self.p.t_now, self.p.t_next = float(t), float(t + 1)

    def _step(self, p0, pstep, nplanes, run_bc):
        p = self.p
        p.p0, p.pstep, p.nplanes, p.run_bc = p0, pstep, nplanes, int(run_bc)
        rc = self.L.emu_step(C.byref(p), _ptr(self.f[self.cur]), _ptr(self.f[self.cur ^ 1]), _ptr(self.cw), _ptr(self.rho), _ptr(self.u),
                             C.c_longlong(self.stride), _ptr(self.bc_list), _ptr(self.bc_extra), C.c_int(len(self.bc_list)), _ptr(self.uin),
                             _ptr(self.types), _ptr(self.desc), _ptr(self.tav))
        assert rc == 0

    def step_all(self):
        self._step(self.ghost, 1, self.cnt, True)

    def step_faces(self, left=None, right=None):
        """the launches of enqueue_step() before the exchange: k_bc and k_step on the two face planes.  With the
        neighbour slabs given, the fused exchange: k_bc / k_step_faces store the outgoing populations into the neighbours'
        ghost planes themselves (here plain host arrays stand in for the peer-mapped lattices)"""
        p = self.p
        if left is not None:
            for side, nb in enumerate((left, right)):
                p.peer_f[side] = nb.f[nb.cur ^ 1].ctypes.data
                p.peer_stride[side] = nb.stride
                p.peer_P[side] = nb.P
            p.faces = 1
        self._step(1, max(self.cnt - 1, 1), 2 if self.cnt > 1 else 1, True)
        p.faces = 0
        for side in range(2):
            p.peer_f[side] = None

    def step_interior(self):
        self._step(2, 1, self.cnt - 2, False)

    def advance(self):
        self.cur ^= 1
        self.t += 1

    def velsrc(self):
        if (not self.case.regularised) and self.case.velocity_ramp is not None and len(self.vel):
            self.L.emu_velsrc(C.byref(self.p), _ptr(self.vel), C.c_int(len(self.vel)), _ptr(self.types), _ptr(self.desc), _ptr(self.u),
                              C.c_longlong(self.stride), _ptr(self.uin))

    # ---- views of the owned planes in the reference's AoS layout ----
    def owned(self):
        a, b = self.ghost * self.MK, (self.P - self.ghost) * self.MK
        out = {"f": np.ascontiguousarray(self.f[self.cur][:, a:b].T).reshape(-1), "rho": self.rho[a:b].copy(),
               "u": np.ascontiguousarray(self.u[:, a:b].T).reshape(-1), "types": self.types[a:b].copy(), "desc": self.desc[a:b].copy()}
        if self.tav is not None:
            D = self.D
            out["rho_timeav"] = self.tav[0, a:b].copy()
            out["ui_timeav"] = np.ascontiguousarray(self.tav[1:1 + D, a:b].T).reshape(-1)
            out["uiuj_timeav"] = np.ascontiguousarray(self.tav[1 + D:, a:b].T).reshape(-1)
        return out


def exchange(slabs, plans, which):
    """the per-step exchange, driven by the PRODUCT's plan (luma_b200_halo_plan): every send is matched, in issue
    order, with the peer's receive of the same pair -- NCCL's matching rule -- and the plane is copied"""
    world = len(slabs)
    for r in range(world):
        for peer in sorted(set(m["peer"] for m in plans[r])):
            sends = [m for m in plans[r] if m["is_send"] and m["peer"] == peer]
            recvs = [m for m in plans[peer] if (not m["is_send"]) and m["peer"] == r]
            assert len(sends) == len(recvs)
            for s, d in zip(sends, recvs):
                assert s["pop"] == d["pop"]
                MK = slabs[r].MK
                src = slabs[r].f[slabs[r].cur ^ which][s["pop"], s["plane"] * MK:(s["plane"] + 1) * MK]
                slabs[peer].f[slabs[peer].cur ^ which][d["pop"], d["plane"] * MK:(d["plane"] + 1) * MK] = src
