#!/usr/bin/env python
"""bench.py -- MLUPS of LUMA's level-0 time step (GridObj::LBM_multi_opt) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c5|c2|c3|c4] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

One JSON line on rank 0.  A "step" is one LBM time step of the whole level-0 grid.

* workload c5 (default at every N): BASELINE.json configs[4], the weak-scaling sweep the metric is quoted on --
  the 3-D lid-driven cavity of configs[1] (D3Q19 BGK Re=1000) at 384^3 cells per GPU; at N GPUs the cavity is
  N*384 x 384 x 384, one x-slab per GPU (weak scaling).  At N=1 the line also carries `configs1_256`: configs[1]
  itself (the same cavity at 256^3) measured in the same process.
  workload c2: configs[1] as the primary (256^3 per GPU); c3 / c4: configs[2] / configs[3] at their fixed global size
  (512^3 periodic channel with Guo forcing; 1024x256x256 inlet/outlet + Smagorinsky + cylinder; strong scaling).
* value  : global cells * K / device time of the K steps (CUDA events on the library's stream, max over ranks), state
           resident in HBM, the library exactly as users run it (no per-kernel events).  The kernel figures for the
           roofline come from a second, shorter pass with per-kernel events on.
* parity_check: before the timed region the same ranks, same transport and same kernels step reduced cases and every
           rank compares its slab bit for bit with the serial oracle (checker use of oracle/).
* e2e    : the same K steps through the reference-facing API with HOST buffers, wall clock: luma_b200_upload of
           the host state the way the drop-in shim hands it over at t = 0 (pinned rho, u, LatTyp; f_aos = NULL
           because f = feq(rho,u) there, see include/luma_b200.h), LBM_multi_opt in LUMA's output cadence
           (L_GRID_OUT_FREQ = 100 steps) and a download of rho,u into pinned host arrays after every interval
           (luma_b200_download_async: the copy of interval n overlaps the steps of interval n+1; everything has
           landed before the clock stops).  `e2e.upload_full_f` repeats it with the 8*Q B/site of f uploaded too
           (the restart path); `e2e_dropin` is the UNMODIFIED LUMA host loop calling the C ABI once per step.
* roofline: the dominant kernel (k_step), 304 B per lattice update (19 x 8 B read + 19 x 8 B write,
           DESIGN.md) against the measured copy bandwidth in MEASURED_PEAKS.json.
* cpu_baseline / --impl reference: the UNMODIFIED reference sources compiled as oracle/_ref (OpenMP build, one
           pinned thread per physical core) on the workload's per-GPU grid when host memory allows (a smaller
           sample of the same case otherwise -- stated in `sample`).
"""
from __future__ import annotations

import argparse
import datetime
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
# stdout carries exactly one JSON line.  Native libraries write there too (NCCL prints its "NCCL version ..."
# banner on stdout), so file descriptor 1 is pointed at stderr for the whole run and the line goes to the
# saved descriptor at the end.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line: dict) -> None:
    sys.stdout.flush()
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())

METRIC = "MLUPS (D3Q19 fp64)"
UNIT = "MLUPS"
BYTES_PER_LUP = 304.0          # D3Q19: 19 populations x 8 B read + 19 x 8 B written (two-lattice pull), DESIGN.md;
                               # main_ours() recomputes it from the workload's lattice and collision operator
OUT_FREQ = 100                 # L_GRID_OUT_FREQ used by the e2e leg


WEAK = ("c2", "c5", "k27")     # cells per GPU fixed; c3 / c4 have a fixed global grid (strong scaling)
DEFAULT_WORKLOAD = "c5"


RES_OVERRIDE = None            # --res: cells per GPU edge of the cavity workloads (studies only; the default is the named size)


def workload_defs(name: str, ngpus: int):
    """SURVEY.md 8(d) table of synthetic inputs, as definitions.h macros."""
    import luma_b200
    if name == "c3":
        # configs[2]: periodic channel 512^3, bounce-back walls in y, Guo forcing along x; nu_lbm = 0.05
        # (omega 1.538), gravity such that the Poiseuille maximum is u_lbm ~ 0.05
        res = 512
        return luma_b200.Definitions(
            L_DIMS=3, L_RESOLUTION=res, L_TIMESTEP=0.05 / res, L_BX=1.0, L_BY=1.0, L_BZ=1.0,
            L_RE=None, L_NU=1.0 / res, L_NO_FLOW=True,
            L_WALL_LEFT=luma_b200.eFluid, L_WALL_RIGHT=luma_b200.eFluid, L_WALL_FRONT=luma_b200.eFluid,
            L_WALL_BACK=luma_b200.eFluid, L_WALL_THICKNESS_CELLS=(0, 0, 1, 1, 0, 0),
            L_GRAVITY_ON=True, L_GRAVITY_FORCE=0.0158, L_GRAVITY_DIRECTION=0)
    if name == "c4":
        # configs[3]: 1024x256x256, velocity inlet / pressure outlet (regularised), Smagorinsky LES, velocity
        # ramp, square cylinder 32x32 spanning z at x ~ 256; omega 1.98
        res = 256
        return luma_b200.Definitions(
            L_DIMS=3, L_RESOLUTION=res, L_TIMESTEP=0.05 / res, L_BX=4.0, L_BY=1.0, L_BZ=1.0,
            L_RE=7600.0, L_NO_FLOW=True, L_USE_BGKSMAG=True, L_CSMAG=0.3, L_VELOCITY_RAMP=0.5,
            L_WALL_LEFT=luma_b200.eVelocity, L_WALL_RIGHT=luma_b200.ePressure, L_WALL_FRONT=luma_b200.eFluid,
            L_WALL_BACK=luma_b200.eFluid, L_WALL_THICKNESS_CELLS=(1, 1, 1, 1, 0, 0),
            body_box=(256, 288, 112, 144, 0, 256))
    if name == "k27":
        # SURVEY 8(f-4) KBC row: periodic channel on D3Q27 with the KBC-N4 operator (non-regularised, as the reference
        # demands on D3Q27), bounce-back walls in y, Guo forcing; 256^3 per GPU
        res = RES_OVERRIDE or 256
        return luma_b200.Definitions(
            L_DIMS=3, L_RESOLUTION=res, L_TIMESTEP=0.05 / res, L_BX=float(ngpus), L_BY=1.0, L_BZ=1.0,
            L_RE=None, L_NU=2.0 / res, L_NO_FLOW=True, L_USE_KBC_COLLISION=True, L_REGULARISED_BOUNDARIES=False,
            L_WALL_LEFT=luma_b200.eFluid, L_WALL_RIGHT=luma_b200.eFluid, L_WALL_FRONT=luma_b200.eFluid,
            L_WALL_BACK=luma_b200.eFluid, L_WALL_THICKNESS_CELLS=(0, 0, 1, 1, 0, 0),
            L_GRAVITY_ON=True, L_GRAVITY_FORCE=0.0158, L_GRAVITY_DIRECTION=0)
    res = RES_OVERRIDE or {"c2": 256, "c5": 384}[name]
    return luma_b200.Definitions(
        L_DIMS=3, L_RESOLUTION=res, L_TIMESTEP=0.05 / res, L_BX=float(ngpus), L_BY=1.0, L_BZ=1.0,
        L_RE=1000.0, L_UX0=1.0, L_WALL_TOP=luma_b200.eVelocity, L_REGULARISED_BOUNDARIES=True, L_NO_FLOW=True)


def workload_name(name: str, ngpus: int) -> str:
    if name == "c3":
        return "BASELINE configs[2]: 3D periodic channel D3Q19, bounce-back walls, Guo forcing, 512x512x512 cells over %d x-slab(s)" % ngpus
    if name == "c4":
        return ("BASELINE configs[3]: flow past a square cylinder D3Q19, velocity inlet / pressure outlet, Smagorinsky LES, "
                "1024x256x256 cells over %d x-slab(s)" % ngpus)
    if name == "k27":
        res = RES_OVERRIDE or 256
        return ("SURVEY 8(f-4): periodic channel on D3Q27 with the KBC-N4 collision operator, bounce-back walls, Guo forcing, "
                "%dx%dx%d cells (%d^3 per GPU, x-slabs)" % (res * ngpus, res, res, res))
    res = RES_OVERRIDE or {"c2": 256, "c5": 384}[name]
    base = {"c2": "BASELINE configs[1]: 3D lid-driven cavity D3Q19 BGK Re=1000",
            "c5": "BASELINE configs[4]: weak-scaling sweep, 3D lid-driven cavity D3Q19 BGK Re=1000 (the case of configs[1])"}[name]
    return "%s, %dx%dx%d cells (%d^3 per GPU, x-slabs)" % (base, res * ngpus, res, res, res)


# ------------------------------------------------------------------------------------------------
# clocks: nvidia-smi sampled while the timed region runs
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("timestamp,index,uuid,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, period_ms=50):
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits", "-lms", str(period_ms)],
                stdout=self.tmp, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self, t0: float, t1: float, uuid: str | None, index: int):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.06)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.tmp.flush()
        rows = []
        with open(self.tmp.name) as fh:
            for line in fh:
                p = [x.strip() for x in line.split(",")]
                if len(p) < 10:
                    continue
                try:
                    ts = datetime.datetime.strptime(p[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                    rows.append((ts, int(p[1]), p[2], float(p[3]), float(p[4]), p[6:10]))
                except Exception:
                    continue
        os.unlink(self.tmp.name)
        mine = [r for r in rows if (uuid and uuid in r[2])] or [r for r in rows if r[1] == index]
        win = [r for r in mine if t0 - 0.02 <= r[0] <= t1 + 0.02] or mine
        if not win:
            return out
        mhz = sorted(r[3] for r in win)
        out["sm_mhz"] = mhz[len(mhz) // 2]
        out["sm_max_mhz"] = max(r[4] for r in win)
        out["samples"] = len(win)
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        out["reasons"] = [n for a, n in enumerate(names) if any(r[5][a].lower().startswith("active") for r in win)]
        return out


# ------------------------------------------------------------------------------------------------
# the reference's CPU implementation (oracle/_ref: unmodified LUMA sources, OpenMP build)
# ------------------------------------------------------------------------------------------------
def physical_cores():
    """(threads to use, description): one thread per physical core this process may run on."""
    try:
        allowed = sorted(os.sched_getaffinity(0))
    except Exception:
        allowed = list(range(os.cpu_count() or 1))
    cores = set()
    try:
        for cpu in allowed:
            base = "/sys/devices/system/cpu/cpu%d/topology/" % cpu
            with open(base + "physical_package_id") as fh:
                pkg = fh.read().strip()
            with open(base + "core_id") as fh:
                core = fh.read().strip()
            cores.add((pkg, core))
    except Exception:
        cores = set()
    n = len(cores) if cores else len(allowed)
    return max(1, n), "%d logical CPUs allowed, %d physical cores" % (len(allowed), n)


def host_mem_available_gb() -> float:
    try:
        with open("/proc/meminfo") as fh:
            for line in fh:
                if line.startswith("MemAvailable:"):
                    return float(line.split()[1]) / 1e6
    except Exception:
        pass
    return 0.0


# per workload: (case of oracle.cases.BENCH_CASES, host GB the reference needs for it, what it is) in order of preference
CPU_CASES = {
    "c5": [("c5_384", 40.0, "the workload's per-GPU grid (384^3)"), ("c2_256", 12.0, "256^3 sample of the 384^3-per-GPU cavity (same Re)"),
           ("c2_128", 2.0, "128^3 sample of the cavity (same omega)")],
    "c2": [("c2_256", 12.0, "the workload's per-GPU grid (256^3)"), ("c2_128", 2.0, "128^3 sample of the 256^3 cavity (same omega)")],
    "c3": [("c3_r128", 2.0, "128^3 sample of the 512^3 channel (same macros)"), ("c2_128", 2.0, "128^3 cavity (channel build missing)")],
    "c4": [("c4_r64", 1.5, "256x64x64 sample of the 1024x256x256 cylinder case (same macros, same omega)"), ("c2_128", 2.0, "128^3 cavity (cylinder build missing)")],
    "k27": [("c2_128", 2.0, "128^3 D3Q19 cavity (no D3Q27 timing build)")],
}


def run_cpu_reference(workload: str, warmup: int, steps: int, budget_s: float):
    """Times the compiled reference (OpenMP build, threads pinned one per physical core) on the workload's per-GPU
    grid when the host has the memory for it, else on a smaller sample of the same case; falls back to the
    single-thread C port when no compiled reference is present."""
    from oracle import port
    from oracle.cases import BENCH_CASES
    threads, cpu_desc = physical_cores()
    free_gb = host_mem_available_gb()
    env = {"OMP_NUM_THREADS": str(threads), "OMP_PROC_BIND": "close", "OMP_PLACES": "cores", "OMP_DYNAMIC": "false"}
    forced = os.environ.get("LUMA_BENCH_CPU_CASE")      # tests: a small sample whatever the host could hold
    for name, need_gb, what in CPU_CASES.get(workload, CPU_CASES["c2"]):
        if not port.ref_binary(name, omp=True) or free_gb < need_gb or (forced and name != forced):
            continue
        case = BENCH_CASES[name]
        cells = case.N * case.M * case.K
        est = cells / (1.2e6 * threads)                 # ~1-2 MLUPS per core is what this build reaches (memory-bound AoS)
        n = max(3, min(steps, int(budget_s / max(est, 1e-6))))
        w = max(1, min(warmup, 3))
        os.environ.update(env)
        res = port.run_ref_bench(name, w, n, threads=threads)
        return {"value": res["mlups"], "unit": UNIT, "cores": int(res["threads"]), "kind": "reference",
                "sample": "unmodified LUMA v1.7.12 LBM_multi_opt (oracle/_ref, -O3 -fopenmp, L_ENABLE_OPENMP: the OpenMP stand-in for "
                          "the MPI build, no MPI in the image), %s: %dx%dx%d cells, %d timed steps after %d warm-up, %.1f s; "
                          "OMP_NUM_THREADS=%d OMP_PROC_BIND=close OMP_PLACES=cores (%s)"
                          % (what, case.N, case.M, case.K, n, w, res["seconds"], threads, cpu_desc),
                "case": name, "same_grid_as_gpu_arm": what.startswith("the workload"),
                "steps_timed": n, "seconds": res["seconds"], "host_mem_available_gb": free_gb}
    case = BENCH_CASES["c2_128"]
    g = port.PortGrid(case)
    g.step(1)
    t0 = time.perf_counter()
    g.step(1)
    per_step = time.perf_counter() - t0
    n = max(1, min(steps, int(budget_s / max(per_step, 1e-6))))
    t0 = time.perf_counter()
    g.step(n)
    secs = time.perf_counter() - t0
    return {"value": case.N * case.M * case.K * n / secs / 1e6, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": "C restatement (oracle/luma_oracle.c), 128^3 sample, %d steps, %.1f s" % (n, secs),
            "case": "c2_128", "same_grid_as_gpu_arm": False, "steps_timed": n, "seconds": secs}


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    t0 = time.perf_counter()
    cb = run_cpu_reference(args.workload, args.warmup, max(args.steps, 3), budget_s=120.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * cb["seconds"] / cb["steps_timed"], "higher_is_better": True,
        "scaling": "weak" if args.workload in WEAK else "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.workload, args.gpus),
                   "note": "host-core baseline on one GPU's share of the workload (or a stated smaller sample): see cpu_baseline.sample"},
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.perf_counter() - t0,
    }
    emit(line)
    return 0


# ------------------------------------------------------------------------------------------------
# correctness evidence inside the bench line: reduced cases on the same ranks, transport and kernels vs the oracle
# ------------------------------------------------------------------------------------------------
PARITY_CASES = (("chan3d", 60), ("cyl3d", 60), ("cav3d_32", 40))


def parity_check(rank, world, local, dist, halo_mode, ring):
    """Every rank steps its x-slab of each case and compares f, rho, u of its planes bit for bit with the serial
    oracle (oracle/luma_oracle.c, itself pinned to the compiled reference) -- the checker, not the measured path."""
    import numpy as np
    import torch
    import luma_b200
    from oracle import port
    from oracle.cases import CASES
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from util import defs_from_case
    done, ok_all, transport = [], True, "none (single GPU)"
    for name, steps in PARITY_CASES:
        case = CASES[name]
        if case.N // world < 4 and world > 1:
            continue
        ref = port.PortGrid(case)
        uid = ring.broadcast_unique_id(dist, rank) if world > 1 else None
        g = luma_b200.GridObj(defs_from_case(case), rank=rank, nranks=world, device=local, unique_id=uid)
        if world > 1:
            transport = "nccl"
            if halo_mode in ("p2p", "fused"):
                os.environ["LUMA_B200_FUSED_HALO"] = "1" if halo_mode == "fused" else "0"
                if ring.attach_p2p(dist, g, rank, world):
                    transport = halo_mode
                os.environ.pop("LUMA_B200_FUSED_HALO", None)
        MK, Q, D = case.M * case.K, case.Q, case.dims
        sl = slice(g.x_offset * MK, (g.x_offset + g.x_count) * MK)
        g.upload(ref.f.reshape(-1, Q)[sl], ref.rho[sl], ref.u.reshape(-1, D)[sl], ref.lattyp[sl], ref.uin(0), ref.uin(1), ref.uin(2))
        ok = True
        for chunk in (1, steps - 1):            # a single-step call, then the rest
            g.LBM_multi_opt(chunk)
            ref.step(chunk)
            got = g.download()
            for nm, width in (("f", Q), ("rho", 1), ("u", D)):
                ok = ok and np.array_equal(got[nm].reshape(-1, width), getattr(ref, nm).reshape(-1, width)[sl])
        g.close(); ref.close()
        if dist is not None:
            flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device="cuda")
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            ok = bool(int(flag.item()))
        ok_all = ok_all and ok
        done.append({"case": name, "steps": steps, "bitwise": ok})
    return {"cases": done, "bitwise": bool(ok_all and done), "ranks": world, "transport": transport,
            "oracle": "oracle/luma_oracle.c (pinned bit-for-bit to the compiled reference: tests/test_oracle_pinned.py)",
            "compared": "f, rho, u of every rank's planes after 1 and after all steps"}


# ------------------------------------------------------------------------------------------------
def make_grid(defs, rank, world, local, dist, halo_mode, ring):
    """A handle on this rank's slab with the requested halo transport; returns (grid, description of the transport)."""
    import luma_b200
    Q = defs.L_NUM_VELS
    uid = ring.broadcast_unique_id(dist, rank) if world > 1 else None
    g = luma_b200.GridObj(defs, rank=rank, nranks=world, device=local, unique_id=uid)
    halo = "none (single GPU)"
    if world > 1:
        nface = {9: 3, 19: 5, 27: 9}[Q]
        halo = "NCCL send/recv of the %d outgoing populations per face" % nface
        if halo_mode in ("p2p", "fused"):
            os.environ["LUMA_B200_FUSED_HALO"] = "1" if halo_mode == "fused" else "0"       # read by luma_b200_p2p_attach
            # peer stores need CUDA IPC peer mappings between ring neighbours; ring.attach_p2p agrees on the outcome across
            # the ranks, and where a box cannot provide the mappings every rank uses the NCCL exchange together
            if ring.attach_p2p(dist, g, rank, world):
                halo = ("device-initiated: the %d outgoing populations per face stored into the neighbour's ghost plane over "
                        "NVLink (CUDA IPC), arrival flags" % nface)
                halo += ("; stores fused into the face kernels' epilogue" if halo_mode == "fused" else "; separate copy kernel")
            else:
                halo += " (peer mapping unavailable on this box)"
            os.environ.pop("LUMA_B200_FUSED_HALO", None)
    return g, halo


def measure_resident(g, K, W, barrier, max_over_ranks, profile_steps):
    """device-resident timing: W warm-up steps, K timed steps with the library as users run it, then a short pass with
    per-kernel events for the roofline of the dominant kernel"""
    g.LBM_multi_opt(W)
    g.sync()
    barrier()
    l0 = g.stats()["kernel_launches"]
    tw0 = time.time()
    g.LBM_multi_opt(K)
    st = g.stats()                      # read point: submits the held-back step, waits for the K steps
    barrier()
    tw1 = time.time()
    ms = max_over_ranks(st["ms_last_call"])
    launches = st["kernel_launches"] - l0
    g.set_profiling(True)
    g.LBM_multi_opt(profile_steps)
    sp = g.stats()
    g.set_profiling(False)
    return ms, launches, sp, (tw0, tw1)


def main_ours(args):
    import numpy as np
    import torch
    import luma_b200
    from luma_b200 import capi, ring

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus %d must be launched with torch.distributed.run --nproc-per-node %d" % (args.gpus, args.gpus))
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- correctness first: reduced cases on these ranks with this transport against the oracle ----
    parity = None
    if not args.no_parity:
        try:
            parity = parity_check(rank, world, local, dist, args.halo, ring)
        except ImportError as ex:       # oracle/libluma_oracle.so not built on this box: say so, never pretend
            parity = {"cases": [], "bitwise": None, "ranks": world, "error": "oracle unavailable: %r" % (ex,)}

    defs = workload_defs(args.workload, world)
    Q = defs.L_NUM_VELS
    # algorithmic bytes per lattice update: Q populations read + Q written (two-lattice pull).  The KBC operator also
    # loads the Q populations of the site itself (optimised.cpp:1150), but those are the very values a neighbouring
    # thread pulls, so they come from L1/L2 and DRAM still moves each population once (ncu: DESIGN.md section 4)
    BYTES_PER_LUP = 8.0 * Q * 2
    g, halo = make_grid(defs, rank, world, local, dist, args.halo, ring)
    g.LBM_initGrid()
    cells_local = g.x_count * g.M_lim * g.K_lim
    cells_global = defs.L_N * defs.L_M * defs.L_K
    K, W = args.steps, max(args.warmup, 3)

    # ---- device-resident timing ----
    sampler = ClockSampler() if rank == 0 else None
    ms, launches, sp, (tw0, tw1) = measure_resident(g, K, W, barrier, max_over_ranks, min(K, 50))
    props = torch.cuda.get_device_properties(local)
    clocks = sampler.stop(tw0, tw1, str(getattr(props, "uuid", "")) or None, local) if sampler else None
    value = cells_global * K / (ms * 1e-3) / 1e6

    # roofline of the dominant kernel: algorithmic bytes per launch / average launch duration
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"

    def roofline_of(sp, bytes_per_lup, kernel, workload):
        k_ms = sp["step_kernel_ms"] / max(sp["step_kernel_launches"], 1)
        k_cells = sp["step_kernel_cells"] / max(sp["step_kernel_launches"], 1)
        achieved = bytes_per_lup * k_cells / (k_ms * 1e-3) / 1e9
        traffic = None
        prof = os.path.join(ROOT, "profiles", "ncu_summary.json")
        if os.path.exists(prof):
            try:
                summ = json.load(open(prof))
                # per-launch DRAM bytes of the ncu capture, valid only for the kernel build it was taken from
                if summ.get("kernel_source_sha256") == luma_b200.kernel_fingerprint():
                    traffic = summ.get("k_step_dram_bytes_per_launch_" + workload)
            except Exception:
                traffic = None
        return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": kernel, "bytes_per_lup": bytes_per_lup,
                "kernel_ms_avg": k_ms, "kernel_launches_timed": sp["step_kernel_launches"], "peak_source": peak_src,
                "how": "CUDA events around every k_step launch on its stream in a separate pass of %d steps (luma_b200_set_profiling)" % sp["step_kernel_launches"]}

    variant = "k_step<D3Q%d,%s,%s>" % (Q, "KBC" if defs.L_USE_KBC_COLLISION else ("Smagorinsky" if defs.L_USE_BGKSMAG else "BGK"),
                                       "Guo force" if defs.L_GRAVITY_ON else "no force")
    roofline = roofline_of(sp, BYTES_PER_LUP, variant, args.workload)

    # ---- end to end through the reference-facing API with host buffers ----
    e2e = None
    if not args.no_e2e:
        pin = lambda n, dt: torch.empty(n, dtype=dt, pin_memory=True).numpy()
        # the host's state at t = 0 as LBM_initGrid leaves it: rho, u, LatTyp (and f = feq(rho,u), which is not uploaded)
        g.close()
        g, halo = make_grid(defs, rank, world, local, dist, args.halo, ring)
        g.LBM_initGrid()
        host = {"rho": pin(cells_local, torch.float64), "u": pin(cells_local * g.D, torch.float64), "lt": pin(cells_local, torch.int32)}
        g.download(capi.RHO | capi.U, out=host)
        host["lt"][:] = g.LatTyp
        outs = [{"rho": pin(cells_local, torch.float64), "u": pin(cells_local * g.D, torch.float64)} for _ in range(2)]
        bc = defs.boundary_site_descriptors(host["lt"], x_offset=g.x_offset)
        ux, uy, uz = defs.inlet_profiles()
        nint = max(1, K // OUT_FREQ)
        per = min(OUT_FREQ, K)

        def run_e2e(f_host):
            barrier()
            t0 = time.perf_counter()
            g.upload(f_host, host["rho"], host["u"], host["lt"], ux, uy, uz, bc_sites=bc)
            t_up = time.perf_counter() - t0
            for n in range(nint):
                g.LBM_multi_opt(per)
                g.download_async(capi.RHO | capi.U, outs[n % 2])
            g.download_wait()
            barrier()
            return max_over_ranks(time.perf_counter() - t0), t_up

        run_e2e(None)                   # untimed: staging buffers, page tables
        secs, t_up = run_e2e(None)
        steps_e2e = nint * per
        h2d = (host["rho"].nbytes + host["u"].nbytes + host["lt"].nbytes) * world
        d2h = (outs[0]["rho"].nbytes + outs[0]["u"].nbytes) * nint * world
        e2e = {"value": cells_global * steps_e2e / secs / 1e6, "unit": UNIT,
               "h2d_bytes_per_step": h2d / steps_e2e, "d2h_bytes_per_step": d2h / steps_e2e,
               "steps": steps_e2e, "seconds": secs, "upload_seconds": t_up,
               "what": "luma_b200_upload (pinned host rho,u,LatTyp; f_aos=NULL: f = feq(rho,u) at t = 0, evaluated on the device) + "
                       "%d x [%d x LBM_multi_opt + download_async rho,u] + download_wait" % (nint, per)}
        if args.e2e_full_f and cells_local * g.Q * 8 < 0.25 * host_mem_available_gb() * 1e9:
            fh = pin(cells_local * g.Q, torch.float64)
            g.upload(None, host["rho"], host["u"], host["lt"], ux, uy, uz, bc_sites=bc)
            g.download(capi.F, out={"f": fh})
            secs_f, t_up_f = run_e2e(fh)
            e2e["upload_full_f"] = {"value": cells_global * steps_e2e / secs_f / 1e6, "seconds": secs_f, "upload_seconds": t_up_f,
                                    "h2d_bytes_per_step": (h2d + fh.nbytes * world) / steps_e2e,
                                    "what": "the same with the populations uploaded too (restart path: 8*Q B per site more)"}
            del fh

    # ---- configs[1] itself (256^3) next to the 384^3-per-GPU primary, same process ----
    configs1 = None
    if world == 1 and args.workload == "c5" and not args.no_configs1 and RES_OVERRIDE is None:
        g.close()
        d2 = workload_defs("c2", 1)
        g = luma_b200.GridObj(d2, device=local)
        g.LBM_initGrid()
        ms2, _, sp2, _ = measure_resident(g, K, W, barrier, max_over_ranks, min(K, 50))
        c2cells = d2.L_N * d2.L_M * d2.L_K
        configs1 = {"workload": workload_name("c2", 1), "value": c2cells * K / (ms2 * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": ms2 / K,
                    "roofline": roofline_of(sp2, BYTES_PER_LUP, variant, "c2")}

    # ---- the unmodified LUMA host loop calling the ABI once per step (drop-in binary), N = 1 ----
    e2e_dropin = None
    if world == 1 and not args.no_e2e and not args.no_dropin:
        g.close()
        e2e_dropin = run_dropin_bench(W, K)

    cpu = None
    if rank == 0 and not args.no_cpu:
        try:
            cpu = run_cpu_reference(args.workload, 2, 30, budget_s=20.0 if world == 1 else 10.0)
        except Exception as ex:     # the baseline must never take the GPU number down with it
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "failed: %r" % (ex,)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak" if args.workload in WEAK else "strong",
            "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.workload, world), "cells_per_gpu": cells_local,
                       "omega": defs.omega, "parallelism": "x-slab x%d; halo exchange: %s" % (world, halo),
                       "l2": "inputs larger than L2 (2 lattices x %.2f GB per GPU)" % (cells_local * Q * 8 / 1e9),
                       "kernel_variant": variant,
                       "arithmetic": "bit-identical to the reference CPU build: this run's parity_check (reduced cases, same ranks and "
                                     "transport); at size by golden digests of the compiled reference (tests/test_gpu_parity.py::"
                                     "test_parity_at_size: 256^3 x 10 steps, 128^3 x 100, 256x64x64 x 1000)"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "parity_check": parity,
            "pct_hbm_roofline": 100.0 * value * BYTES_PER_LUP / 1e3 / (world * peak),
        }
        if configs1 is not None:
            line["configs1_256"] = configs1
        if e2e_dropin is not None:
            line["e2e_dropin"] = e2e_dropin
        emit(line)
    try:
        g.close()
    except Exception:
        pass
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def run_dropin_bench(warmup: int, steps: int):
    """oracle/_ref/luma_dropin_c2_256: the UNMODIFIED LUMA host (GridManager, GridObj, ObjectManager, its own time loop,
    src/main_lbm.cpp:422-572) linked with luma_b200/host/GridObj_ops_lbm_b200.cpp -- LBM_multi_opt() once per step through
    the C ABI, upload at the first call, rho/u downloaded when the loop ends.  None when the binary was not built
    (needs /root/reference at build time)."""
    exe = os.path.join(ROOT, "oracle", "_ref", "luma_dropin_c2_256")
    if not os.path.exists(exe) or host_mem_available_gb() < 14.0:
        return None
    try:
        out = subprocess.run([exe, "bench", str(int(warmup)), str(int(steps))], check=True, stdout=subprocess.PIPE,
                             stderr=subprocess.DEVNULL, timeout=600).stdout.decode()
        d = json.loads(out.strip().splitlines()[-1])
    except Exception as ex:
        return {"value": None, "error": repr(ex)}
    return {"value": d["mlups_e2e"], "unit": UNIT, "steps": d["steps"], "seconds": d["seconds_e2e"],
            "steps_only_mlups": d["mlups"], "upload_seconds": d.get("first_call_seconds"), "per_call_us": d.get("per_call_us"),
            "what": "unmodified LUMA host loop, 256^3 cavity (configs[1]): first LBM_multi_opt call (case description + upload of "
                    "rho,u,LatTyp) + %d x LBM_multi_opt() one step per call + final host sync of rho,u; steps_only_mlups excludes the "
                    "first call and the final download; per_call_us = host time inside LBM_multi_opt per step" % d["steps"]}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", choices=["c2", "c3", "c4", "c5", "k27"], default=DEFAULT_WORKLOAD)
    ap.add_argument("--halo", choices=["p2p", "nccl", "fused"], default="fused",
                    help="multi-GPU halo exchange: peer stores fused into the face kernels' epilogue (default), peer stores by a separate "
                         "copy kernel (p2p), or NCCL send/recv")
    ap.add_argument("--res", type=int, default=None, help="cavity edge per GPU for c2/c5 (scaling studies; not the named config)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the reduced-case parity check against the oracle")
    ap.add_argument("--no-configs1", action="store_true", help="N=1, workload c5: skip the secondary 256^3 measurement")
    ap.add_argument("--no-dropin", action="store_true", help="N=1: skip the drop-in binary's end-to-end figure")
    ap.add_argument("--e2e-full-f", action="store_true", help="also time the e2e leg with the populations uploaded")
    a = ap.parse_args()
    RES_OVERRIDE = a.res
    sys.exit(main_reference(a) if a.impl == "reference" else main_ours(a))
