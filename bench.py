#!/usr/bin/env python
"""bench.py -- MLUPS of LUMA's level-0 time step (GridObj::LBM_multi_opt) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c4|c5] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

One JSON line on rank 0.  A "step" is one LBM time step of the whole level-0 grid.

* workload c2 (default): BASELINE.json configs[1], 3-D lid-driven cavity D3Q19 BGK Re=1000, 256^3 cells
  per GPU (at N GPUs the cavity is N*256 x 256 x 256, x-slab per GPU -> weak scaling);
  workload c5: configs[4], 384^3 cells per GPU; c3 / c4: configs[2] / configs[3] at their fixed global size
  (512^3 periodic channel with Guo forcing; 1024x256x256 inlet/outlet + Smagorinsky + cylinder).
* value  : global cells * K / device time of the K steps (CUDA events on the library's stream,
           max over ranks), state resident in HBM.
* e2e    : the same K steps through the reference-facing API with HOST buffers: upload of the host
           state (f, rho, u, LatTyp; pinned memory), LBM_multi_opt in LUMA's output cadence
           (L_GRID_OUT_FREQ = 100 steps) and a download of rho,u into pinned host arrays after every
           interval (luma_b200_download_async: the copy of interval n overlaps the steps of interval
           n+1, two host buffers in turn; everything has landed before the clock stops), wall clock.
* roofline: the dominant kernel (k_step), 304 B per lattice update (19 x 8 B read + 19 x 8 B write,
           DESIGN.md) against the measured copy bandwidth in MEASURED_PEAKS.json.
* cpu_baseline / --impl reference: the UNMODIFIED reference sources compiled as oracle/_ref
           (OpenMP build, all host threads) on a bounded 128^3 sample of the same case.
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
            "c5": "BASELINE configs[4]: weak-scaling cavity D3Q19 BGK"}[name]
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
def run_cpu_reference(warmup: int, steps: int, budget_s: float):
    """Times the compiled reference on a bounded 128^3 sample of the c2 case (same omega); falls back
    to the single-thread C port when the compiled reference is not present."""
    from oracle import port
    from oracle.cases import BENCH_CASES
    cores = os.cpu_count() or 1
    name = "c2_128"
    if port.ref_binary(name, omp=True):
        probe = port.run_ref_bench(name, 1, 2, threads=cores)
        per_step = probe["seconds"] / 2.0
        n = max(1, min(steps, int(budget_s / max(per_step, 1e-6))))
        res = port.run_ref_bench(name, min(warmup, 3), n, threads=cores)
        return {"value": res["mlups"], "unit": UNIT, "cores": int(res["threads"]), "kind": "reference",
                "sample": "unmodified LUMA v1.7.12 LBM_multi_opt (oracle/_ref, -O3 -fopenmp, L_ENABLE_OPENMP), "
                          "128^3 sample of the 256^3 cavity (same omega), %d timed steps after %d warm-up, "
                          "%.1f s" % (n, min(warmup, 3), res["seconds"]),
                "steps_timed": n, "seconds": res["seconds"]}
    case = BENCH_CASES[name]
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
            "steps_timed": n, "seconds": secs}


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    t0 = time.perf_counter()
    cb = run_cpu_reference(args.warmup, args.steps, budget_s=150.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * cb["seconds"] / cb["steps_timed"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.workload, args.gpus),
                   "note": "CPU reference timed on a bounded 128^3 sample of the workload; MLUPS is flat in grid size beyond cache"},
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.perf_counter() - t0,
    }
    emit(line)
    return 0


# ------------------------------------------------------------------------------------------------
def main_ours(args):
    import numpy as np
    import torch
    import luma_b200
    from luma_b200 import capi

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
    uid = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        box = [luma_b200.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        uid = box[0]

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

    defs = workload_defs(args.workload, world)
    Q = defs.L_NUM_VELS
    # algorithmic bytes per lattice update: Q populations read + Q written (two-lattice pull).  The KBC operator also
    # loads the Q populations of the site itself (optimised.cpp:1150), but those are the very values a neighbouring
    # thread pulls, so they come from L1/L2 and DRAM still moves each population once (ncu: DESIGN.md section 4)
    BYTES_PER_LUP = 8.0 * Q * 2
    g = luma_b200.GridObj(defs, rank=rank, nranks=world, device=local, unique_id=uid)
    halo = "none (single GPU)"
    if world > 1:
        nface = {9: 3, 19: 5, 27: 9}[Q]
        halo = "NCCL send/recv of the %d outgoing populations per face" % nface
        if args.halo in ("p2p", "fused"):
            from luma_b200 import ring
            if args.halo == "fused":
                os.environ["LUMA_B200_FUSED_HALO"] = "1"       # read by luma_b200_p2p_attach
            # peer stores need CUDA IPC peer mappings between ring neighbours; where a box cannot provide them every rank
            # falls back to the NCCL exchange together (same kernels, same results: tests/test_gpu_multi.py)
            ok = 1
            try:
                blob = g.p2p_export()
            except Exception:
                blob, ok = b"", 0
            blobs = [None] * world
            dist.all_gather_object(blobs, blob)
            if ok and all(len(b) == 256 for b in blobs):
                try:
                    g.p2p_attach(blobs[(rank - 1) % world], blobs[(rank + 1) % world])
                except Exception as ex:
                    sys.stderr.write("rank %d: p2p_attach failed (%r)\n" % (rank, ex))
                    ok = 0
            else:
                ok = 0
            flag = torch.tensor([ok], dtype=torch.int32, device="cuda")
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag.item()) == 1:
                halo = ("device-initiated: the %d outgoing populations per face stored into the neighbour's ghost plane over "
                        "NVLink (CUDA IPC), arrival flags" % nface)
                if args.halo == "fused":
                    halo += "; stores fused into the face kernels' epilogue"

            else:
                # a handle cannot be detached: start over without peer mappings
                g.close()
                uid = ring.broadcast_unique_id(dist, rank)
                g = luma_b200.GridObj(defs, rank=rank, nranks=world, device=local, unique_id=uid)
                halo += " (peer mapping unavailable on this box)"
    g.LBM_initGrid()
    cells_local = g.x_count * g.M_lim * g.K_lim
    cells_global = defs.L_N * defs.L_M * defs.L_K
    K, W = args.steps, max(args.warmup, 3)

    # host copy of the initial state in pinned memory (the e2e leg uploads it inside its timed region)
    host = None
    if not args.no_e2e:
        pin = lambda n, dt: torch.empty(n, dtype=dt, pin_memory=True).numpy()
        host = {"f": pin(cells_local * g.Q, torch.float64), "rho": pin(cells_local, torch.float64),
                "u": pin(cells_local * g.D, torch.float64), "lt": pin(cells_local, torch.int32)}
        g.download(capi.F | capi.RHO | capi.U, out=host)
        host["lt"][:] = g.LatTyp
        outs = [{"rho": pin(cells_local, torch.float64), "u": pin(cells_local * g.D, torch.float64)} for _ in range(2)]
        bc = defs.boundary_site_descriptors(host["lt"], x_offset=g.x_offset)
        ux, uy, uz = defs.inlet_profiles()

    # ---- device-resident timing ----
    sampler = ClockSampler() if rank == 0 else None
    g.LBM_multi_opt(W)
    barrier()
    g.set_profiling(True)
    l0 = g.stats()["kernel_launches"]
    tw0 = time.time()
    g.LBM_multi_opt(K)
    barrier()
    tw1 = time.time()
    st = g.stats()
    g.set_profiling(False)
    ms = max_over_ranks(st["ms_last_call"])
    launches = st["kernel_launches"] - l0
    props = torch.cuda.get_device_properties(local)
    clocks = sampler.stop(tw0, tw1, str(getattr(props, "uuid", "")) or None, local) if sampler else None
    value = cells_global * K / (ms * 1e-3) / 1e6

    # roofline of the dominant kernel: algorithmic bytes per launch / average launch duration
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    k_ms = st["step_kernel_ms"] / max(st["step_kernel_launches"], 1)
    k_cells = st["step_kernel_cells"] / max(st["step_kernel_launches"], 1)
    achieved = BYTES_PER_LUP * k_cells / (k_ms * 1e-3) / 1e9
    traffic = None
    prof = os.path.join(ROOT, "profiles", "ncu_summary.json")
    if os.path.exists(prof):
        try:
            traffic = json.load(open(prof)).get("k_step_dram_bytes_per_launch_" + args.workload)
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": "k_step<D%dQ%d>" % (defs.L_DIMS, Q), "bytes_per_lup": BYTES_PER_LUP,
                "kernel_ms_avg": k_ms, "kernel_launches_timed": st["step_kernel_launches"], "peak_source": peak_src}

    # ---- end to end through the reference-facing API with host buffers ----
    e2e = None
    if host is not None:
        nint = max(1, K // OUT_FREQ)
        per = min(OUT_FREQ, K)
        barrier()
        t0 = time.perf_counter()
        g.upload(host["f"], host["rho"], host["u"], host["lt"], ux, uy, uz, bc_sites=bc)
        t_up = time.perf_counter() - t0
        for n in range(nint):
            g.LBM_multi_opt(per)
            g.download_async(capi.RHO | capi.U, outs[n % 2])
        g.download_wait()
        barrier()
        secs = max_over_ranks(time.perf_counter() - t0)
        out = outs[0]
        steps_e2e = nint * per
        h2d = (host["f"].nbytes + host["rho"].nbytes + host["u"].nbytes + host["lt"].nbytes) * world
        d2h = (out["rho"].nbytes + out["u"].nbytes) * nint * world
        e2e = {"value": cells_global * steps_e2e / secs / 1e6, "unit": UNIT,
               "h2d_bytes_per_step": h2d / steps_e2e, "d2h_bytes_per_step": d2h / steps_e2e,
               "steps": steps_e2e, "seconds": secs, "upload_seconds": t_up,
               "what": "luma_b200_upload (pinned host f,rho,u,LatTyp) + %d x [%d x LBM_multi_opt + download_async rho,u] + download_wait" % (nint, per)}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            cpu = run_cpu_reference(2, 400, budget_s=20.0)
        except Exception as ex:     # the baseline must never take the GPU number down with it
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "failed: %r" % (ex,)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak" if args.workload in WEAK else "strong",
            "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.workload, world), "cells_per_gpu": cells_local,
                       "omega": g.omega, "parallelism": "x-slab x%d; halo exchange: %s" % (world, halo),
                       "l2": "inputs larger than L2 (2 lattices x %.2f GB per GPU)" % (cells_local * Q * 8 / 1e9),
                       "kernel_variant": "k_step<D3Q%d,%s,%s>" % (Q, "KBC" if defs.L_USE_KBC_COLLISION else ("Smagorinsky" if defs.L_USE_BGKSMAG else "BGK"),
                                                                   "Guo force" if defs.L_GRAVITY_ON else "no force"),
                       "arithmetic": "bit-identical to the reference CPU build (tests/test_gpu_parity.py)"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "pct_hbm_roofline": 100.0 * value * BYTES_PER_LUP / 1e3 / (world * peak),
        }
        emit(line)
    g.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", choices=["c2", "c3", "c4", "c5", "k27"], default="c2")
    ap.add_argument("--halo", choices=["p2p", "nccl", "fused"], default="p2p",
                    help="multi-GPU halo exchange: peer stores by a copy kernel (default), NCCL send/recv, or (experimental) peer stores "
                         "fused into the face kernels' epilogue")
    ap.add_argument("--res", type=int, default=None, help="cavity edge per GPU for c2/c5 (scaling studies; not the named config)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    RES_OVERRIDE = a.res
    sys.exit(main_reference(a) if a.impl == "reference" else main_ours(a))
