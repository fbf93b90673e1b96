"""Worker of tests/test_gpu_multi.py: run under torch.distributed.run, one rank per GPU.
Every rank steps its x-slab on its GPU and compares it bit-for-bit with the same planes of the
serial oracle (the reference's MPI result equals its serial result on core sites)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import luma_b200  # noqa: E402
from luma_b200 import capi, ring  # noqa: E402
from oracle import port  # noqa: E402
from oracle.cases import CASES  # noqa: E402
from util import defs_from_case, first_diff  # noqa: E402


def fullsize(rank, world, local, workload, steps):
    """order-independent checksums of f, rho, u of a full-size bench workload after `steps` steps"""
    import bench
    defs = bench.workload_defs(workload, world)
    uid = ring.broadcast_unique_id(dist, rank) if world > 1 else None
    g = luma_b200.GridObj(defs, rank=rank, nranks=world, device=local, unique_id=uid)
    if world > 1:
        ring.attach_p2p(dist, g, rank, world)
    g.LBM_initGrid()
    g.LBM_multi_opt(steps)
    got = g.download()
    acc = []
    for nm in ("f", "rho", "u"):
        bits = got[nm].view(np.uint64)
        acc += [int(np.add.reduce(bits, dtype=np.uint64)), int(np.bitwise_xor.reduce(bits))]
    F = g.computeLiftDrag()
    g.close()
    if world > 1:
        t = torch.tensor([[a & 0xFFFFFFFF, a >> 32] for a in acc], dtype=torch.int64, device="cuda")
        parts = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(parts, t)
        tot = [0] * len(acc)
        for p in parts:
            for i, (lo, hi) in enumerate(p.cpu().tolist()):
                v = lo | (hi << 32)
                tot[i] = (tot[i] + v) & 0xFFFFFFFFFFFFFFFF if i % 2 == 0 else tot[i] ^ v
        acc = tot
    if rank == 0:
        print("fullsize checksums %s on %d GPU(s): %s" % (workload, world, " ".join("%016x" % a for a in acc)), flush=True)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if sys.argv[1].startswith("fullsize:"):
        _, workload, steps = sys.argv[1].split(":")
        fullsize(rank, world, local, workload, int(steps))
        dist.barrier()
        dist.destroy_process_group()
        return
    names = sys.argv[1].split(",")
    for name in names:
        case = CASES[name]
        defs = defs_from_case(case)
        ref = port.PortGrid(case)
        Q, D, N, MK = case.Q, case.dims, case.N, case.M * case.K
        # both state paths and both transports, one handle (= one NCCL communicator bootstrap) each
        # three transports: peer stores by the copy kernel, NCCL send/recv, peer stores fused into the face kernels' epilogue
        modes = ("device_init+fused", "upload+nccl", "device_init+copy")
        if os.environ.get("LUMA_TEST_NO_FUSED"):
            modes = modes[1:]
        for mode in modes:
            uid = ring.broadcast_unique_id(dist, rank)      # one ncclUniqueId per communicator / handle
            g = luma_b200.GridObj(defs, rank=rank, nranks=world, device=local, unique_id=uid)
            if not mode.endswith("+nccl"):
                os.environ["LUMA_B200_FUSED_HALO"] = "1" if mode.endswith("+fused") else "0"    # read by luma_b200_p2p_attach
                attached = ring.attach_p2p(dist, g, rank, world)       # device-initiated halo exchange; "+nccl" keeps send/recv
                os.environ.pop("LUMA_B200_FUSED_HALO", None)
                assert attached, "no CUDA IPC peer mapping between ring neighbours on this box"
            x0, cnt = g.x_offset, g.x_count
            ref = port.PortGrid(case)
            if mode.startswith("device_init"):
                g.LBM_initGrid()
            else:
                sl = slice(x0 * MK, (x0 + cnt) * MK)
                g.upload(ref.f.reshape(-1, Q)[sl], ref.rho[sl], ref.u.reshape(-1, D)[sl], ref.lattyp[sl],
                         ref.uin(0), ref.uin(1), ref.uin(2))
            done = 0
            for s in ((1, 2, 10, 25) if case.kbc else (1, 2, 10, 50)):      # the reference's KBC operator diverges early
                g.LBM_multi_opt(s - done)
                ref.step(s - done)
                done = s
                got = g.download()
                sl = slice(x0 * MK, (x0 + cnt) * MK)
                for nm, width in (("f", Q), ("rho", 1), ("u", D)):
                    a = got[nm].reshape(-1, width)
                    b = getattr(ref, nm).reshape(-1, width)[sl]
                    assert np.array_equal(a, b), "rank %d %s %s t=%d %s: %s" % (rank, name, mode, s, nm, first_diff(a.ravel(), b.ravel()))
                if case.time_averaged:
                    tav = g.download_timeav()
                    for nm, width in (("rho_timeav", 1), ("ui_timeav", D), ("uiuj_timeav", 3 * D - 3)):
                        a = tav[nm].reshape(-1, width)
                        b = getattr(ref, nm).reshape(-1, width)[sl]
                        assert np.array_equal(a, b), "rank %d %s %s t=%d %s: %s" % (rank, name, mode, s, nm, first_diff(a.ravel(), b.ravel()))
            if case.ld_out:
                # no barrier between the last step and forces(): a rank's share only reads its own planes
                F = torch.tensor(g.computeLiftDrag(), dtype=torch.float64, device="cuda")
                dist.all_reduce(F)
                Fr = ref.force
                assert np.all(np.abs(F.cpu().numpy() - Fr) <= 1e-10 * max(1.0, np.abs(Fr).max())), (F, Fr)
            st = g.stats()
            assert st["halo_bytes_per_step"] == 2 * {9: 3, 19: 5, 27: 9}[Q] * MK * 8
            g.close(); ref.close()
        dist.barrier()
        if rank == 0:
            print("mgpu ok: %s on %d GPUs (%s), bit-identical to the serial oracle" % (name, world, ", ".join(modes)), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
