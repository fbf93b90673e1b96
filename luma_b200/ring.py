"""Host-side plumbing of the x-slab ring: bootstrap over torch.distributed and the halo plan.

One process per GPU (the analogue of LUMA's MPI ranks with L_MPI_XCORES = world size,
L_MPI_YCORES = L_MPI_ZCORES = 1).  torch.distributed is used only to bootstrap -- broadcast the
NCCL unique id that luma_b200_comm_init consumes, barriers and max-over-ranks of timings; the
per-step exchange itself is issued natively by the library (NCCL send/recv on its comm stream).

`execute_plan_on_host` replays the library's halo plan (luma_b200_halo_plan) on host tensors with
whatever backend torch.distributed was initialised with; it exists so the protocol (who sends which
population plane to whom, in which order) can be checked with `gloo` on machines without GPUs and
is not part of the time step.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, List, Optional

from . import capi
from .definitions import Definitions


def slab_params(defs: Definitions, rank: int, nranks: int) -> capi.LumaCaseParams:
    p = capi.default_params()
    p.dims, p.num_vels = defs.L_DIMS, defs.L_NUM_VELS
    p.N, p.M, p.K = defs.L_N, defs.L_M, defs.L_K
    p.rank, p.nranks = rank, nranks
    p.x_offset, p.x_count = capi.slab(defs.L_N, nranks, rank)
    p.omega = defs.omega
    return p


def halo_plan(defs: Definitions, rank: int, nranks: int) -> List[dict]:
    """The per-step messages of `rank`, in issue order: dicts with is_send, peer, pop, plane."""
    p = slab_params(defs, rank, nranks)
    n = C.c_int32()
    L = capi.load()
    capi.check(L.luma_b200_halo_plan(C.byref(p), None, 0, C.byref(n)))
    arr = (capi.LumaHaloMsg * max(n.value, 1))()
    capi.check(L.luma_b200_halo_plan(C.byref(p), arr, n.value, C.byref(n)))
    return [dict(is_send=bool(m.is_send), peer=m.peer, pop=m.pop, plane=m.plane) for m in arr[: n.value]]


def broadcast_unique_id(dist, rank: int, make_id: Optional[Callable[[], bytes]] = None) -> bytes:
    """Rank 0 creates the 128-byte ncclUniqueId, everyone receives it (MPI_Bcast in a LUMA MPI build)."""
    from .gridobj import comm_unique_id
    box = [(make_id or comm_unique_id)() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    uid = box[0]
    if not isinstance(uid, (bytes, bytearray)) or len(uid) != 128:
        raise RuntimeError("unique id broadcast failed")
    return bytes(uid)


def attach_p2p(dist, grid, rank: int, nranks: int) -> bool:
    """All-gather the ranks' IPC blobs and hand `grid` those of its ring neighbours (MPI_Allgather in a LUMA
    MPI build): from then on its halo exchange is device-initiated (include/luma_b200.h).

    The decision is COLLECTIVE: peer stores are used only if every rank could export its blob.  Attaching itself
    is then tried on all ranks and the outcomes are agreed on with a second reduction; a rank whose neighbours
    attached while it could not would leave the ring with mixed transports (one side waiting for flags the other
    never writes), so in that case every rank raises.  Returns True when the ring runs on peer stores, False when
    all ranks stay on the NCCL exchange."""
    try:
        blob = grid.p2p_export()
    except Exception:
        blob = b""
    blobs = [None] * nranks
    dist.all_gather_object(blobs, blob)
    if not all(isinstance(b, (bytes, bytearray)) and len(b) == 256 for b in blobs):
        return False                    # same list on every rank: everybody stays on NCCL
    ok = 1
    err = None
    try:
        grid.p2p_attach(blobs[(rank - 1) % nranks], blobs[(rank + 1) % nranks])
    except Exception as ex:             # e.g. no peer access between two GPUs of the box
        ok, err = 0, ex
    flags = [None] * nranks
    dist.all_gather_object(flags, ok)
    if all(flags):
        return True
    if not any(flags):
        return False                    # nobody attached: NCCL everywhere
    raise RuntimeError("p2p_attach succeeded on ranks %s only (%r): recreate the handles and use the NCCL exchange"
                       % ([r for r, f in enumerate(flags) if f], err))


def execute_plan_on_host(dist, plan: List[dict], lattice):
    """lattice: torch tensor [Q, P, M*K] (SoA, ghost planes 0 and P-1) on the backend's device.
    Posts every message of the plan in order as non-blocking p2p and waits (one 'group')."""
    ops = []
    for m in plan:
        buf = lattice[m["pop"], m["plane"]]
        ops.append(dist.P2POp(dist.isend if m["is_send"] else dist.irecv, buf, m["peer"]))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return lattice
