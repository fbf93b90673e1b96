"""World-size-2 (and 3) `gloo` tests of the N>1 host logic, on CPU: slab decomposition, ring
bootstrap and the halo plan the library issues every step (luma_b200_halo_plan), replayed on host
tensors and checked against the global periodic lattice of the oracle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from luma_b200 import capi, ring
from oracle import port
from oracle.cases import CASES
from util import defs_from_case


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port_no, name, q):
    try:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port_no)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        case = CASES[name]
        defs = defs_from_case(case)
        # bootstrap: the 128-byte id travels from rank 0 to everyone (id factory injected: no NCCL on CPU boxes)
        uid = ring.broadcast_unique_id(dist, rank, make_id=lambda: bytes(range(128)))
        assert uid == bytes(range(128))

        # every rank holds the oracle's global state (stand-in for LUMA's per-rank arrays)
        g = port.PortGrid(case)
        g.step(3)
        Q, N, M, K = case.Q, case.N, case.M, case.K
        f = g.f.reshape(N, M * K, Q)                       # AoS -> [x, site-in-plane, v]
        x0, cnt = capi.slab(N, world, rank)
        # local SoA lattice with ghost planes, owned planes filled, ghosts poisoned
        lat = torch.full((Q, cnt + 2, M * K), float("nan"), dtype=torch.float64)
        lat[:, 1:cnt + 1, :] = torch.from_numpy(np.ascontiguousarray(f[x0:x0 + cnt].transpose(2, 0, 1)))
        plan = ring.halo_plan(defs, rank, world)
        ring.execute_plan_on_host(dist, plan, lat)

        cx = {19: (1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 0, 0, 0, 0, 1, -1, -1, 1, 0), 9: (1, -1, 0, 0, 1, -1, 1, -1, 0),
              # D3Q27 (src/stdafx.cpp:41-70)
              27: (1, -1, 0, 0, 0, 0, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1, 1, -1, -1, 1, -1, 1, 1, -1, 0)}[Q]
        lo, hi = (x0 - 1) % N, (x0 + cnt) % N
        for v in range(Q):
            if cx[v] == 1:      # pulled from x-1: must be in the low ghost plane
                assert np.array_equal(lat[v, 0].numpy(), f[lo, :, v]), ("low ghost", v)
                assert np.isnan(lat[v, cnt + 1].numpy()).all()
            elif cx[v] == -1:   # pulled from x+1: high ghost plane
                assert np.array_equal(lat[v, cnt + 1].numpy(), f[hi, :, v]), ("high ghost", v)
                assert np.isnan(lat[v, 0].numpy()).all()
            else:               # never exchanged
                assert np.isnan(lat[v, 0].numpy()).all() and np.isnan(lat[v, cnt + 1].numpy()).all()
        # bytes on the wire: 5 (3) populations per face instead of the reference's 19 (9)
        sends = [m for m in plan if m["is_send"]]
        assert len(sends) == {19: 10, 9: 6, 27: 18}[Q]
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as e:   # pragma: no cover
        import traceback
        q.put((rank, "FAIL: " + traceback.format_exc()))


@pytest.mark.parametrize("name,world", [("chan3d", 2), ("chan2d", 2), ("cav3d_32", 3), ("kbc3d_chan", 2)])
def test_halo_plan_over_gloo(name, world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port_no = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port_no, name, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res


def test_plan_is_empty_for_a_single_rank_and_symmetric_otherwise():
    defs = defs_from_case(CASES["chan3d"])
    assert ring.halo_plan(defs, 0, 1) == []
    for world in (2, 4, 8):
        plans = [ring.halo_plan(defs, r, world) for r in range(world)]
        for r, plan in enumerate(plans):
            for m in plan:
                if m["is_send"]:
                    # the peer must post the matching receive of the same population
                    assert any((not o["is_send"]) and o["peer"] == r and o["pop"] == m["pop"] for o in plans[m["peer"]])
        # per peer, sends and the peer's receives are issued in the same population order (NCCL matches in order)
        for r in range(world):
            for peer in set(m["peer"] for m in plans[r]):
                s = [m["pop"] for m in plans[r] if m["is_send"] and m["peer"] == peer]
                rcv = [m["pop"] for m in plans[peer] if (not m["is_send"]) and m["peer"] == r]
                assert s == rcv, (world, r, peer)


class _FakeGrid:
    """stands in for a GridObj in the transport negotiation (no GPU here): export / attach succeed or fail on demand"""
    def __init__(self, export_ok, attach_ok):
        self.export_ok, self.attach_ok, self.attached = export_ok, attach_ok, False

    def p2p_export(self):
        if not self.export_ok:
            raise RuntimeError("no IPC on this device")
        return bytes(256)

    def p2p_attach(self, left, right):
        assert len(left) == 256 and len(right) == 256
        if not self.attach_ok:
            raise RuntimeError("cudaIpcOpenMemHandle failed")
        self.attached = True


SCENARIOS = (("all_ok", True), ("export_fails_on_1", False), ("attach_fails_everywhere", False), ("attach_fails_on_1", "raised"))


def _attach_worker(rank, world, port_no, q):
    try:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port_no)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        out = []
        for scenario, _ in SCENARIOS:
            export_ok = not (scenario == "export_fails_on_1" and rank == 1)
            attach_ok = not (scenario == "attach_fails_on_1" and rank == 1) and scenario != "attach_fails_everywhere"
            g = _FakeGrid(export_ok, attach_ok)
            try:
                res = ring.attach_p2p(dist, g, rank, world)
            except RuntimeError:
                res = "raised"
            dist.barrier()
            out.append((scenario, res, g.attached))
        dist.destroy_process_group()
        q.put((rank, out))
    except Exception:   # pragma: no cover
        import traceback
        q.put((rank, "FAIL: " + traceback.format_exc()))


def test_transport_choice_is_collective():
    """ring.attach_p2p: every rank ends with the same answer -- peer stores everywhere, NCCL everywhere (and then NO rank has
    attached when the export failed), or an exception everywhere when the mappings opened on some ranks only"""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port_no = _free_port()
    procs = [ctx.Process(target=_attach_worker, args=(r, world, port_no, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert all(isinstance(r[1], list) for r in res), res
    for i, (scenario, expect) in enumerate(SCENARIOS):
        got = [r[1][i] for r in res]
        assert [g[1] for g in got] == [expect] * world, (scenario, got)
        if scenario == "export_fails_on_1":
            assert not any(g[2] for g in got), got
