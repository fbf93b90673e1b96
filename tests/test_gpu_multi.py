"""Multi-GPU parity (needs >= 2 GPUs, skipped otherwise): x-slabs over NCCL must reproduce the serial
reference bit-for-bit on every rank's planes."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("world", [2, 4, 8])
def test_slabs_match_serial_oracle(world):
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    # cases whose x extent leaves >= 4 planes per rank and whose BC extrapolation stays inside a slab
    names = "chan3d,cyl3d,chan2d,cyl2d,slipchan3d,sliptunnel2d,sliptunnel3d,fevel2d,fevel3d,fevel2d_tav,tunnel2d_tav"
    if world <= 4:
        names += ",cav3d_32,cav3d_tav,felid3d,kbc2d_cyl,kbc3d_chan"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(29610 + world), os.path.join(HERE, "mgpu_worker.py"), names]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=1500)
    out = r.stdout.decode()
    assert r.returncode == 0, out[-4000:]
    assert out.count("mgpu ok") == len(names.split(",")), out[-4000:]


@pytest.mark.parametrize("workload,steps", [("c4", 20), ("c3", 12)])
def test_full_size_slabs_equal_one_gpu(workload, steps):
    """BASELINE configs[2] / configs[3] at full size: the x-slab run on all GPUs of the box must leave the same bits
    as the single-GPU run (compared through order-independent 64-bit wrap-sums and xors of f, rho, u)."""
    world = torch.cuda.device_count()
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    sums = {}
    for n in (1, world):
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
               "--master-addr", "127.0.0.1", "--master-port", str(29640 + n), os.path.join(HERE, "mgpu_worker.py"),
               "fullsize:%s:%d" % (workload, steps)]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=1500)
        out = r.stdout.decode()
        assert r.returncode == 0, out[-4000:]
        line = [l for l in out.splitlines() if l.startswith("fullsize checksums")]
        assert len(line) == 1, out[-4000:]
        sums[n] = line[0].split(":", 1)[1].strip()
    assert sums[1] == sums[world], sums
