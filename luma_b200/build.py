"""In-tree build of libluma_b200.so (hand-written CUDA for sm_100a + the C ABI).

nvcc cross-compiles without a GPU; the built library sits next to this file so that it travels
with the repository snapshot to the GPU box (it is git-ignored, not gpurun-ignored).
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.environ.get("LUMA_B200_LIB") or os.path.join(HERE, "libluma_b200.so")
STAMP = LIB + ".srchash"
SOURCES = ["kernels_d3q19.cu", "kernels_d2q9.cu", "kernels_d3q27.cu", "kernels_common.cu", "api.cu"]
HEADERS = ["lattice.cuh", "kernels.cuh", "kernels_impl.cuh", os.path.join("..", "..", "include", "luma_b200.h")]

# -fmad=false: the reference is built without FMA contraction (makefile CFLAGS: -O3 -std=c++0x);
# parity is bit-for-bit, so the only fused operations are the explicit fma() calls in lattice.cuh.
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false", "-Xcompiler", "-fPIC", "-Xcompiler", "-O2",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _src_hash() -> str:
    h = hashlib.sha256()
    for nm in SOURCES + HEADERS:
        with open(os.path.join(CSRC, nm), "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_current() -> bool:
    if os.environ.get("LUMA_B200_LIB"):
        return os.path.exists(LIB)
    if not (os.path.exists(LIB) and os.path.exists(STAMP)):
        return False
    with open(STAMP) as fh:
        return fh.read().strip() == _src_hash()


def build_variant(tag: str, defines) -> str:
    """Development aid: an extra library whose D3Q19 kernels are built with -D tuning knobs (kernels_impl.cuh), e.g.
    build_variant("t256b3", ["-DLUMA_STEP_THREADS=256", "-DLUMA_MIN_BLOCKS=3"]) -> luma_b200/libluma_b200_t256b3.so
    (select it with LUMA_B200_LIB).  The other translation units are the default build's objects."""
    build()
    nvcc = _nvcc()
    bdir = os.path.join(HERE, "build")
    out = os.path.join(HERE, "libluma_b200_%s.so" % tag)
    obj = os.path.join(bdir, "kernels_d3q19_%s.o" % tag)
    subprocess.run([nvcc] + NVCC_FLAGS + list(defines) + ["-I", "/usr/include", "-c", os.path.join(CSRC, "kernels_d3q19.cu"), "-o", obj], check=True)
    objs = [os.path.join(bdir, s.replace(".cu", ".o")) for s in SOURCES if s != "kernels_d3q19.cu"] + [obj]
    subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", out] + objs + ["-ldl"], check=True)
    return out


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the library if the sources changed; returns its path."""
    if os.environ.get("LUMA_B200_LIB"):
        return LIB
    if not force and is_current():
        return LIB
    nvcc = _nvcc()
    inc = []
    for cand in ("/usr/include", os.path.join(os.path.dirname(os.path.dirname(nvcc)), "include")):
        if os.path.exists(os.path.join(cand, "nccl.h")):
            inc = ["-I", cand]
            break
    objs = []
    bdir = os.path.join(HERE, "build")
    os.makedirs(bdir, exist_ok=True)
    procs = []
    for src in SOURCES:
        obj = os.path.join(bdir, src.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + inc + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
        objs.append(obj)
    for cmd, pr in procs:
        out = pr.communicate()[0].decode()
        if pr.returncode != 0:
            raise RuntimeError("nvcc failed: %s\n%s" % (" ".join(cmd), out))
        if verbose:
            print(out)
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] + objs + ["-ldl"]
    subprocess.run(cmd, check=True)
    with open(STAMP, "w") as fh:
        fh.write(_src_hash() + "\n")
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
