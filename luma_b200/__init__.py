"""luma_b200 -- B200-native (sm_100a CUDA + NCCL) implementation of LUMA's level-0 LBM time step,
GridObj::LBM_multi_opt, behind the C ABI of include/luma_b200.h.

Python here is plumbing and the host-side mirror of the reference's interface (Definitions ~
inc/definitions.h, GridObj ~ inc/GridObj.h); all lattice arithmetic is in luma_b200/csrc/*.cu.
There is no CPU fallback: without the built CUDA library, or without a GPU, calls fail loudly.
"""
from .definitions import Definitions, eExtrapolateRight, eFluid, ePressure, eSlip, eSolid, eVelocity  # noqa: F401
from .gridobj import GridObj, comm_unique_id  # noqa: F401
from . import capi  # noqa: F401


def kernel_fingerprint() -> str:
    """sha256 over the CUDA sources and nvcc flags the loaded library was built from: profiles/ncu_summary.json is
    stamped with it, so that bench.py quotes an ncu DRAM-traffic figure only for the kernel build it was measured on."""
    from . import build
    return build._src_hash()

__all__ = ["Definitions", "GridObj", "comm_unique_id", "capi", "kernel_fingerprint", "eSolid", "eFluid", "eVelocity", "ePressure", "eSlip",
           "eExtrapolateRight"]
