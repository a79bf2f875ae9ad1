"""recs_b200 -- the B200 (sm_100a) data-parallel hot path of the recs ECS engine.

The compute lives in librecs_gpu.so (hand-written CUDA behind the C ABI of include/recs_gpu.h and
include/recs_ecs.h); this package is the ctypes plumbing plus the host-side mirror of the
reference's workload set-up.  Importing it without the built library raises: there is no CPU,
PyTorch or oracle fallback.
"""
from . import _lib  # noqa: F401  (raises RecsLibraryMissing if librecs_gpu.so was not built)
from .gpu import GpuContext, RecsGpuError  # noqa: F401

__all__ = ["GpuContext", "RecsGpuError"]
