"""Index sharding of the body archetype over R ranks (one process per GPU).

Rank r owns rows [r*S, min((r+1)*S, N)).  S is a whole number of CHUNKS of the gravity kernel's summation tree (a chunk
= K = ceil(tiles / 32) tiles of 256 bodies, a function of N alone): every rank's slice of the pair-interleaved
position mirror is one contiguous, equally sized run of whole chunks, so a rank's own tiles and the other ranks' tiles
are two runs of whole chunks and a row's result does not depend on the rank count (recs_b200/csrc/kernels.cuh).  The
partition is computed by librecs_gpu.so (recs_gpu_shard_plan); this module only wraps it for host-side tests:

    tick:  gravity(own rows x all bodies)  ->  movement(own rows)  ->  acceleration(own rows)
           -> the refreshed positions of the own rows reach every rank (peer push by the integration tail, or the
              NCCL all-gather) before the next gravity reads them
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

from ._lib import lib


@dataclass(frozen=True)
class ShardPlan:
    n_rows: int
    rank: int
    world_size: int
    row_begin: int
    row_end: int
    slice_rows: int  # rows per rank in the padded gather buffer

    @property
    def padded_rows(self) -> int:
        return self.slice_rows * self.world_size

    @property
    def gather_bytes_per_rank(self) -> int:
        return self.slice_rows * 16  # {x, y, z, G*m} per body


def shard_plan(n_rows: int, rank: int, world_size: int) -> ShardPlan:
    b, e, s = C.c_uint64(), C.c_uint64(), C.c_uint64()
    code = lib.recs_gpu_shard_plan(n_rows, rank, world_size, C.byref(b), C.byref(e), C.byref(s))
    if code:
        raise ValueError(f"recs_gpu_shard_plan({n_rows}, {rank}, {world_size}) failed with {code}")
    return ShardPlan(n_rows, rank, world_size, int(b.value), int(e.value), int(s.value))
