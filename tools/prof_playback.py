"""Cost of one structural change on resident columns, 4 Mi raindrops {Velocity, Position}: k rows removed (swap-remove
replay as row moves) and k rows appended, against the download + re-upload round trip it replaces."""
import ctypes as C
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from recs_b200 import GpuContext, _lib  # noqa: E402
from recs_b200._lib import lib  # noqa: E402

n = 4 * 1024 * 1024
VEL, POS, ARCH = 1, 2, 3
out = {}
with GpuContext(device=0, use_cuda_graph=False) as ctx:
    cap = n + (1 << 20)
    vel = np.zeros((cap, 3), np.float32)
    pos = np.random.default_rng(1).random((cap, 3), dtype=np.float32)
    ctx.bind_column(ARCH, VEL, vel[:n])
    ctx.bind_column(ARCH, POS, pos[:n])
    ctx.upload(ARCH, VEL)
    ctx.upload(ARCH, POS)
    ctx.synchronize()
    comps = (C.c_uint64 * 2)(VEL, POS)
    for k in (1000, 40000):
        rng = np.random.default_rng(k)
        times = []
        count = n
        for rep in range(6):
            # append k rows, then remove k random rows (fold of swap-removes: distinct dst < new count, src from the tail)
            t0 = time.perf_counter()
            for comp, arr in ((VEL, vel), (POS, pos)):
                st = lib.recs_gpu_column_append(ctx._ctx, ARCH, comp, arr.ctypes.data_as(C.c_void_p), count + k)
                assert st == 0, ctx.last_error()
            dst = np.sort(rng.choice(count, size=k, replace=False)).astype(np.uint32)
            src = np.arange(count, count + k, dtype=np.uint32)
            st = lib.recs_gpu_archetype_move_rows(
                ctx._ctx, ARCH, comps, 2, dst.ctypes.data_as(C.POINTER(C.c_uint32)), src.ctypes.data_as(C.POINTER(C.c_uint32)), k, count)
            assert st == 0, ctx.last_error()
            ctx.synchronize()
            times.append((time.perf_counter() - t0) * 1e3)
        out[f"resident_append_and_remove_{k}_rows_ms"] = float(np.median(times[1:]))
    times = []
    for rep in range(4):
        t0 = time.perf_counter()
        ctx.download(ARCH, VEL)
        ctx.download(ARCH, POS)
        ctx.bind_column(ARCH, VEL, vel[:n])
        ctx.bind_column(ARCH, POS, pos[:n])
        ctx.upload(ARCH, VEL)
        ctx.upload(ARCH, POS)
        ctx.synchronize()
        times.append((time.perf_counter() - t0) * 1e3)
    out["round_trip_download_rebind_upload_ms"] = float(np.median(times[1:]))
print(json.dumps(out))
