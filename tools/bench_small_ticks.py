"""Ticks per second of small n-body worlds: recs_gpu_tick(n) as one cooperative launch (nbody_resident_ticks_kernel)
against the tick-by-tick chain (CUDA graph of the two launches).  python tools/bench_small_ticks.py [sizes]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from recs_b200 import GpuContext, scenes  # noqa: E402
from recs_b200.nbody import NBodyGpuApp  # noqa: E402

sizes = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [256, 1000, 2048, 3000, 4096]
ticks = 1000
for n in sizes:
    row = []
    for max_rows in ("4096", "0"):
        os.environ["RECS_GPU_RESIDENT_MAX_ROWS"] = max_rows
        with GpuContext(device=0, use_cuda_graph=True) as ctx:
            cols = scenes.spawn_bodies(scenes.all_heavy_random_cube_with_bodies(n), 1)
            app = NBodyGpuApp(ctx, *cols)
            app.upload()
            app.tick(ticks)
            ctx.synchronize()
            best = 1e9
            for _ in range(5):
                ctx.timer_start()
                app.tick(ticks)
                best = min(best, ctx.timer_stop())
            row.append(best / ticks * 1e3)
    print(f"n={n:5d}  resident {row[0]:7.2f} us/tick   chain {row[1]:7.2f} us/tick   x{row[1] / row[0]:.2f}   "
          f"({n * n / row[0] / 1e3:.1f} vs {n * n / row[1] / 1e3:.1f} G interactions/s)", flush=True)
