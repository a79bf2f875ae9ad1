"""Whole-tick time (CUDA graph of the chain) of the three problem-size shapes at the sizes between them.
python tools/sweep_mid_sizes.py [sizes] [variants]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from recs_b200 import GpuContext, scenes  # noqa: E402
from recs_b200.nbody import NBodyGpuApp  # noqa: E402

sizes = [int(a) for a in sys.argv[1].split(",")] if len(sys.argv) > 1 else [4097, 6144, 8192, 12288, 16384, 20480, 24576, 32768, 49152, 65536]
variants = sys.argv[2].split(",") if len(sys.argv) > 2 else ["auto", "2", "3", "0"]
os.environ["RECS_GPU_RESIDENT_MAX_ROWS"] = "0"
for n in sizes:
    cols = scenes.spawn_bodies(scenes.all_heavy_random_cube_with_bodies(n), 1)
    line = f"n={n:6d}"
    for v in variants:
        os.environ.pop("RECS_GPU_GRAVITY_VARIANT", None)
        if v != "auto":
            os.environ["RECS_GPU_GRAVITY_VARIANT"] = v
        with GpuContext(device=0, use_cuda_graph=True) as ctx:
            app = NBodyGpuApp(ctx, *[c.copy() for c in cols])
            app.upload()
            ticks = max(10, min(200, int(4e10 / (n * n))))
            app.tick(ticks)
            ctx.synchronize()
            best = 1e9
            for _ in range(3):
                ctx.timer_start()
                app.tick(ticks)
                best = min(best, ctx.timer_stop())
            us = best / ticks * 1e3
            line += f"  v{v}: {us:8.2f} us ({n * n / us / 1e3:6.0f} G/s)"
    print(line, flush=True)
