"""A few eager ticks at a mid size, for an ncu launch list: python tools/prof_mid_ticks.py [bodies] [ticks]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from recs_b200 import GpuContext, scenes  # noqa: E402
from recs_b200.nbody import NBodyGpuApp  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
ticks = int(sys.argv[2]) if len(sys.argv) > 2 else 4
with GpuContext(device=0, use_cuda_graph=False) as ctx:
    app = NBodyGpuApp(ctx, *scenes.spawn_bodies(scenes.all_heavy_random_cube_with_bodies(n), 1))
    app.upload()
    for _ in range(ticks):
        app.tick(1)
    ctx.synchronize()
