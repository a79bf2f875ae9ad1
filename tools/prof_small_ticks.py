"""configs[0] as the launch list sees it: 1 000 bodies, recs_gpu_tick(100) twice (python tools/prof_small_ticks.py, under ncu)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from recs_b200 import GpuContext, scenes  # noqa: E402
from recs_b200.nbody import NBodyGpuApp  # noqa: E402

with GpuContext(device=0) as ctx:
    app = NBodyGpuApp(ctx, *scenes.spawn_bodies(scenes.EVERYTHING_HEAVY, 1))
    app.upload()
    app.tick(100)
    app.tick(100)
    ctx.synchronize()
