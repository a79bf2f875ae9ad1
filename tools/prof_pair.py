"""The generic nested-query system (reference-order gravity, recs_b200/nbody.py) at N outer x N query entities:
timing, and a target for ncu.  usage: python tools/prof_pair.py [bodies] [reps]"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from recs_b200 import GpuContext, scenes  # noqa: E402
from recs_b200.nbody import ReferenceOrderNBodyApp  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
pos, mass, vel, acc = scenes.spawn_bodies(scenes.all_heavy_random_cube_with_bodies(n), 1)
with GpuContext(device=0, use_cuda_graph=False) as ctx:
    app = ReferenceOrderNBodyApp(ctx, pos, mass, vel, acc)
    app.upload()
    app.run("gravity")
    ctx.synchronize()
    ctx.timer_start()
    for _ in range(reps):
        app.run("gravity", sync=False)
    ms = ctx.timer_stop() / reps
print(f"generic reference-order gravity, {n} bodies: {ms:.3f} ms  {n * n / ms / 1e6:.1f} G interactions/s")
