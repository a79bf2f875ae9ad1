"""Short gravity-only run for ncu (N bodies, a few launches)."""
import sys

sys.path.insert(0, ".")
from recs_b200 import GpuContext, scenes  # noqa: E402
from recs_b200.nbody import NBodyGpuApp  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
with GpuContext(device=0, use_cuda_graph=False) as ctx:
    pos, mass, vel, acc = scenes.spawn_bodies(scenes.all_heavy_random_cube_with_bodies(n), 1)
    app = NBodyGpuApp(ctx, pos, mass, vel, acc)
    app.upload()
    for _ in range(reps):
        app.run("gravity")
    ctx.synchronize()
print("ok")
