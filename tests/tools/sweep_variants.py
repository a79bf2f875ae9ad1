"""Times the gravity kernel variants (RECS_GPU_GRAVITY_VARIANT) on one GPU."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, ".")
from recs_b200 import GpuContext, scenes  # noqa: E402
from recs_b200.nbody import NBodyGpuApp  # noqa: E402
import oracle  # noqa: E402  (checker only)

sizes = [int(a) for a in sys.argv[1].split(",")] if len(sys.argv) > 1 else [65536, 262144]
variants = [int(a) for a in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 1, 2, 3]
orc = oracle.load_nbody()
res = {}
for v in variants:
    os.environ["RECS_GPU_GRAVITY_VARIANT"] = str(v)
    with GpuContext(device=0, use_cuda_graph=False) as ctx:
        for n in sizes:
            pos, mass, vel, acc = scenes.spawn_bodies(scenes.all_heavy_random_cube_with_bodies(n), 1)
            app = NBodyGpuApp(ctx, pos, mass, vel, acc)
            app.upload()
            for _ in range(3):
                app.run("gravity", sync=False)
            ctx.synchronize()
            reps = max(3, int(2e11 / (n * n)))
            ctx.enable_timing(True)
            ctx.reset_timing()
            for _ in range(reps):
                app.run("gravity", sync=False)
            t = ctx.timing()
            ctx.enable_timing(False)
            ms = t["gravity_ms"] / reps
            app.download()
            rows = np.random.default_rng(3).choice(n, size=32, replace=False)
            ref32, _ = orc.gravity_sample(pos, mass, rows, f64=False)
            err = np.linalg.norm(app.acc[rows].astype(np.float64) - ref32, axis=1) / np.linalg.norm(ref32, axis=1)
            res[f"v{v}_n{n}"] = {"ms": ms, "ginter_s": n * n / ms / 1e6, "max_rel_err": float(err.max())}
            print(f"v{v} n={n}: {ms:.3f} ms  {n*n/ms/1e6:.1f} Ginter/s  err {err.max():.2e}", flush=True)
print(json.dumps(res))
