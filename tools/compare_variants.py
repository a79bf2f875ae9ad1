"""Bitwise comparison of the accelerations two gravity kernel variants produce (RECS_GPU_GRAVITY_VARIANT)
on the same bodies: used to show that the range-shifted variants equal the unshifted formulation bit for bit."""
import os
import sys

import numpy as np

sys.path.insert(0, ".")
from recs_b200 import GpuContext, scenes  # noqa: E402
from recs_b200.nbody import NBodyGpuApp  # noqa: E402

va, vb = int(sys.argv[1]), int(sys.argv[2])
sizes = [int(a) for a in sys.argv[3].split(",")] if len(sys.argv) > 3 else [4097, 65536]
for n in sizes:
    out = {}
    for v in (va, vb):
        os.environ["RECS_GPU_GRAVITY_VARIANT"] = str(v)
        with GpuContext(device=0, use_cuda_graph=False) as ctx:
            pos, mass, vel, acc = scenes.spawn_bodies(scenes.all_heavy_random_cube_with_bodies(n), 1)
            # a few coincident and sub-epsilon pairs so the exact path is exercised
            pos[7] = pos[3]
            pos[n // 2] = pos[n - 1] + np.float32(1e-4)
            app = NBodyGpuApp(ctx, pos, mass, vel, acc)
            app.upload()
            app.run("gravity", sync=True)
            app.download()
            out[v] = app.acc.copy()
    a, b = out[va].view(np.uint32), out[vb].view(np.uint32)
    diff = int((a != b).sum())
    rel = np.abs(out[va].astype(np.float64) - out[vb]).max() / np.abs(out[vb]).max()
    print(f"n={n}: v{va} vs v{vb}: {diff} of {a.size} words differ, max abs diff / max = {rel:.3e}, "
          f"finite={np.isfinite(out[va]).all() and np.isfinite(out[vb]).all()}", flush=True)
