"""Quick on-box probe: FP32 pipe rate (scalar vs packed), gravity timing, streaming bandwidth."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from recs_b200 import GpuContext, scenes  # noqa: E402
from recs_b200.nbody import NBodyGpuApp  # noqa: E402

out = {}
with GpuContext(device=0, use_cuda_graph=False) as ctx:
    for packed in (False, True):
        tf, mhz = ctx.measure_fp32_peak(packed)
        out["fp32_peak_packed" if packed else "fp32_peak_scalar"] = {"tflops": tf, "sm_mhz": mhz}
    for n, steps in ((65536, 20), (262144, 5), (1048576, 2)):
        pos, mass, vel, acc = scenes.spawn_bodies(scenes.all_heavy_random_cube_with_bodies(n), 1)
        app = NBodyGpuApp(ctx, pos, mass, vel, acc)
        app.upload()
        app.tick(2)
        ctx.synchronize()
        ctx.enable_timing(True)
        ctx.reset_timing()
        t0 = time.time()
        app.tick(steps)
        ctx.synchronize()
        wall = time.time() - t0
        t = ctx.timing()
        ctx.enable_timing(False)
        g_ms = t["gravity_ms"] / steps
        out[f"nbody_{n}"] = {
            "wall_ms_per_tick": wall * 1e3 / steps,
            "gravity_ms": g_ms,
            "reduce_ms": t["gravity_reduce_ms"] / steps,
            "movement_ms": t["movement_ms"] / steps,
            "accel_ms": t["acceleration_ms"] / steps,
            "pack_ms": t["pack_ms"] / steps,
            "ginter_per_s": n * n / (g_ms * 1e-3) / 1e9,
            "tflops20": 20.0 * n * n / (g_ms * 1e-3) / 1e12,
            "movement_GBps": 36.0 * n / (t["movement_ms"] / steps * 1e-3) / 1e9,
        }
print(json.dumps(out, indent=1))
