"""Fused rain tick (wind -> gravity -> movement, 48 B per drop) at several drop counts, single ticks after an L2 flush and
ticks back to back.  4 Mi drops (2 x 50 MB columns) fit the 126 MB L2; 16 Mi and 32 Mi do not, so they show the
HBM-bound rate.  RECS_GPU_RAIN_STAGED=1 selects the LDG/STS staged kernel instead of the TMA pipeline."""
import json
import sys

import numpy as np

sys.path.insert(0, ".")
from recs_b200 import GpuContext, _lib, scenes  # noqa: E402

out = {}
for n_mi in (4, 16, 32):
    n = n_mi * 1024 * 1024
    with GpuContext(device=0, use_cuda_graph=False) as ctx:
        VEL, POS, ARCH = 201, 202, 8
        vel, mass, pos = scenes.rain_drops(n, 1)
        ctx.bind_column(ARCH, VEL, vel)
        ctx.bind_column(ARCH, POS, pos)
        ctx.upload(ARCH, VEL)
        ctx.upload(ARCH, POS)
        w = ctx.create_system(_lib.SYS_RAIN_WIND, [VEL])
        g = ctx.create_system(_lib.SYS_RAIN_GRAVITY, [VEL])
        m = ctx.create_system(_lib.SYS_RAIN_MOVEMENT, [POS, VEL])
        for s in (w, g, m):
            ctx.set_archetypes(s, [ARCH])
        ctx.set_tick_order([w, g, m])
        for _ in range(3):
            ctx.tick(1)
        ms = []
        for _ in range(10):
            ctx.flush_l2()
            ctx.timer_start()
            ctx.tick(1)
            ms.append(ctx.timer_stop())
        ctx.synchronize()
        ctx.timer_start()
        for _ in range(20):
            ctx.tick(1)
        b2b = ctx.timer_stop() / 20
        out[f"{n_mi}Mi"] = {"flushed_ms": float(np.median(ms)), "flushed_GBps": 48.0 * n / (np.median(ms) * 1e-3) / 1e9,
                            "back_to_back_ms": b2b, "back_to_back_GBps": 48.0 * n / (b2b * 1e-3) / 1e9}
print(json.dumps(out))
