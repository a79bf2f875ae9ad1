"""Where a tick's time goes beside the interaction kernel: python tools/prof_tail.py [bodies]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from recs_b200 import GpuContext, scenes  # noqa: E402
from recs_b200.nbody import NBodyGpuApp  # noqa: E402

for n in [int(a) for a in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["131072"])]:
    with GpuContext(device=0, use_cuda_graph=False) as ctx:
        app = NBodyGpuApp(ctx, *scenes.spawn_bodies(scenes.all_heavy_random_cube_with_bodies(n), 1))
        app.upload()
        for _ in range(3):
            app.tick(1)
        ms = []
        for _ in range(10):
            ctx.flush_l2()
            ctx.timer_start()
            app.tick(1)
            ms.append(ctx.timer_stop())
        ctx.enable_timing(True)
        ctx.reset_timing()
        for _ in range(10):
            ctx.flush_l2()
            app.tick(1)
        t = ctx.timing()
        ctx.enable_timing(False)
        print(f"n={n}: tick {np.mean(ms) * 1e3:.1f} us (min {np.min(ms) * 1e3:.1f}); interaction kernel {t['gravity_ms'] * 100:.1f} us, "
              f"chunk sums + reduce + tail {t['gravity_reduce_ms'] * 100:.1f} us, launches/tick {t['kernel_launches'] / 10}")
