"""Average tick duration at the reference's own benchmark sizes (2^0 .. 2^14 bodies, 100 ticks per
sample, crates/ecs_bench_suite/benches/n_body.rs:4,16-21): GPU chain (CUDA graph) vs the CPU port of
the reference's worker pool on this box."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import oracle  # noqa: E402  (CPU baseline)
from recs_b200 import GpuContext, scenes  # noqa: E402
from recs_b200.nbody import NBodyGpuApp  # noqa: E402

orc = oracle.load_nbody()
threads = len(os.sched_getaffinity(0))
print(f"# host threads {threads}; 100 ticks per sample, best of 3")
print("# bodies   gpu_us_per_tick  cpu_pool_us_per_tick   gpu G inter/s   cpu G inter/s")
for e in range(0, 15):
    n = 1 << e
    pos, mass, vel, acc = scenes.spawn_bodies(scenes.all_heavy_random_cube_with_bodies(n), 1)
    cpu = [a.copy() for a in (pos, mass, vel, acc)]
    with GpuContext(device=0, use_cuda_graph=True) as ctx:
        app = NBodyGpuApp(ctx, pos, mass, vel, acc)
        app.upload()
        app.tick(100)
        ctx.synchronize()
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            app.tick(100)
            ctx.synchronize()
            best = min(best, time.perf_counter() - t0)
    cbest = 1e9
    for _ in range(3 if n <= 4096 else 1):
        cbest = min(cbest, orc.run_pool(cpu[0], cpu[1], cpu[2], cpu[3], 100, threads))
    print(f"{n:8d} {best*1e4:16.2f} {cbest*1e4:20.2f} {n*n/(best/100)/1e9:14.3f} {n*n/(cbest/100)/1e9:14.3f}", flush=True)
