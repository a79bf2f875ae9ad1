"""Each streaming kernel of configs[3] / configs[4] three times (for `ncu --set full`: dram bytes per launch).
Launch order: axpy_rn_vec_kernel x3 (simple_iter 10 M), recs_generic_system x3 (the same as a generic system),
rain_tma_kernel<2> x3 at 4 Mi drops, rain_tma_kernel<2> x3 at 16 Mi drops."""
import sys

import numpy as np

sys.path.insert(0, ".")
from recs_b200 import GpuContext, _lib, scenes  # noqa: E402

with GpuContext(device=0, use_cuda_graph=False) as ctx:
    n = 10_000_000
    p = np.ones((n, 3), np.float32)
    v = np.ones((n, 3), np.float32)
    ctx.bind_column(7, 101, p)
    ctx.bind_column(7, 102, v)
    ctx.upload(7, 101)
    ctx.upload(7, 102)
    sid = ctx.create_system(_lib.SYS_QUERY_ADD, [101, 102])
    body = ("struct Position { float x, y, z; };\nstruct Velocity { float x, y, z; };\n"
            "void movement(Position& p, const Velocity& v) { p.x += v.x; p.y += v.y; p.z += v.z; }\n")
    gid = ctx.create_kernel_system("simple_iter", body, "movement", [(101, 12, "write", "Position"), (102, 12, "read", "Velocity")])
    for s in (sid, gid):
        ctx.set_archetypes(s, [7])
        for _ in range(3):
            ctx.run_system(s)
    for drops in (4 * 1024 * 1024, 16 * 1024 * 1024):
        vel, mass, pos = scenes.rain_drops(drops, 1)
        ctx.bind_column(8, 201, vel)
        ctx.bind_column(8, 202, pos)
        ctx.upload(8, 201)
        ctx.upload(8, 202)
        if drops == 4 * 1024 * 1024:
            w = ctx.create_system(_lib.SYS_RAIN_WIND, [201])
            g = ctx.create_system(_lib.SYS_RAIN_GRAVITY, [201])
            m = ctx.create_system(_lib.SYS_RAIN_MOVEMENT, [202, 201])
            for s in (w, g, m):
                ctx.set_archetypes(s, [8])
            ctx.set_tick_order([w, g, m])
        for _ in range(3):
            ctx.tick(1)
            ctx.synchronize()
print("ok")
