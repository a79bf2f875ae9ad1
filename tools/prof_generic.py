"""Generic (NVRTC-generated) systems against the hand-written streaming kernel on simple_iter, 10 M entities,
and frag_iter (26 archetypes x 20 entities) launch cost.  RECS_GPU_KERNEL_DIRECT=1 selects the unstaged form."""
import json
import sys

import numpy as np

sys.path.insert(0, ".")
from recs_b200 import GpuContext, _lib  # noqa: E402

SRC = ("struct Position { float x, y, z; };\nstruct Velocity { float x, y, z; };\n"
       "void movement(Position& p, const Velocity& v) { p.x += v.x; p.y += v.y; p.z += v.z; }\n")
FRAG = "struct Data { float v; };\nvoid doubling(Data& d) { d.v *= 2.0f; }\n"
out = {}
with GpuContext(device=0, use_cuda_graph=False) as ctx:
    n = 10_000_000
    p = np.ones((n, 3), np.float32)
    v = np.ones((n, 3), np.float32)
    ctx.bind_column(1, 1, p)
    ctx.bind_column(1, 2, v)
    ctx.upload(1, 1)
    ctx.upload(1, 2)
    hand = ctx.create_system(_lib.SYS_QUERY_ADD, [1, 2])
    gen = ctx.create_kernel_system("movement", SRC, "movement", [(1, 12, "write", "Position"), (2, 12, "read", "Velocity")])
    for name, sid in (("hand_written", hand), ("generic", gen)):
        ctx.set_archetypes(sid, [1])
        for _ in range(3):
            ctx.run_system(sid, sync=False)
        ctx.synchronize()
        ctx.timer_start()
        for _ in range(20):
            ctx.run_system(sid, sync=False)
        ms = ctx.timer_stop() / 20
        out[f"simple_iter_10M_{name}"] = {"ms": ms, "GBps": 36.0 * n / (ms * 1e-3) / 1e9}
    cols = []
    for a in range(26):
        c = np.ones((20, 1), np.float32)
        cols.append(c)
        ctx.bind_column(10 + a, 5, c)
        ctx.upload(10 + a, 5)
    frag = ctx.create_kernel_system("doubling", FRAG, "doubling", [(5, 4, "write", "Data")])
    ctx.set_archetypes(frag, list(range(10, 36)))
    for _ in range(3):
        ctx.run_system(frag, sync=False)
    ctx.synchronize()
    ctx.timer_start()
    for _ in range(100):
        ctx.run_system(frag, sync=False)
    out["frag_iter_26x20_us_per_run"] = ctx.timer_stop() * 10.0
    # gravity written as a generic nested-query system (reference order, IEEE div + sqrt) against the hand-written kernel
    sys.path.insert(0, "tests")
    from test_pair_kernel_systems import GRAVITY, GRAVITY_PARAMS  # noqa: E402
    from recs_b200 import scenes  # noqa: E402

    for n in (16384, 65536):
        pos, mass, vel, acc = scenes.spawn_bodies(scenes.all_heavy_random_cube_with_bodies(n), 1)
        arch = 100 + n
        for comp, arr in ((1, pos), (2, mass), (4, acc)):
            ctx.bind_column(arch, comp, arr)
            ctx.upload(arch, comp)
        gsys = ctx.create_pair_kernel_system("gravity", GRAVITY, "Total", "pair", "finish", GRAVITY_PARAMS)
        hsys = ctx.create_system(_lib.SYS_NBODY_GRAVITY, [1, 4, 2])
        for sid in (gsys, hsys):
            ctx.set_archetypes(sid, [arch])
            ctx.set_query_archetypes(sid, [arch])
        for name, sid in (("generic_reference_order", gsys), ("hand_written", hsys)):
            ctx.run_system(sid)
            ctx.timer_start()
            for _ in range(3):
                ctx.run_system(sid, sync=False)
            ms = ctx.timer_stop() / 3
            out[f"gravity_{n}_{name}"] = {"ms": ms, "G_interactions_per_s": n * n / ms / 1e6}
print(json.dumps(out))
