"""Times the `collision` system (crates/n-body/examples/n_body_demo.rs:54-71) alone and inside the demo's system set
through the Application (gravity, collision -> acceleration -> movement), at 16 384 and 65 536 bodies.
Reports G pair-tests/s (one pair test = one (entity, other entity) distance comparison, N*N per run when nothing exits
early) and the share of the FP32 issue rate: the loop issues 6 packed + 6.5 scalar instructions per two rows and body."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from recs_b200 import GpuContext, _lib, ecs, scenes  # noqa: E402

POS, MASS, VEL, ACC = 1, 2, 3, 4
F3 = np.dtype((np.float32, 3))
out = {}
for n in (16384, 65536):
    pos, mass, vel, acc = scenes.spawn_bodies(scenes.all_heavy_random_cube_with_bodies(n), 3)
    with GpuContext(device=0, use_cuda_graph=False) as ctx:
        ctx.bind_column(1, POS, pos)
        ctx.upload(1, POS)
        sid = ctx.create_system(_lib.SYS_NBODY_COLLISION, [POS])
        ctx.set_archetypes(sid, [1])
        ctx.set_query_archetypes(sid, [1])
        for _ in range(3):
            ctx.run_system(sid, sync=False)
        ctx.synchronize()
        reps = 20
        ctx.timer_start()
        for _ in range(reps):
            ctx.run_system(sid, sync=False)
        ms = ctx.timer_stop() / reps
        hits = len(ctx.system_removals(sid, 1))
        out[f"collision_alone_{n}"] = {"ms": ms, "G_pair_tests_per_s": n * n / ms / 1e6, "rows_flagged": hits}
    # the demo's system set through the Application (positions in a box where nothing collides, so no entity leaves)
    with GpuContext(device=0) as ctx:
        app = ecs.Application(ctx)
        w = app.world
        for comp, dt, name in ((POS, F3, "Position"), (MASS, np.float32, "Mass"), (VEL, F3, "Velocity"), (ACC, F3, "Acceleration")):
            w.register_component(comp, dt, name)
        app.add_gpu_system(_lib.SYS_NBODY_MOVEMENT, [POS, VEL], "movement")
        app.add_gpu_system(_lib.SYS_NBODY_ACCELERATION, [VEL, ACC], "acceleration")
        app.add_gpu_system(_lib.SYS_NBODY_GRAVITY, [POS, ACC, MASS], "gravity")
        app.add_gpu_system(_lib.SYS_NBODY_COLLISION, [POS], "collision")
        w.create_entities(n, {POS: pos * np.float32(50.0), MASS: mass, VEL: vel, ACC: acc})
        app.run(3)
        ctx.synchronize()
        t0 = time.perf_counter()
        ticks = 20
        app.run(ticks)
        ctx.synchronize()
        wall = (time.perf_counter() - t0) / ticks * 1e3
        order, _ = app.last_tick_order()
        out[f"demo_tick_{n}"] = {"ms_per_tick_wall": wall, "order": [app.system_names[i] for i in order],
                                 "removed_last_tick": app.last_playback()[1]}
        app.close()
print(json.dumps(out, indent=1))
