"""The reference's rain_simulation benchmark (crates/ecs_bench_suite/src/recs/rain_simulation.rs: all seven systems, clouds
from create_evenly_interspersed_clouds, >= 100 ticks) with the WHOLE tick on the device: the streaming systems, `rain`
(Commands::create) and `vapor_accumulation` as generic systems, ground_collision deciding removals; the host World plays
the commands back every tick and the resident columns follow by appends / row moves.  Wall-clock per tick (host + device).
python tools/bench_rain_app.py [clouds,...] [ticks]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from recs_b200 import GpuContext, _lib, ecs  # noqa: E402

F3 = np.dtype((np.float32, 3))
VEL, POS, MASS, CLOUD = 20, 21, 22, 23
SOURCE = """
struct Cloud { float vapor_accumulation_rate, width, height; };
struct Position { float x, y, z; };
struct Mass { float m; };
struct Drop { float vx, vy, vz; float mass; float px, py, pz; };
void rain(const Cloud& cloud, const Position& position, Mass& mass, const recs_commands& commands) {
    while (mass.m > 0.1f) {
        commands.create(Drop{0.0f, 0.0f, 0.0f, 0.1f, position.x, position.y, position.z});
        mass.m -= 0.1f;
    }
}
void vapor_accumulation(const Cloud& cloud, Mass& mass) { mass.m += cloud.vapor_accumulation_rate * 0.05f; }
"""

sizes = [int(a) for a in sys.argv[1].split(",")] if len(sys.argv) > 1 else [1024, 16384]
ticks = int(sys.argv[2]) if len(sys.argv) > 2 else 100
for n_clouds in sizes:
    with GpuContext(device=0) as ctx:
        app = ecs.Application(gpu=ctx)
        w = app.world
        for c, dt, name in ((VEL, F3, "Velocity"), (POS, F3, "Position"), (MASS, np.float32, "Mass"), (CLOUD, F3, "Cloud")):
            w.register_component(c, dt, name)
        app.add_gpu_system(_lib.SYS_RAIN_GRAVITY, [VEL], "gravity")
        app.add_gpu_system(_lib.SYS_RAIN_WIND, [VEL], "wind")
        app.add_gpu_system(_lib.SYS_RAIN_WIND_DIRECTION, [], "wind_direction")
        app.add_gpu_system(_lib.SYS_RAIN_MOVEMENT, [POS, VEL], "movement")
        app.add_gpu_kernel_system("rain", SOURCE, "rain", [("read", CLOUD), ("read", POS), ("write", MASS), ("commands",)],
                                  ["Cloud", "Position", "Mass"], creates=[VEL, MASS, POS], create_record_type="Drop")
        app.add_gpu_system(_lib.SYS_RAIN_GROUND_COLLISION, [POS], "ground_collision")
        app.add_gpu_kernel_system("vapor_accumulation", SOURCE, "vapor_accumulation", [("read", CLOUD), ("write", MASS)], ["Cloud", "Mass"])
        side = int(np.sqrt(np.float32(n_clouds)))
        xs, zs = np.meshgrid(np.arange(side, dtype=np.float32), np.arange(side, dtype=np.float32), indexing="ij")
        pos = np.stack([xs.ravel() * 10, np.full(side * side, 10.0, np.float32), zs.ravel() * 10], axis=1).astype(np.float32)
        cloud = np.tile(np.array([1.0, 10.0, 10.0], np.float32), (side * side, 1))
        w.create_entities(side * side, {CLOUD: cloud, POS: pos, MASS: np.zeros(side * side, np.float32)})
        app.run(100)  # reach the steady state: drops are landing
        ctx.synchronize()
        t0 = time.perf_counter()
        created = removed = 0
        for _ in range(ticks):
            app.run(1)
            c, r = app.last_playback()
            created += c
            removed += r
        ctx.synchronize()
        sec = time.perf_counter() - t0
        drops = len(w.archetype_entities(2)) if w.archetype_count() > 2 else 0
        print(f"clouds={side * side:6d}: {sec / ticks * 1e3:8.3f} ms per tick over {ticks} ticks (wall clock, host playback included); "
              f"{drops} drops alive, {created / ticks:.0f} created and {removed / ticks:.0f} removed per tick", flush=True)
        app.close()
