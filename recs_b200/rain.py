"""The reference's rain simulation assembled on the C ABI with the WHOLE tick on the device
(crates/ecs_bench_suite/src/recs/rain_simulation.rs:56-65 registers gravity, wind, wind_direction, movement, rain,
ground_collision, vapor_accumulation; crates/rain-simulation/src/lib.rs:62-160 defines them).  ctypes plumbing only: the
streaming systems are built-in kinds, `rain` (Commands::create) and `vapor_accumulation` are generic systems whose bodies
are the two functions below, `ground_collision` decides removals; the host World plays the commands back every tick."""
from __future__ import annotations

import numpy as np

from . import _lib, ecs
from .gpu import GpuContext

F3 = np.dtype((np.float32, 3))
VELOCITY, POSITION, MASS, CLOUD = 20, 21, 22, 23

# crates/rain-simulation/src/lib.rs:128-148 (deterministic_point_on_surface = the cloud's centre, :45-47) and :158-160
RAIN_SYSTEMS_SOURCE = """
struct Cloud { float vapor_accumulation_rate, width, depth; };
struct Position { float x, y, z; };
struct Mass { float m; };
struct Drop { float vx, vy, vz; float mass; float px, py, pz; };   // (Velocity::default(), Mass(0.1), Position { drop_position })
void rain(const Cloud& cloud, const Position& position, Mass& mass, const recs_commands& commands) {
    while (mass.m > 0.1f) {
        commands.create(Drop{0.0f, 0.0f, 0.0f, 0.1f, position.x, position.y, position.z});
        mass.m -= 0.1f;
    }
}
void vapor_accumulation(const Cloud& cloud, Mass& mass) { mass.m += cloud.vapor_accumulation_rate * 0.05f; }
"""


def evenly_interspersed_clouds(max_clouds: int):
    """`create_evenly_interspersed_clouds` (crates/rain-simulation/src/scene.rs:27-54): a square grid of clouds 10 m apart
    at height 10, vapor accumulation rate 1, emit area 10 x 10.  Returns (cloud[N,3], position[N,3], mass[N])."""
    side = int(np.sqrt(np.float32(max_clouds)))
    xs, zs = np.meshgrid(np.arange(side, dtype=np.float32), np.arange(side, dtype=np.float32), indexing="ij")
    n = side * side
    pos = np.stack([xs.ravel() * np.float32(10), np.full(n, 10.0, np.float32), zs.ravel() * np.float32(10)], axis=1).astype(np.float32)
    cloud = np.tile(np.array([1.0, 10.0, 10.0], np.float32), (n, 1))
    return cloud, pos, np.zeros(n, np.float32)


def rain_application(ctx: GpuContext, max_clouds: int) -> ecs.Application:
    """The benchmark's application: every system a GPU system, clouds created, nothing run yet."""
    app = ecs.Application(gpu=ctx)
    w = app.world
    for comp, dt, name in ((VELOCITY, F3, "Velocity"), (POSITION, F3, "Position"), (MASS, np.float32, "Mass"), (CLOUD, F3, "Cloud")):
        w.register_component(comp, dt, name)
    app.add_gpu_system(_lib.SYS_RAIN_GRAVITY, [VELOCITY], "gravity")
    app.add_gpu_system(_lib.SYS_RAIN_WIND, [VELOCITY], "wind")
    app.add_gpu_system(_lib.SYS_RAIN_WIND_DIRECTION, [], "wind_direction")
    app.add_gpu_system(_lib.SYS_RAIN_MOVEMENT, [POSITION, VELOCITY], "movement")
    app.add_gpu_kernel_system("rain", RAIN_SYSTEMS_SOURCE, "rain",
                              [("read", CLOUD), ("read", POSITION), ("write", MASS), ("commands",)], ["Cloud", "Position", "Mass"],
                              creates=[VELOCITY, MASS, POSITION], create_record_type="Drop")
    app.add_gpu_system(_lib.SYS_RAIN_GROUND_COLLISION, [POSITION], "ground_collision")
    app.add_gpu_kernel_system("vapor_accumulation", RAIN_SYSTEMS_SOURCE, "vapor_accumulation", [("read", CLOUD), ("write", MASS)],
                              ["Cloud", "Mass"])
    cloud, pos, mass = evenly_interspersed_clouds(max_clouds)
    w.create_entities(cloud.shape[0], {CLOUD: cloud, POSITION: pos, MASS: mass})
    return app
