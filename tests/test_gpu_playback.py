"""Removal-deciding GPU systems + command playback against the host World:
the n-body demo's `collision` system (crates/n-body/examples/n_body_demo.rs:54-71) and the complete
rain-simulation tick (crates/rain-simulation/src/lib.rs:62-160, registration order of
crates/ecs_bench_suite/src/recs/rain_simulation.rs:56-65), compared with the Python/C oracle."""
import ctypes as C

import numpy as np
import pytest

from oracle import ecs_oracle as O
from recs_b200 import _lib, ecs
from tests.util import rel_err_rows

pytestmark = pytest.mark.gpu
F3 = np.dtype((np.float32, 3))
f32 = np.float32


def _arr(ptr, count, width):
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_float)), shape=(count, width) if width > 1 else (count,))


# ---------------------------------------------------------------------------------------------------
# collision kernel alone
# ---------------------------------------------------------------------------------------------------
def _collision_mask(pos, radius=f32(0.1), qpos=None, same=True):
    """cgmath's `distance2` = (dx*dx + dy*dy) + dz*dz, every product and sum rounded on its own; in row blocks."""
    qpos = pos if qpos is None else qpos
    out = np.zeros(len(pos), bool)
    for b in range(0, len(pos), 1024):
        d = qpos[None, :, :] - pos[b:b + 1024, None, :]  # other - self, f32
        d2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]
        if same:
            idx = np.arange(b, min(b + 1024, len(pos)))
            d2[idx - b, idx] = np.inf
        out[b:b + 1024] = (d2 < radius * radius).any(axis=1)
    return out


@pytest.mark.parametrize("n,box", [(1, 1.5), (2, 1.5), (257, 1.5), (1500, 1.5), (1024, 2.0), (9001, 4.0), (30000, 6.0)])
def test_collision_flags_match_oracle_bit_exact(gpu_ctx, n, box):
    """Sizes on both sides of the kernel's tile (256), CTA (1024 rows) and query-split boundaries; densities from
    "nearly everything collides" (the warp-level early exit) to a few per cent."""
    POS = 1
    rng = np.random.default_rng(n)
    pos = (rng.random((n, 3), dtype=np.float32) * f32(box)).astype(np.float32)
    if n >= 2:
        pos[1] = pos[0]  # coincident distinct entities do collide (only the entity itself is skipped)
    if n >= 2000:
        pos[n // 2 : n // 2 + 300] = pos[n // 2]  # a clump: whole warps hit early
    gpu_ctx.bind_column(1, POS, pos)
    gpu_ctx.upload(1, POS)
    sid = gpu_ctx.create_system(_lib.SYS_NBODY_COLLISION, [POS])
    gpu_ctx.set_archetypes(sid, [1])
    gpu_ctx.set_query_archetypes(sid, [1])
    gpu_ctx.run_system(sid)
    rows = gpu_ctx.system_removals(sid, 1)
    want = np.flatnonzero(_collision_mask(pos)).astype(np.uint32)
    assert np.array_equal(rows, want)
    if n >= 2:
        assert 0 in rows and 1 in rows
    if n >= 2000:
        assert 0 < len(want) < n  # neither trivial case


def test_collision_over_two_archetypes(gpu_ctx):
    """Outer rows in two archetypes, the nested query over both: an entity is only skipped against ITSELF (same row of
    the same archetype); rows with equal index in the other archetype are ordinary neighbours."""
    POS = 1
    rng = np.random.default_rng(77)
    a = (rng.random((1300, 3), dtype=np.float32) * f32(3.0)).astype(np.float32)
    b = (rng.random((700, 3), dtype=np.float32) * f32(3.0)).astype(np.float32)
    b[5] = a[5]  # same row index, other archetype: a collision
    for arch, col in ((1, a), (2, b)):
        gpu_ctx.bind_column(arch, POS, col)
        gpu_ctx.upload(arch, POS)
    sid = gpu_ctx.create_system(_lib.SYS_NBODY_COLLISION, [POS])
    gpu_ctx.set_archetypes(sid, [1, 2])
    gpu_ctx.set_query_archetypes(sid, [1, 2])
    gpu_ctx.run_system(sid)
    want_a = _collision_mask(a) | _collision_mask(a, qpos=b, same=False)
    want_b = _collision_mask(b) | _collision_mask(b, qpos=a, same=False)
    assert np.array_equal(gpu_ctx.system_removals(sid, 1), np.flatnonzero(want_a).astype(np.uint32))
    assert np.array_equal(gpu_ctx.system_removals(sid, 2), np.flatnonzero(want_b).astype(np.uint32))
    assert want_a[5] and want_b[5]


def test_ground_collision_flags(gpu_ctx):
    POS = 1
    pos = np.array([[0, 1, 0], [0, 0, 0], [0, -0.0, 5], [1, -3, 1], [2, 1e-9, 2]], np.float32)
    gpu_ctx.bind_column(2, POS, pos)
    gpu_ctx.upload(2, POS)
    sid = gpu_ctx.create_system(_lib.SYS_RAIN_GROUND_COLLISION, [POS])
    gpu_ctx.set_archetypes(sid, [2])
    gpu_ctx.run_system(sid)
    assert list(gpu_ctx.system_removals(sid, 2)) == [1, 2, 3]  # y <= 0 (crates/rain-simulation/src/lib.rs:153)


# ---------------------------------------------------------------------------------------------------
# n-body demo system set through the Application: gravity, collision -> acceleration -> movement
# ---------------------------------------------------------------------------------------------------
def test_nbody_demo_with_collision_through_application(gpu_ctx, nbody_oracle):
    POS, MASS, VEL, ACC = 1, 2, 3, 4
    n, ticks = 240, 4
    rng = np.random.default_rng(42)
    pos = (rng.random((n, 3), dtype=np.float32) * f32(1.2)).astype(np.float32)
    mass = np.full(n, 1.0e3, np.float32)  # light bodies: gravity barely moves them within a few ticks
    vel = ((rng.random((n, 3), dtype=np.float32) - f32(0.5)) * f32(200.0)).astype(np.float32)
    acc = np.zeros((n, 3), np.float32)

    app = ecs.Application(gpu=gpu_ctx)
    w = app.world
    for c, dt, name in ((POS, F3, "Position"), (MASS, np.float32, "Mass"), (VEL, F3, "Velocity"), (ACC, F3, "Acceleration")):
        w.register_component(c, dt, name)
    # registration order of the demo (n_body_demo.rs:39-42)
    app.add_gpu_system(_lib.SYS_NBODY_MOVEMENT, [POS, VEL], "movement")
    app.add_gpu_system(_lib.SYS_NBODY_ACCELERATION, [VEL, ACC], "acceleration")
    app.add_gpu_system(_lib.SYS_NBODY_GRAVITY, [POS, ACC, MASS], "gravity")
    app.add_gpu_system(_lib.SYS_NBODY_COLLISION, [POS], "collision")
    w.create_entities(n, {POS: pos, MASS: mass, VEL: vel, ACC: acc})

    # oracle: same World operations + the CPU systems in the order the schedule produces
    ow = O.World()
    ents = [ow.create_entity({POS: tuple(pos[k]), MASS: float(mass[k]), VEL: tuple(vel[k]), ACC: tuple(acc[k])}) for k in range(n)]
    removed_total = 0
    for tick in range(ticks):
        app.run(1)
        order, batches = app.last_tick_order()
        names = [app.system_names[i] for i in order]
        # {gravity, collision} -> {acceleration} -> {movement}
        assert names.index("gravity") < names.index("acceleration") < names.index("movement")
        assert names.index("collision") < names.index("movement")

        cols = ow.archetypes[1].columns
        p = np.array(cols[POS], np.float32).reshape(-1, 3)
        m = np.array(cols[MASS], np.float32)
        v = np.array(cols[VEL], np.float32).reshape(-1, 3)
        a = nbody_oracle.gravity(np.ascontiguousarray(p), np.ascontiguousarray(m))
        mask = _collision_mask(p)
        v = np.ascontiguousarray(v)
        p = np.ascontiguousarray(p)
        nbody_oracle.axpy(v, a, nbody_oracle.dt)  # acceleration
        nbody_oracle.axpy(p, v, nbody_oracle.dt)  # movement
        cols[POS] = [tuple(r) for r in p]
        cols[VEL] = [tuple(r) for r in v]
        cols[ACC] = [tuple(r) for r in a]
        doomed = [ow.archetypes[1].entities[i] for i in np.flatnonzero(mask)]  # ascending component index
        for e in doomed:
            ow.delete_entity(e)
        removed_total += len(doomed)

        created, removed = app.last_playback()
        assert (created, removed) == (0, len(doomed))
        app.sync_to_host()
        assert w.archetype_entities(1) == [(e.id, e.generation) for e in ow.archetypes[1].entities]
        op = np.array(ow.archetypes[1].columns[POS], np.float32).reshape(-1, 3)
        ov = np.array(ow.archetypes[1].columns[VEL], np.float32).reshape(-1, 3)
        assert rel_err_rows(w.column(1, POS), op, floor=1.0).max() <= 1e-5
        assert rel_err_rows(w.column(1, VEL), ov, floor=1.0).max() <= 1e-4
    assert removed_total > 10  # the scene does collide


def test_nbody_demo_scenes_two_body_archetypes(gpu_ctx, nbody_oracle):
    """The demo's own scenes (crates/n-body/examples/n_body_demo.rs:44-45): SMALL_HEAVY_CLUSTERS planets and the
    SINGLE_HEAVY_BODY_AT_ORIGIN sun, which carries a PointLight and therefore lives in a SECOND body archetype
    (crates/n-body/src/scenes.rs:146-168).  Every system of the demo matches both archetypes; the nested queries of
    gravity and collision run over their concatenation in archetype order."""
    POS, MASS, VEL, ACC, LIGHT = 1, 2, 3, 4, 5
    from recs_b200 import scenes

    planets = scenes.spawn_clusters(scenes.SMALL_HEAVY_CLUSTERS, 3)
    sun = scenes.spawn_bodies(scenes.ONE_HEAVY_BODY, 4)
    app = ecs.Application(gpu=gpu_ctx)
    w = app.world
    for c, dt, name in ((POS, F3, "Position"), (MASS, np.float32, "Mass"), (VEL, F3, "Velocity"), (ACC, F3, "Acceleration"),
                        (LIGHT, F3, "PointLight")):
        w.register_component(c, dt, name)
    app.add_gpu_system(_lib.SYS_NBODY_MOVEMENT, [POS, VEL], "movement")
    app.add_gpu_system(_lib.SYS_NBODY_ACCELERATION, [VEL, ACC], "acceleration")
    app.add_gpu_system(_lib.SYS_NBODY_GRAVITY, [POS, ACC, MASS], "gravity")
    app.add_gpu_system(_lib.SYS_NBODY_COLLISION, [POS], "collision")
    n = len(planets[1])
    w.create_entities(n, {POS: planets[0], MASS: planets[1], VEL: planets[2], ACC: planets[3]})
    w.create_entity({LIGHT: np.array([0.9, 0.8, 0.1], np.float32), POS: sun[0][0], MASS: sun[1][0], VEL: sun[2][0], ACC: sun[3][0]})
    assert w.archetype_count() == 3
    pos = np.concatenate([planets[0], sun[0]])
    mass = np.concatenate([planets[1], sun[1]])
    vel = np.concatenate([planets[2], sun[2]])
    alive = np.ones(n + 1, bool)
    ticks, removed_total = 5, 0
    for tick in range(ticks):
        app.run(1)
        names = [app.system_names[i] for i in app.last_tick_order()[0]]
        assert names.index("gravity") < names.index("acceleration") < names.index("movement")
        p = np.ascontiguousarray(pos[alive])
        m = np.ascontiguousarray(mass[alive])
        v = np.ascontiguousarray(vel[alive])
        a = nbody_oracle.gravity(p, m)
        hit = _collision_mask(p)
        nbody_oracle.axpy(v, a, nbody_oracle.dt)
        nbody_oracle.axpy(p, v, nbody_oracle.dt)
        pos[alive], vel[alive] = p, v
        assert app.last_playback() == (0, int(hit.sum())), tick
        removed_total += int(hit.sum())
        # entities leave by swap_remove: compare as sets of rows keyed by entity id (planet k = entity k, the sun = entity n)
        idx = np.flatnonzero(alive)
        alive[idx[hit]] = False
    app.sync_to_host()
    got = {}
    for arch in (1, 2):
        ents = w.archetype_entities(arch)
        cp, cv = w.column(arch, POS), w.column(arch, VEL)
        for row, (eid, gen) in enumerate(ents):
            got[eid] = (cp[row].copy(), cv[row].copy())
    assert sorted(got) == list(np.flatnonzero(alive))
    gp = np.array([got[k][0] for k in sorted(got)])
    gv = np.array([got[k][1] for k in sorted(got)])
    assert rel_err_rows(gp, pos[alive], floor=1.0).max() <= 1e-5
    assert rel_err_rows(gv, vel[alive], floor=1.0).max() <= 1e-4
    app.close()


# ---------------------------------------------------------------------------------------------------
# the complete rain tick: clouds spawn drops (CPU, Commands::create), drops fall (GPU), drops that
# reach the ground are removed (GPU decides, host plays back)
# ---------------------------------------------------------------------------------------------------
RAIN_ON_DEVICE = """
struct Cloud { float vapor_accumulation_rate, width, height; };
struct Position { float x, y, z; };
struct Mass { float m; };
struct Drop { float vx, vy, vz; float mass; float px, py, pz; };   // (Velocity::default(), Mass(0.1), Position { drop_position })
// crates/rain-simulation/src/lib.rs:128-148
void rain(const Cloud& cloud, const Position& position, Mass& mass, const recs_commands& commands) {
    while (mass.m > 0.1f) {
        commands.create(Drop{0.0f, 0.0f, 0.0f, 0.1f, position.x, position.y, position.z});  // deterministic_point_on_surface = center (:45-47)
        mass.m -= 0.1f;
    }
}
// crates/rain-simulation/src/lib.rs:158-160
void vapor_accumulation(const Cloud& cloud, Mass& mass) { mass.m += cloud.vapor_accumulation_rate * 0.05f; }
"""


@pytest.mark.parametrize("worker_threads,rain_on_device", [(0, False), (3, False), (0, True), (3, True)])
def test_full_rain_tick_through_application(gpu_ctx, nbody_oracle, worker_threads, rain_on_device):
    """worker_threads = 0: Sequential executor; 3: WorkerPool -- the CPU systems of a batch run on worker threads
    while the GPU systems of the same batch are enqueued from the scheduling thread.
    rain_on_device: `rain` (which spawns the drops with Commands::create) and `vapor_accumulation` are generic GPU
    systems too, so the WHOLE tick of the reference runs on the device; only the create records and the removal
    rows come back for playback on the host World."""
    VEL, POS, MASS, CLOUD = 20, 21, 22, 23
    n_clouds, ticks = 9, 60
    app = ecs.Application(gpu=gpu_ctx)
    app.set_executor(worker_threads)
    w = app.world
    for c, dt, name in ((VEL, F3, "Velocity"), (POS, F3, "Position"), (MASS, np.float32, "Mass"), (CLOUD, F3, "Cloud")):
        w.register_component(c, dt, name)
    DT, MASS_PER_DROP = f32(0.05), f32(0.1)

    def rain(arch, count, cols):  # rain(Read<Cloud>, Read<Position>, Write<Mass>, Commands)   lib.rs:128-148
        pos, mass = _arr(cols[1], count, 3), _arr(cols[2], count, 1)
        for k in range(count):
            while mass[k] > MASS_PER_DROP:
                app.command_create({VEL: np.zeros(3, np.float32), MASS: MASS_PER_DROP, POS: pos[k].copy()})
                mass[k] = mass[k] - MASS_PER_DROP

    def vapor(arch, count, cols):  # vapor_accumulation(Read<Cloud>, Write<Mass>)   lib.rs:158-160
        cloud, mass = _arr(cols[0], count, 3), _arr(cols[1], count, 1)
        mass += cloud[:, 0] * DT

    ticks_seen = []
    # registration order of the benchmark (recs/rain_simulation.rs:56-65)
    app.add_gpu_system(_lib.SYS_RAIN_GRAVITY, [VEL], "gravity")
    app.add_gpu_system(_lib.SYS_RAIN_WIND, [VEL], "wind")
    app.add_gpu_system(_lib.SYS_RAIN_WIND_DIRECTION, [], "wind_direction")
    app.add_gpu_system(_lib.SYS_RAIN_MOVEMENT, [POS, VEL], "movement")
    if rain_on_device:
        app.add_gpu_kernel_system("rain", RAIN_ON_DEVICE, "rain", [("read", CLOUD), ("read", POS), ("write", MASS), ("commands",)],
                                  ["Cloud", "Position", "Mass"], creates=[VEL, MASS, POS], create_record_type="Drop")
    else:
        app.add_cpu_system("rain", [("read", CLOUD), ("read", POS), ("write", MASS), ("commands",)], rain)
    app.add_gpu_system(_lib.SYS_RAIN_GROUND_COLLISION, [POS], "ground_collision")
    if rain_on_device:
        app.add_gpu_kernel_system("vapor_accumulation", RAIN_ON_DEVICE, "vapor_accumulation", [("read", CLOUD), ("write", MASS)],
                                  ["Cloud", "Mass"])
    else:
        app.add_cpu_system("vapor_accumulation", [("read", CLOUD), ("write", MASS)], vapor)
    app.add_cpu_system("benchmark_system", [], lambda a, c, cols: ticks_seen.append(1))
    # create_evenly_interspersed_clouds (crates/rain-simulation/src/scene.rs:27-54); low clouds so drops land soon
    side = int(np.sqrt(n_clouds))
    ow = O.World()
    for x in range(side):
        for z in range(side):
            comps = {CLOUD: (1.0, 10.0, 10.0), POS: (x * 10.0, 2.0, z * 10.0), MASS: 0.0}
            w.create_entity({k: np.asarray(v, np.float32) for k, v in comps.items()})
            ow.create_entity(dict(comps))

    direction = np.array([1.0, 0.0, 0.0], np.float32)
    total_created = total_removed = 0
    for tick in range(ticks):
        app.run(1)
        order, _ = app.last_tick_order()
        names = [app.system_names[i] for i in order]
        for before, after in (("wind", "gravity"), ("gravity", "movement"), ("rain", "movement"),
                              ("ground_collision", "movement"), ("vapor_accumulation", "rain")):
            assert names.index(before) < names.index(after)
        # ---- oracle tick in the same order: wind, wind_direction, ground_collision, vapor, [gravity, rain], movement
        drops = ow.archetypes[2] if len(ow.archetypes) > 2 else None
        clouds = ow.archetypes[1]
        to_remove = []
        if drops is not None and drops.entities:
            v = np.array(drops.columns[VEL], np.float32).reshape(-1, 3)
            p = np.array(drops.columns[POS], np.float32).reshape(-1, 3)
            nbody_oracle.rain_wind(v, direction)
        direction = nbody_oracle.rain_wind_direction(direction)
        # ground_collision over every archetype with a Position, ascending archetype / row
        for ai in (1, 2):
            if ai < len(ow.archetypes):
                arch = ow.archetypes[ai]
                pp = p if (ai == 2 and drops is not None and drops.entities) else np.array(arch.columns[POS], np.float32).reshape(-1, 3)
                to_remove += [arch.entities[i] for i in range(len(arch.entities)) if pp[i, 1] <= 0.0]
        cm = np.array(clouds.columns[MASS], np.float32)
        cm = (cm + np.array([c[0] for c in clouds.columns[CLOUD]], np.float32) * DT).astype(np.float32)
        if drops is not None and drops.entities:
            nbody_oracle.rain_gravity(v)
        creates = []
        for k in range(len(cm)):
            while cm[k] > MASS_PER_DROP:
                creates.append({VEL: (0.0, 0.0, 0.0), MASS: float(MASS_PER_DROP), POS: clouds.columns[POS][k]})
                cm[k] = cm[k] - MASS_PER_DROP
        clouds.columns[MASS] = [float(x) for x in cm]
        if drops is not None and drops.entities:
            nbody_oracle.rain_movement(p, v)
            drops.columns[VEL] = [tuple(r) for r in v]
            drops.columns[POS] = [tuple(r) for r in p]
        for comps in creates:  # playback: creates first, then removes
            ow.create_entity(dict(comps))
        for e in to_remove:
            ow.delete_entity(e)
        total_created += len(creates)
        total_removed += len(to_remove)
        assert app.last_playback() == (len(creates), len(to_remove)), tick

    # structural changes did not cost a round trip: the drop columns stayed on the device for all 60 ticks, only the
    # rows of new drops (12 B velocity + 12 B position each), the swap-remove move lists (8 B per move) and the nine
    # clouds' columns crossed PCIe, and nothing was downloaded before the final sync
    h2d, d2h = gpu_ctx.transfer_bytes()
    # (on the device, every create record comes back once: 16-byte order tag + the 28-byte tuple rounded up to 8)
    assert d2h == (total_created * 48 if rain_on_device else 0)
    assert h2d <= total_created * 24 + total_removed * 8 + 9 * 12 * 4, (h2d, total_created, total_removed)
    app.sync_to_host()
    assert len(ticks_seen) == ticks and total_created > 20 and total_removed > 5
    assert w.archetype_count() == len(ow.archetypes) == 3
    assert w.archetype_components(2) == sorted([VEL, MASS, POS])
    for ai in (1, 2):
        assert w.archetype_entities(ai) == [(e.id, e.generation) for e in ow.archetypes[ai].entities]
    # every streaming system is bit-exact, so the whole state is
    d = ow.archetypes[2]
    assert np.array_equal(w.column(2, POS).view(np.uint32), np.array(d.columns[POS], np.float32).reshape(-1, 3).view(np.uint32))
    assert np.array_equal(w.column(2, VEL).view(np.uint32), np.array(d.columns[VEL], np.float32).reshape(-1, 3).view(np.uint32))
    assert np.array_equal(w.column(1, MASS).view(np.uint32), np.array(ow.archetypes[1].columns[MASS], np.float32).view(np.uint32))


# ---------------------------------------------------------------------------------------------------
# device-resident structural changes under random create / remove batches
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_resident_columns_follow_random_creates_and_removes(gpu_ctx, seed):
    """Columns written by a GPU system every tick (so their newest copy is always on the device) must follow the
    host World through batches of Commands::create / Commands::remove without a round trip: appended rows are
    uploaded alone (recs_gpu_column_append), swap-removes are replayed as row moves
    (recs_gpu_archetype_move_rows), for 12-byte, 8-byte and 6-byte (byte path) elements.  Oracle: the same
    operations on the Python World (crates/ecs/src/archetypes.rs:200-243) + the body in numpy."""
    A, B, SIX = 31, 32, 33
    six_t = np.dtype((np.uint8, 6))
    src = ("struct A { float x, y, z; };\nstruct B { unsigned long long k; };\nstruct Six { unsigned char v[6]; };\n"
           "void bump(A& a, B& b, Six& s) { a.x += 1.0f; a.z -= 0.5f; b.k += 3ull; s.v[0] += 1; s.v[5] += 2; }\n")
    rng = np.random.default_rng(seed)
    app = ecs.Application(gpu=gpu_ctx)
    w = app.world
    w.register_component(A, F3, "A")
    w.register_component(B, np.dtype(np.uint64), "B")
    w.register_component(SIX, six_t, "Six")
    app.add_gpu_kernel_system("bump", src, "bump", [("write", A), ("write", B), ("write", SIX)], ["A", "B", "Six"])
    pending = {"creates": [], "removes": []}

    def commands(arch, count, cols):  # a system with only Commands runs once per tick
        for comps in pending["creates"]:
            app.command_create(comps)
        for e in pending["removes"]:
            app.command_remove(e)

    app.add_cpu_system("commands", [("commands",)], commands)
    ow = O.World()
    serial = [0]

    def new_values():
        serial[0] += 1
        k = serial[0]
        return {A: np.array([k, -k, 0.25 * k], np.float32), B: np.uint64(1000 * k), SIX: np.array([k % 200, 1, 2, 3, 4, k % 100], np.uint8)}

    def oracle_create(vals):
        ow.create_entity({A: tuple(float(x) for x in vals[A]), B: int(vals[B]), SIX: tuple(int(x) for x in vals[SIX])})

    for _ in range(700):  # initial population, created directly
        vals = new_values()
        w.create_entity(vals)
        oracle_create(vals)
    for tick in range(25):
        live = list(ow.archetypes[1].entities) if len(ow.archetypes) > 1 else []
        n_remove = int(rng.integers(0, min(len(live), 120) + 1)) if tick % 5 != 4 else 0
        victims = [live[i] for i in rng.choice(len(live), size=n_remove, replace=False)] if n_remove else []
        if victims and tick % 3 == 0:
            victims += victims[:3]  # duplicates collapse (the reference collects removals into a set)
        creates = [new_values() for _ in range(int(rng.integers(0, 90)) if tick % 4 != 3 else 0)]
        pending["creates"] = creates
        pending["removes"] = [ecs.Entity(e.id, e.generation) for e in victims]
        app.run(1)
        # oracle: the body over every row alive during the tick, then playback (creates first, then removes)
        arch = ow.archetypes[1]
        arch.columns[A] = [(a[0] + 1.0, a[1], a[2] - 0.5) for a in arch.columns[A]]
        arch.columns[B] = [b + 3 for b in arch.columns[B]]
        arch.columns[SIX] = [((s[0] + 1) % 256, s[1], s[2], s[3], s[4], (s[5] + 2) % 256) for s in arch.columns[SIX]]
        for vals in creates:
            oracle_create(vals)
        done = set()
        for e in victims:
            if (e.id, e.generation) not in done:
                done.add((e.id, e.generation))
                ow.delete_entity(e)
        assert app.last_playback() == (len(creates), len(done)), tick
    h2d, d2h = gpu_ctx.transfer_bytes()
    assert d2h == 0  # nothing came back before the final sync
    assert h2d <= (700 + serial[0]) * (12 + 8 + 6) + 25 * 120 * 8 + 1024
    app.sync_to_host()
    arch = ow.archetypes[1]
    assert w.archetype_entities(1) == [(e.id, e.generation) for e in arch.entities]
    n = len(arch.entities)
    assert np.array_equal(w.column(1, A)[:n], np.array(arch.columns[A], np.float32).reshape(n, 3))
    assert np.array_equal(w.column(1, B)[:n].reshape(-1), np.array(arch.columns[B], np.uint64))
    assert np.array_equal(w.column(1, SIX)[:n].reshape(n, 6), np.array(arch.columns[SIX], np.uint8).reshape(n, 6))
    app.close()


def test_resident_column_primitives_report_misuse(gpu_ctx):
    """recs_gpu_column_append / recs_gpu_archetype_move_rows: status codes for unbound columns, shrinking appends and
    moves that would need rows which were never appended; a correct call sequence is checked against numpy."""
    import ctypes as C

    from recs_b200 import RecsGpuError
    from recs_b200._lib import lib

    ctx = gpu_ctx
    COMP, ARCH = 5, 3
    store = np.arange(40, dtype=np.float32).reshape(10, 4)  # capacity 10 rows of 16 B
    st = lib.recs_gpu_column_append(ctx._ctx, ARCH, COMP, store.ctypes.data_as(C.c_void_p), 4)
    assert st == 5  # RECS_ERR_MISSING_PARAMETER: not bound
    ctx.clear_error()
    ctx.bind_column(ARCH, COMP, store[:6])
    ctx.upload(ARCH, COMP)
    assert lib.recs_gpu_column_append(ctx._ctx, ARCH, COMP, store.ctypes.data_as(C.c_void_p), 5) == 1  # rows can only be added
    ctx.clear_error()
    comps = (C.c_uint64 * 1)(COMP)
    dst, src = (C.c_uint32 * 2)(0, 3), (C.c_uint32 * 2)(7, 6)
    assert lib.recs_gpu_archetype_move_rows(ctx._ctx, ARCH, comps, 1, dst, src, 2, 8) == 7  # RECS_ERR_SIZE_MISMATCH: append first
    ctx.clear_error()
    # append rows 6..7, then fold of swap_remove(0), swap_remove(3): row 0 <- row 7, row 3 <- row 6, 6 rows remain
    assert lib.recs_gpu_column_append(ctx._ctx, ARCH, COMP, store.ctypes.data_as(C.c_void_p), 8) == 0
    assert lib.recs_gpu_archetype_move_rows(ctx._ctx, ARCH, comps, 1, dst, src, 2, 6) == 0
    want = store[:8].copy()
    want[0], want[3] = store[7], store[6]
    out = np.zeros((6, 4), np.float32)
    assert lib.recs_gpu_column_rebind_host(ctx._ctx, ARCH, COMP, out.ctypes.data_as(C.c_void_p)) == 0
    ctx.download(ARCH, COMP)
    assert np.array_equal(out, want[:6])
    with pytest.raises(RecsGpuError):
        ctx._check(lib.recs_gpu_column_rebind_host(ctx._ctx, ARCH, 99, None), "rebind")
    ctx.clear_error()


def test_add_and_remove_component_move_rows_between_resident_archetypes(gpu_ctx):
    """Application::add_component / remove_component (crates/ecs/src/lib.rs:213-228) between ticks: the entity changes
    archetype (crates/ecs/src/archetypes.rs:474-518, 556-651).  Both archetypes' columns are written by a GPU system
    every tick, so their newest copy is on the device; they must follow the moves without the invalidate -> re-bind ->
    full re-upload path: per move one row per column comes down, one row per column goes up, one row move is replayed."""
    POS, VEL, TAG = 41, 42, 43
    app = ecs.Application(gpu=gpu_ctx)
    w = app.world
    w.register_component(POS, F3, "Position")
    w.register_component(VEL, F3, "Velocity")
    w.register_component(TAG, np.dtype(np.float32), "Tag")
    app.add_gpu_system(_lib.SYS_QUERY_ADD, [POS, VEL], "movement")  # position += velocity over every archetype with both
    ow = O.World()
    rng = np.random.default_rng(19)
    n = 5000
    ents = []
    for k in range(n):
        comps = {POS: rng.standard_normal(3).astype(np.float32), VEL: rng.standard_normal(3).astype(np.float32)}
        if k % 7 == 0:
            comps[TAG] = np.float32(k)
        ents.append(w.create_entity(comps))
        ow.create_entity({c: (tuple(float(x) for x in v) if c != TAG else float(v)) for c, v in comps.items()})
    oracle_ents = [e for a in ow.archetypes for e in a.entities]
    by_id = {e.id: e for e in oracle_ents}

    def oracle_tick():
        for arch in ow.archetypes:
            if POS in arch.columns and VEL in arch.columns:
                p = np.array(arch.columns[POS], np.float32).reshape(-1, 3)
                v = np.array(arch.columns[VEL], np.float32).reshape(-1, 3)
                arch.columns[POS] = [tuple(r) for r in (p + v)]

    app.run(2)
    oracle_tick()
    oracle_tick()
    h2d_before, d2h_before = gpu_ctx.transfer_bytes()
    moved = 0
    for round_ in range(6):
        picks = rng.choice(n, size=40, replace=False)
        for k in picks:
            e = ents[k]
            arch = w.entity_archetype(e)
            has_tag = TAG in w.archetype_components(arch)
            if has_tag:
                app.remove_components(e, [TAG])
                ow.remove_components(by_id[e.id], [TAG])
            else:
                app.add_components(e, {TAG: np.float32(1000 + k)})
                ow.add_components(by_id[e.id], {TAG: float(1000 + k)})
            moved += 1
        app.run(2)
        oracle_tick()
        oracle_tick()
    h2d, d2h = gpu_ctx.transfer_bytes()
    # per move: the moved row of Position and Velocity down (24 B) and up (24 B) + one 8-byte move-list entry; nothing
    # resembling the 2 x 60 KB columns per archetype a re-upload would cost each time
    assert d2h - d2h_before <= moved * 24
    assert h2d - h2d_before <= moved * (24 + 8)
    app.sync_to_host()
    assert w.archetype_count() == len(ow.archetypes)
    for ai, arch in enumerate(ow.archetypes):
        assert w.archetype_entities(ai) == [(e.id, e.generation) for e in arch.entities], ai
        if POS in arch.columns and arch.entities:
            m = len(arch.entities)
            assert np.array_equal(w.column(ai, POS)[:m].view(np.uint32), np.array(arch.columns[POS], np.float32).reshape(m, 3).view(np.uint32)), ai
            assert np.array_equal(w.column(ai, VEL)[:m].view(np.uint32), np.array(arch.columns[VEL], np.float32).reshape(m, 3).view(np.uint32)), ai
    app.close()
