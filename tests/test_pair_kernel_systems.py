"""Generic GPU systems with a nested `Query<(...)>` (recs_gpu_system_create_pair_kernel): the shape of the reference's
`gravity` (crates/n-body/src/lib.rs:104-135) and `collision` (crates/n-body/examples/n_body_demo.rs:54-71).

The body is called for the nested query's entities in query order, one thread per outer entity, compiled with
--fmad=false and IEEE division / square root -- so `gravity` written this way must reproduce the CPU oracle's
reference-order f32 result BIT FOR BIT (the hand-written kernel is only within 1e-4 of it: it sums in tiles and uses
rsqrt).  That makes it the exactness cross-check of the restated arithmetic on the device."""
import numpy as np
import pytest

from recs_b200 import scenes
from recs_b200.gpu import compile_pair_kernel_system

POS, MASS, VEL, ACC = 1, 2, 3, 4

# crates/n-body/src/lib.rs:104-135, one expression per line of the Rust body (shipped as an example in recs_b200/nbody.py)
from recs_b200.nbody import REFERENCE_ORDER_GRAVITY_PARAMS as GRAVITY_PARAMS  # noqa: E402
from recs_b200.nbody import REFERENCE_ORDER_GRAVITY_SOURCE as GRAVITY  # noqa: E402

# crates/n-body/examples/n_body_demo.rs:54-71
COLLISION = """
struct Position { float x, y, z; };
struct Hit { int any; };
struct Radius { float r2; };
void pair(Hit& hit, const unsigned long long& entity, const Position& position, const unsigned long long& other_entity,
          const Position& other_position, const Radius& u) {
    if (other_entity == entity) return;
    const float dx = other_position.x - position.x, dy = other_position.y - position.y, dz = other_position.z - position.z;
    if ((dx * dx + dy * dy) + dz * dz < u.r2) hit.any = 1;       // distance2 < COLLISION_RADIUS * COLLISION_RADIUS
}
void finish(const unsigned long long& entity, const Position& position, const recs_commands& commands, const Hit& hit,
            const Radius& u) {
    if (hit.any) commands.remove();
}
"""
COLLISION_PARAMS = [(0, 8, "entity", "unsigned long long"), (POS, 12, "read", "Position"),
                    (0, 8, "query_entity", "unsigned long long"), (POS, 12, "query_read", "Position"), (0, 0, "commands", None)]


def test_pair_kernels_compile_without_a_gpu():
    st, log, cubin = compile_pair_kernel_system(GRAVITY, "Total", "pair", "finish", GRAVITY_PARAMS)
    assert st == 0 and cubin > 1000, log
    st, log, cubin = compile_pair_kernel_system(COLLISION, "Hit", "pair", "finish", COLLISION_PARAMS, "Radius")
    assert st == 0 and cubin > 1000, log
    st, log, _ = compile_pair_kernel_system(GRAVITY.replace("total.z += cz;", "total.w += cz;"), "Total", "pair", "finish", GRAVITY_PARAMS)
    assert st == 12 and "has no member" in log and "recs_system_body.cu" in log  # the user's own line numbers


def _bind(ctx, arch, comp, array):
    ctx.bind_column(arch, comp, array)
    ctx.upload(arch, comp)


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 2, 255, 257, 1000, 5000])
def test_generic_gravity_reproduces_the_reference_order_sum_bitwise(gpu_ctx, nbody_oracle, n):
    ctx = gpu_ctx
    pos, mass, vel, acc = scenes.spawn_bodies(scenes.all_heavy_random_cube_with_bodies(n), 13)
    if n >= 255:
        pos[7] = pos[3]  # coincident
        pos[11] = pos[5] + np.array([1e-4, 0, 0], np.float32)  # below epsilon
    for comp, arr in ((POS, pos), (MASS, mass), (ACC, acc)):
        _bind(ctx, 1, comp, arr)
    sid = ctx.create_pair_kernel_system("gravity", GRAVITY, "Total", "pair", "finish", GRAVITY_PARAMS)
    # accesses in parameter order: Read<Position>, Write<Acceleration>, then the query's (crates/ecs/src/systems/mod.rs:973-978)
    assert ctx.describe_system(sid) == ("gravity", [(POS, False), (ACC, True), (POS, False), (MASS, False)])
    ctx.set_archetypes(sid, [1])
    ctx.set_query_archetypes(sid, [1])
    ctx.run_system(sid)
    ctx.download(1, ACC)
    want = nbody_oracle.gravity(pos, mass)
    assert np.array_equal(acc.view(np.uint32), want.view(np.uint32))


@pytest.mark.gpu
def test_generic_gravity_over_several_archetypes_follows_query_order(gpu_ctx, nbody_oracle):
    """Outer rows in two archetypes, nested query over three (one of them not an outer archetype, one empty): every outer
    body must see the concatenation of the query archetypes in the order given."""
    ctx = gpu_ctx
    pos, mass, vel, acc = scenes.spawn_bodies(scenes.all_heavy_random_cube_with_bodies(900), 14)
    cut = [(0, 300), (300, 300), (300, 650), (650, 900)]  # archetypes 1..4; archetype 2 is empty
    accs = {}
    for a, (b, e) in enumerate(cut, start=1):
        _bind(ctx, a, POS, np.ascontiguousarray(pos[b:e]))
        _bind(ctx, a, MASS, np.ascontiguousarray(mass[b:e]))
        accs[a] = np.zeros((e - b, 3), np.float32)
        _bind(ctx, a, ACC, accs[a])
    sid = ctx.create_pair_kernel_system("gravity", GRAVITY, "Total", "pair", "finish", GRAVITY_PARAMS)
    ctx.set_archetypes(sid, [1, 2, 3])           # outer: bodies 0..650
    ctx.set_query_archetypes(sid, [4, 2, 1, 3])  # query order: 650..900, (empty), 0..300, 300..650
    ctx.run_system(sid)
    order = np.concatenate([np.arange(650, 900), np.arange(0, 300), np.arange(300, 650)])
    want = nbody_oracle.gravity(pos[:650].copy(), mass, qpos=np.ascontiguousarray(pos[order]), qmass=np.ascontiguousarray(mass[order]))
    for a, (b, e) in ((1, cut[0]), (3, cut[2])):
        ctx.download(a, ACC)
        assert np.array_equal(accs[a].view(np.uint32), want[b:e].view(np.uint32)), a


@pytest.mark.gpu
@pytest.mark.parametrize("n", [2, 257, 1500])
def test_generic_collision_matches_the_builtin_kind(gpu_ctx, n):
    import struct

    from recs_b200 import _lib

    ctx = gpu_ctx
    rng = np.random.default_rng(n)
    pos = (rng.random((n, 3), dtype=np.float32) * np.float32(1.5)).astype(np.float32)
    pos[1] = pos[0]  # coincident distinct entities do collide; only the entity itself is skipped
    _bind(ctx, 1, POS, pos)
    _bind(ctx, 1, _lib.ENTITY_COMPONENT, (np.arange(n, dtype=np.uint64) + (3 << 32)))
    gen = ctx.create_pair_kernel_system("collision", COLLISION, "Hit", "pair", "finish", COLLISION_PARAMS, "Radius", 4)
    assert ctx.describe_system(gen) == ("collision", [(POS, False), (POS, False)])
    ctx.set_system_uniform(gen, struct.pack("f", np.float32(0.1) * np.float32(0.1)))
    builtin = ctx.create_system(_lib.SYS_NBODY_COLLISION, [POS])
    for s in (gen, builtin):
        ctx.set_archetypes(s, [1])
        ctx.set_query_archetypes(s, [1])
        ctx.run_system(s)
    rows = ctx.system_removals(gen, 1)
    assert np.array_equal(rows, ctx.system_removals(builtin, 1))
    assert 0 in rows and 1 in rows


@pytest.mark.gpu
def test_pair_kernel_errors(gpu_ctx):
    from recs_b200 import RecsGpuError

    ctx = gpu_ctx
    with pytest.raises(RecsGpuError) as e:  # no nested query parameter
        ctx.create_pair_kernel_system("g", GRAVITY, "Total", "pair", "finish", GRAVITY_PARAMS[:2])
    assert e.value.code == 1
    ctx.clear_error()
    sid = ctx.create_pair_kernel_system("gravity", GRAVITY, "Total", "pair", "finish", GRAVITY_PARAMS)
    pos = np.zeros((4, 3), np.float32)
    _bind(ctx, 1, POS, pos)
    _bind(ctx, 1, ACC, np.zeros((4, 3), np.float32))
    ctx.set_archetypes(sid, [1])
    ctx.set_query_archetypes(sid, [1])
    with pytest.raises(RecsGpuError) as e:  # the query's Mass column is not bound
        ctx.run_system(sid)
    assert e.value.code == 5
    ctx.clear_error()


@pytest.mark.gpu
def test_whole_nbody_tick_bit_exact_through_the_application(gpu_ctx, nbody_oracle):
    """The n-body benchmark's system set (crates/ecs_bench_suite/src/recs/n_body.rs:54-59) with `gravity` as a generic
    nested-query system and the built-in movement / acceleration kinds, scheduled by the PrecedenceGraph: after five
    ticks positions, velocities AND accelerations equal the oracle's sequential run bit for bit -- the reference's
    arithmetic, executed on the device end to end."""
    from recs_b200 import _lib
    from recs_b200.ecs import Application

    F3 = np.dtype((np.float32, 3))
    app = Application(gpu_ctx)
    w = app.world
    for comp, dt, name in ((POS, F3, "Position"), (MASS, np.float32, "Mass"), (VEL, F3, "Velocity"), (ACC, F3, "Acceleration")):
        w.register_component(comp, dt, name)
    n, ticks = 1000, 5  # configs[0]: 1 000 bodies, EVERYTHING_HEAVY
    pos, mass, vel, acc = scenes.spawn_bodies(scenes.EVERYTHING_HEAVY, 1)
    ref = [a.copy() for a in (pos, mass, vel, acc)]
    app.add_gpu_system(_lib.SYS_NBODY_MOVEMENT, [POS, VEL], "movement")
    app.add_gpu_system(_lib.SYS_NBODY_ACCELERATION, [VEL, ACC], "acceleration")
    app.add_gpu_pair_kernel_system(
        "gravity", GRAVITY, "Total", "pair", "finish",
        [("read", POS), ("write", ACC), ("query", [("read", POS), ("read", MASS)])],
        ["Position", "Acceleration", "Position", "Mass"])
    w.create_entities(n, {POS: pos, MASS: mass, VEL: vel, ACC: acc})
    app.run(ticks)
    order, _ = app.last_tick_order()
    assert [app.system_names[i] for i in order] == ["gravity", "movement", "acceleration"]
    app.sync_to_host()
    nbody_oracle.run_sequential(ref[0], ref[1], ref[2], ref[3], ticks)
    for comp, want in ((POS, ref[0]), (VEL, ref[2]), (ACC, ref[3])):
        got = w.column(1, comp)[:n].reshape(n, 3)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), comp
    app.close()
