"""Bit-exact parity of the HBM-bound streaming systems (rain, simple_iter / generic Write<A>,Read<B>
query) against the CPU oracle, through the C ABI."""
import numpy as np
import pytest

from recs_b200 import _lib, scenes

pytestmark = pytest.mark.gpu
VEL, POS, ARCH = 201, 202, 3


def _rain(ctx, n, seed, speed=0.0):
    vel, mass, pos = scenes.rain_drops(n, seed)
    if speed:
        rng = np.random.default_rng(seed + 100)
        vel[...] = (rng.random((n, 3), dtype=np.float32) - np.float32(0.5)) * np.float32(2 * speed)
    ctx.bind_column(ARCH, VEL, vel)
    ctx.bind_column(ARCH, POS, pos)
    ctx.upload(ARCH, VEL)
    ctx.upload(ARCH, POS)
    ids = {
        "wind": ctx.create_system(_lib.SYS_RAIN_WIND, [VEL]),
        "gravity": ctx.create_system(_lib.SYS_RAIN_GRAVITY, [VEL]),
        "movement": ctx.create_system(_lib.SYS_RAIN_MOVEMENT, [POS, VEL]),
        "wind_direction": ctx.create_system(_lib.SYS_RAIN_WIND_DIRECTION, []),
    }
    for k in ("wind", "gravity", "movement"):
        ctx.set_archetypes(ids[k], [ARCH])
    return vel, pos, ids


@pytest.mark.parametrize("n", [1, 3, 4, 5, 1000, 4099])
@pytest.mark.parametrize("speed", [0.0, 12.0])  # 12 m/s: many drops beyond TERMINAL_VELOCITY = 9
def test_rain_systems_individually_bit_exact(gpu_ctx, nbody_oracle, n, speed):
    vel, pos, ids = _rain(gpu_ctx, n, 1, speed)
    rv, rp = vel.copy(), pos.copy()
    direction = np.array([0.6, 0.0, 0.8], np.float32)
    gpu_ctx.set_wind_direction(direction)
    for name in ("gravity", "wind", "movement", "wind", "gravity"):
        gpu_ctx.run_system(ids[name])
        {"gravity": lambda: nbody_oracle.rain_gravity(rv), "wind": lambda: nbody_oracle.rain_wind(rv, direction),
         "movement": lambda: nbody_oracle.rain_movement(rp, rv)}[name]()
    gpu_ctx.download(ARCH, VEL)
    gpu_ctx.download(ARCH, POS)
    assert np.array_equal(vel.view(np.uint32), rv.view(np.uint32))
    assert np.array_equal(pos.view(np.uint32), rp.view(np.uint32))


@pytest.mark.parametrize("n", [7, 4096, 100003])
def test_rain_fused_tick_matches_unfused_oracle_order(gpu_ctx, nbody_oracle, n):
    """Tick order of the rain benchmark's DAG: wind -> gravity -> movement, with
    wind_direction rotating the global between ticks; 20 ticks so drops reach terminal velocity."""
    vel, pos, ids = _rain(gpu_ctx, n, 2)
    rv, rp = vel.copy(), pos.copy()
    gpu_ctx.set_tick_order([ids["wind_direction"], ids["wind"], ids["gravity"], ids["movement"]])
    direction = np.array([1.0, 0.0, 0.0], np.float32)
    for _ in range(20):
        direction = nbody_oracle.rain_wind_direction(direction)
        nbody_oracle.rain_wind(rv, direction)
        nbody_oracle.rain_gravity(rv)
        nbody_oracle.rain_movement(rp, rv)
    gpu_ctx.tick(20)
    gpu_ctx.download(ARCH, VEL)
    gpu_ctx.download(ARCH, POS)
    assert np.array_equal(gpu_ctx.get_wind_direction().view(np.uint32), direction.view(np.uint32))
    assert np.array_equal(vel.view(np.uint32), rv.view(np.uint32))
    assert np.array_equal(pos.view(np.uint32), rp.view(np.uint32))
    assert np.abs(rv[:, 1]).max() >= 9.0  # the terminal-velocity branch was exercised


@pytest.mark.parametrize("n", [0, 1, 2, 3, 10000, 1 << 20])
def test_simple_iter_query_add_bit_exact(gpu_ctx, nbody_oracle, n):
    """crates/ecs_bench_suite/src/recs/simple_iter.rs:47-49: position.0 += velocity.0; 100 passes."""
    A, B, ARCH2 = 11, 12, 5
    rng = np.random.default_rng(3)
    p = rng.standard_normal((n, 3)).astype(np.float32)
    v = rng.standard_normal((n, 3)).astype(np.float32)
    rp = p.copy()
    gpu_ctx.bind_column(ARCH2, A, p)
    gpu_ctx.bind_column(ARCH2, B, v)
    gpu_ctx.upload(ARCH2, A)
    gpu_ctx.upload(ARCH2, B)
    sid = gpu_ctx.create_system(_lib.SYS_QUERY_ADD, [A, B])
    gpu_ctx.set_archetypes(sid, [ARCH2])
    gpu_ctx.set_tick_order([sid])
    gpu_ctx.tick(100)
    for _ in range(100):
        nbody_oracle.add(rp, v)
    gpu_ctx.download(ARCH2, A)
    assert np.array_equal(p.view(np.uint32), rp.view(np.uint32))


def test_query_over_two_archetypes(gpu_ctx, nbody_oracle):
    """A (Write<A>, Read<B>) query matching several archetypes visits each one's rows (segments may
    span archetypes on the CPU, crates/ecs/src/systems/mod.rs:473-540; one launch per archetype here)."""
    A, B = 21, 22
    cols = {}
    for arch, n in ((2, 1000), (4, 37)):
        p = np.full((n, 3), float(arch), np.float32)
        v = np.full((n, 3), 0.25, np.float32)
        gpu_ctx.bind_column(arch, A, p)
        gpu_ctx.bind_column(arch, B, v)
        gpu_ctx.upload(arch, A)
        gpu_ctx.upload(arch, B)
        cols[arch] = p
    sid = gpu_ctx.create_system(_lib.SYS_QUERY_ADD, [A, B])
    gpu_ctx.set_archetypes(sid, [2, 4])
    gpu_ctx.run_system(sid)
    for arch, p in cols.items():
        gpu_ctx.download(arch, A)
        assert np.array_equal(p, np.full_like(p, arch + 0.25))


@pytest.mark.parametrize("split", [300, 301, 1])
def test_gravity_query_spanning_two_archetypes(gpu_ctx, nbody_oracle, split):
    """Nested Query<(Read<Position>, Read<Mass>)> matches two archetypes (bodies with and without an
    extra component); every outer body must feel all of them, in query order."""
    from recs_b200.nbody import ACCELERATION, MASS, POSITION
    from tests.util import rel_err_rows

    pos, mass, vel, acc = scenes.spawn_bodies(scenes.all_heavy_random_cube_with_bodies(777), 8)

    parts = {1: slice(0, split), 2: slice(split, 777)}
    bound = {}
    for arch, sl in parts.items():
        p, m, a = (np.ascontiguousarray(x[sl]) for x in (pos, mass, acc))
        gpu_ctx.bind_column(arch, POSITION, p)
        gpu_ctx.bind_column(arch, MASS, m)
        gpu_ctx.bind_column(arch, ACCELERATION, a)
        for c in (POSITION, MASS, ACCELERATION):
            gpu_ctx.upload(arch, c)
        bound[arch] = a
    sid = gpu_ctx.create_system(_lib.SYS_NBODY_GRAVITY, [POSITION, ACCELERATION, MASS])
    gpu_ctx.set_archetypes(sid, [1, 2])
    gpu_ctx.set_query_archetypes(sid, [1, 2])
    gpu_ctx.run_system(sid)
    ref = nbody_oracle.gravity(pos, mass)
    for arch, sl in parts.items():
        gpu_ctx.download(arch, ACCELERATION)
        assert rel_err_rows(bound[arch], ref[sl]).max() <= 1e-4
