"""GPU parity of the n-body systems against the CPU oracle (through the C ABI).

Tolerances (BASELINE.json north_star / SURVEY 8d): positions after the first tick are bit-exact
(movement runs before acceleration, so they do not depend on the new accelerations); acceleration
is within 1e-4 norm-wise per body of the reference-order f32 oracle; velocity is bit-exact given
identical acceleration input.
"""
import numpy as np
import pytest

from recs_b200 import scenes
from tests.util import rel_err_rows

pytestmark = pytest.mark.gpu

ACC_TOL = 1e-4  # north_star: "Accelerations and positions must agree within a relative tolerance of 1e-4"


def _app(ctx, n, seed, scene=None):
    from recs_b200.nbody import NBodyGpuApp

    scene = scene or scenes.all_heavy_random_cube_with_bodies(n)
    pos, mass, vel, acc = scenes.spawn_bodies(scene, seed)
    host = tuple(a.copy() for a in (pos, mass, vel, acc))
    app = NBodyGpuApp(ctx, pos, mass, vel, acc)
    app.upload()
    return app, host


@pytest.mark.parametrize("n", [1, 2, 3, 31, 255, 256, 257, 1000, 1024, 1025, 4097])
def test_gravity_matches_oracle(gpu_ctx, nbody_oracle, n):
    app, (pos, mass, vel, acc) = _app(gpu_ctx, n, seed=1)
    app.run("gravity")
    app.download()
    ref = nbody_oracle.gravity(pos, mass)
    err = rel_err_rows(app.acc, ref)
    assert np.all(np.isfinite(app.acc))
    assert err.max() <= ACC_TOL, f"n={n}: max norm-wise rel err {err.max():.3e}"
    # gravity writes only Acceleration
    assert np.array_equal(app.pos, pos) and np.array_equal(app.vel, vel)


def test_gravity_overwrites_initial_acceleration(gpu_ctx, nbody_oracle):
    # `*acceleration = total_acceleration` (crates/n-body/src/lib.rs:134): the scene's random
    # initial acceleration is discarded, a lone body ends with exactly zero.
    app, _ = _app(gpu_ctx, 1, seed=3)
    assert np.any(app.acc != 0)
    app.run("gravity")
    app.download()
    assert np.array_equal(app.acc, np.zeros((1, 3), np.float32))


def test_gravity_epsilon_mask_and_coincident_bodies(gpu_ctx, nbody_oracle):
    # pairs with d2 <= f32::EPSILON contribute zero (crates/n-body/src/lib.rs:118-120): coincident
    # bodies and bodies 1e-4 apart (d2 = 1e-8) must not blow up, bodies 1e-3 apart must interact.
    from recs_b200.nbody import NBodyGpuApp

    pos = np.array(
        [[0, 0, 0], [0, 0, 0], [1e-4, 0, 0], [5.0, 0, 0], [5.001, 0, 0], [-3.0, 4.0, 0.5]], dtype=np.float32
    )
    mass = np.full(6, 5.9e14, dtype=np.float32)
    vel = np.zeros((6, 3), np.float32)
    acc = np.ones((6, 3), np.float32)
    app = NBodyGpuApp(gpu_ctx, pos.copy(), mass, vel, acc)
    app.upload()
    app.run("gravity")
    app.download()
    ref = nbody_oracle.gravity(pos, mass)
    assert np.all(np.isfinite(app.acc))
    assert rel_err_rows(app.acc, ref).max() <= ACC_TOL
    # the close pair (1e-3 apart) dominates bodies 3 and 4
    assert abs(app.acc[3, 0]) > 1e9 and abs(app.acc[4, 0]) > 1e9


@pytest.mark.parametrize("n", [600, 3000, 20000])
def test_gravity_many_coincident_and_sub_epsilon_pairs(gpu_ctx, nbody_oracle, n):
    """Every body has a coincident twin in another tile and a neighbour 1e-4 m away (d2 = 1e-8 <=
    epsilon): the unmasked fast path must detect each affected tile and redo it exactly."""
    from recs_b200.nbody import NBodyGpuApp

    pos, mass, vel, acc = scenes.spawn_bodies(scenes.all_heavy_random_cube_with_bodies(n), 6)
    third = n // 3
    pos[third : 2 * third] = pos[:third]  # coincident twins
    pos[2 * third : 3 * third] = pos[:third] + np.array([1e-4, 0, 0], np.float32)  # sub-epsilon neighbours
    pos = np.ascontiguousarray(pos)
    app = NBodyGpuApp(gpu_ctx, pos.copy(), mass, vel, acc)
    app.upload()
    app.run("gravity")
    app.download()
    ref = nbody_oracle.gravity(pos, mass)
    assert np.all(np.isfinite(app.acc))
    # bodies ~1e-4 apart in f32 at |x| ~ 100 collapse to a few ulps: compare against the oracle on
    # the SAME f32 positions, norm-wise
    assert rel_err_rows(app.acc, ref).max() <= ACC_TOL


def test_movement_and_acceleration_bit_exact(gpu_ctx, nbody_oracle):
    app, (pos, mass, vel, acc) = _app(gpu_ctx, 1000, seed=2)
    app.run("movement")
    app.run("acceleration")
    app.download()
    nbody_oracle.axpy(pos, vel, nbody_oracle.dt)
    nbody_oracle.axpy(vel, acc, nbody_oracle.dt)
    assert np.array_equal(app.pos.view(np.uint32), pos.view(np.uint32))
    assert np.array_equal(app.vel.view(np.uint32), vel.view(np.uint32))


@pytest.mark.parametrize("n", [1, 5, 1000, 4099])
def test_streaming_tail_sizes_bit_exact(gpu_ctx, nbody_oracle, n):
    app, (pos, mass, vel, acc) = _app(gpu_ctx, n, seed=4)
    app.run("movement")
    app.download()
    nbody_oracle.axpy(pos, vel, nbody_oracle.dt)
    assert np.array_equal(app.pos.view(np.uint32), pos.view(np.uint32))


def test_one_tick_config0(gpu_ctx, nbody_oracle):
    """configs[0]: 1 000 bodies, EVERYTHING_HEAVY; one tick in schedule order G -> M -> A."""
    app, (pos, mass, vel, acc) = _app(gpu_ctx, 1000, seed=1, scene=scenes.EVERYTHING_HEAVY)
    app.tick(1)
    app.download()
    nbody_oracle.run_sequential(pos, mass, vel, acc, 1)
    assert np.array_equal(app.pos.view(np.uint32), pos.view(np.uint32)), "positions after tick 1 must be bit-exact"
    assert rel_err_rows(app.acc, acc).max() <= ACC_TOL
    assert rel_err_rows(app.vel, vel).max() <= ACC_TOL


@pytest.mark.parametrize("use_graph", [False, True])
def test_100_ticks_config0_drift(nbody_oracle, use_graph):
    """configs[0] over 100 ticks (MINIMUM_NUMBER_OF_TICKS, crates/ecs_bench_suite/benches/n_body.rs:4).
    Drift bound: max_i ||p_gpu - p_ref|| / max(||p_ref||, 1 m) <= 1e-4, same for velocity."""
    from recs_b200 import GpuContext

    with GpuContext(device=0, use_cuda_graph=use_graph) as ctx:
        app, (pos, mass, vel, acc) = _app(ctx, 1000, seed=1, scene=scenes.EVERYTHING_HEAVY)
        app.tick(100)
        app.download()
    nbody_oracle.run_sequential(pos, mass, vel, acc, 100)
    assert rel_err_rows(app.pos, pos, floor=1.0).max() <= 1e-4
    assert rel_err_rows(app.vel, vel, floor=1.0).max() <= 1e-4
    assert rel_err_rows(app.acc, acc).max() <= 1e-3  # accelerations of drifted positions


def test_graph_and_eager_ticks_agree_bitwise():
    from recs_b200 import GpuContext

    outs = []
    for use_graph in (False, True):
        with GpuContext(device=0, use_cuda_graph=use_graph) as ctx:
            app, _ = _app(ctx, 3000, seed=5)
            app.tick(7)
            app.download()
            outs.append((app.pos.copy(), app.vel.copy(), app.acc.copy()))
    for a, b in zip(*outs):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_gravity_65536_config1(gpu_ctx, nbody_oracle):
    """configs[1]: 65 536 bodies on one GPU; full-N check of sampled rows against the oracle
    (all j), f32 reference order and f64 truth."""
    n = 65536
    app, (pos, mass, vel, acc) = _app(gpu_ctx, n, seed=1)
    app.run("gravity")
    app.download()
    rows = np.random.default_rng(7).choice(n, size=256, replace=False)
    ref32, ref64 = nbody_oracle.gravity_sample(pos, mass, rows)
    err32 = rel_err_rows(app.acc[rows], ref32)
    err64 = rel_err_rows(app.acc[rows], ref64)
    ref_err64 = rel_err_rows(ref32, ref64)
    assert err32.max() <= ACC_TOL, f"vs reference-order f32: {err32.max():.3e}"
    assert err64.max() <= ACC_TOL, f"vs f64 truth: {err64.max():.3e}"
    # the tiled GPU sum should not be further from the truth than the reference's own sum
    assert err64.max() <= max(ref_err64.max() * 2, 2e-6)


def test_missing_column_is_reported(gpu_ctx):
    from recs_b200 import RecsGpuError, _lib
    from recs_b200.nbody import ACCELERATION, MASS, POSITION

    sid = gpu_ctx.create_system(_lib.SYS_NBODY_GRAVITY, [POSITION, ACCELERATION, MASS])
    gpu_ctx.set_archetypes(sid, [1])
    gpu_ctx.set_query_archetypes(sid, [1])
    with pytest.raises(RecsGpuError) as e:
        gpu_ctx.run_system(sid)
    assert e.value.code == 5  # RECS_ERR_MISSING_PARAMETER (SystemError::MissingParameter)
    # errors are sticky until cleared
    with pytest.raises(RecsGpuError):
        gpu_ctx.synchronize()
    gpu_ctx.clear_error()
    gpu_ctx.synchronize()


def test_invalidate_makes_binding_stale(gpu_ctx):
    from recs_b200 import RecsGpuError

    app, _ = _app(gpu_ctx, 100, seed=1)
    gpu_ctx.invalidate(1)
    with pytest.raises(RecsGpuError) as e:
        app.run("movement")
    assert e.value.code == 6  # RECS_ERR_STALE_BINDING


def test_fused_tick_tail_matches_separate_systems_bitwise(nbody_oracle):
    """recs_gpu_tick fuses movement + acceleration into the gravity reduce kernel when the schedule
    order is gravity -> movement -> acceleration; running the three systems one by one must give
    bit-identical columns (and so must a different, unfused order of the same tick)."""
    from recs_b200 import GpuContext

    outs = []
    for mode in ("tick", "separate"):
        with GpuContext(device=0, use_cuda_graph=False) as ctx:
            app, _ = _app(ctx, 5000, seed=7)
            for _ in range(3):
                if mode == "tick":
                    app.tick(1)
                else:
                    for name in ("gravity", "movement", "acceleration"):
                        app.run(name)
            app.download()
            outs.append((app.pos.copy(), app.vel.copy(), app.acc.copy()))
    for a, b in zip(*outs):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


@pytest.mark.parametrize("order", [("gravity", "acceleration", "movement"), ("movement", "acceleration", "gravity")])
def test_other_tick_orders_follow_the_oracle(gpu_ctx, nbody_oracle, order):
    """The demo's system set yields gravity -> acceleration -> movement and the schedule test's triple
    yields movement -> acceleration -> gravity: the GPU chain must follow whatever order
    the schedule produced."""
    from oracle.nbody import SYS_ACCELERATION, SYS_GRAVITY, SYS_MOVEMENT

    code = {"gravity": SYS_GRAVITY, "movement": SYS_MOVEMENT, "acceleration": SYS_ACCELERATION}
    app, (pos, mass, vel, acc) = _app(gpu_ctx, 2000, seed=9)
    app.set_tick_order(order)
    app.tick(3)
    app.download()
    nbody_oracle.run_sequential(pos, mass, vel, acc, 3, order=tuple(code[o] for o in order))
    assert rel_err_rows(app.pos, pos, floor=1.0).max() <= 1e-4
    assert rel_err_rows(app.vel, vel, floor=1.0).max() <= 1e-4
    assert rel_err_rows(app.acc, acc).max() <= 1e-4


@pytest.mark.parametrize("n", [24576, 24577])
def test_kernel_shape_switch_sizes(gpu_ctx, nbody_oracle, n):
    """Sizes on both sides of the small-problem / default kernel-shape switch."""
    app, (pos, mass, vel, acc) = _app(gpu_ctx, n, seed=3)
    app.run("gravity")
    app.download()
    rows = np.random.default_rng(1).choice(n, size=64, replace=False)
    ref32, _ = nbody_oracle.gravity_sample(pos, mass, rows, f64=False)
    assert rel_err_rows(app.acc[rows], ref32).max() <= ACC_TOL
