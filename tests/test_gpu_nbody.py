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


@pytest.mark.parametrize("n", [16384, 16385, 49152, 49153])
def test_kernel_shape_switch_sizes(gpu_ctx, nbody_oracle, n):
    """Sizes on both sides of the small / medium / default kernel-shape switches."""
    app, (pos, mass, vel, acc) = _app(gpu_ctx, n, seed=3)
    app.run("gravity")
    app.download()
    rows = np.random.default_rng(1).choice(n, size=64, replace=False)
    ref32, _ = nbody_oracle.gravity_sample(pos, mass, rows, f64=False)
    assert rel_err_rows(app.acc[rows], ref32).max() <= ACC_TOL


# ---- kernel shapes (recs_b200/csrc/gravity.cu, kVariants) --------------------------------------------------------
RESULT_VARIANTS = [0, 1, 2, 3, 4, 5, 6, 7]  # 8 and 9 are timing-only ablations with wrong results by design


def _forced_variant_ctx(monkeypatch, variant):
    from recs_b200 import GpuContext

    monkeypatch.setenv("RECS_GPU_GRAVITY_VARIANT", str(variant))
    return GpuContext(device=0, use_cuda_graph=False)


def _bodies_with_close_pairs(n, seed):
    pos, mass, vel, acc = scenes.spawn_bodies(scenes.all_heavy_random_cube_with_bodies(n), seed)
    if n >= 64:
        k = n // 8
        pos[k : 2 * k] = pos[:k]  # coincident twins
        pos[2 * k : 3 * k] = pos[:k] + np.array([1e-4, 0, 0], np.float32)  # sub-epsilon neighbours
        pos[3 * k : 4 * k] = pos[:k] + np.array([0, 7e-4, 0], np.float32)  # just above epsilon where f32 resolves it
    return np.ascontiguousarray(pos), mass, vel, acc


@pytest.mark.parametrize("variant", RESULT_VARIANTS)
@pytest.mark.parametrize("n", [5, 257, 1537, 4097])
def test_every_kernel_shape_matches_oracle(monkeypatch, nbody_oracle, variant, n):
    """Each compiled shape, forced for every size (i-block / tile edges included), with coincident,
    sub-epsilon and just-above-epsilon pairs spread over other tiles."""
    from recs_b200.nbody import NBodyGpuApp

    pos, mass, vel, acc = _bodies_with_close_pairs(n, seed=11)
    with _forced_variant_ctx(monkeypatch, variant) as ctx:
        app = NBodyGpuApp(ctx, pos.copy(), mass, vel, acc)
        app.upload()
        app.run("gravity")
        app.download()
    ref = nbody_oracle.gravity(pos, mass)
    assert np.all(np.isfinite(app.acc))
    assert rel_err_rows(app.acc, ref).max() <= ACC_TOL


def test_range_shift_is_bit_exact(monkeypatch):
    """Variant 0 works on positions scaled by 2^-32 and detects d2 <= epsilon through the overflow of r^-3;
    variant 4 is the same operation order on unscaled data with an explicit min-d2 detector.  Power-of-two
    scalings commute with every rounding involved, so wherever both take the exact path for the same tiles
    (coincident and sub-epsilon pairs; light bodies, so that G*m/d^3 stays below 2^32) they must agree bit for
    bit -- several ticks of the fused tail, which refreshes the shifted mirror, included."""
    from recs_b200.nbody import NBodyGpuApp

    outs = []
    for variant in (0, 4):
        scene = scenes.replace(scenes.EVERYTHING_LIGHT, body_count=20000)
        pos, mass, vel, acc = scenes.spawn_bodies(scene, 12)
        k = 2500
        pos[k : 2 * k] = pos[:k]  # coincident twins
        pos[2 * k : 3 * k] = pos[:k] + np.array([1e-4, 0, 0], np.float32)  # sub-epsilon neighbours
        with _forced_variant_ctx(monkeypatch, variant) as ctx:
            app = NBodyGpuApp(ctx, np.ascontiguousarray(pos), mass, vel, acc)
            app.upload()
            app.run("gravity")
            app.download()
            first = app.acc.copy()
            app.tick(3)
            app.download()
            outs.append((first, app.pos.copy(), app.vel.copy(), app.acc.copy()))
    assert np.all(np.isfinite(outs[0][0])) and np.any(outs[0][0] != 0)
    for a, b in zip(*outs):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


@pytest.mark.parametrize(
    "name,pos_scale,mass_lo,mass_hi",
    [
        ("light", 1.0, 100.0, 1000.0),  # EVERYTHING_LIGHT (crates/n-body/src/scenes.rs:23-27)
        ("dust", 1.0, 1e-30, 1e-28),  # G*m around 1e-40: denormal in the reference as well
        ("grains", 1e-3, 1e-12, 1e-9),
        ("stars", 1e12, 1e29, 1e31),
        ("galactic", 1e16, 1e36, 1e37),  # d2 up to ~1e37, close to the f32 limit
        ("superdense", 1e-2, 1e20, 1e21),  # G*m/d^3 >= 2^32: every tile takes the exact path
        ("mixed", 1.0, 1e-6, 1e24),
    ],
)
@pytest.mark.parametrize("variant", [0, 3])
def test_range_extremes(monkeypatch, nbody_oracle, variant, name, pos_scale, mass_lo, mass_hi):
    """The shifted kernels over the dynamic range of f32 inputs: no input may silently lose accuracy --
    whatever does not fit the shifted range must reach the exact path.  (Not covered: separations beyond
    1.8e19 m, where the reference's d2 overflows to +inf and it drops the pair; the shifted path keeps the pair's
    true, tiny contribution -- DESIGN.md 3.1.)"""
    from recs_b200.nbody import NBodyGpuApp

    n = 3000
    rng = np.random.default_rng(21)
    pos = (rng.uniform(-100.0, 100.0, size=(n, 3)) * pos_scale).astype(np.float32)
    if name == "mixed":
        mass = np.exp(rng.uniform(np.log(mass_lo), np.log(mass_hi), size=n)).astype(np.float32)
        mass[::17] = 0.0  # massless bodies are legal rows
    else:
        mass = rng.uniform(mass_lo, mass_hi, size=n).astype(np.float32)
    vel = np.zeros((n, 3), np.float32)
    acc = np.zeros((n, 3), np.float32)
    with _forced_variant_ctx(monkeypatch, variant) as ctx:
        app = NBodyGpuApp(ctx, pos.copy(), mass, vel, acc)
        app.upload()
        app.run("gravity")
        app.download()
    ref = nbody_oracle.gravity(pos, mass)
    ref64 = nbody_oracle.gravity_f64(pos, mass)
    assert np.all(np.isfinite(app.acc)) or not np.all(np.isfinite(ref))
    # judged like everywhere else: norm-wise per body against the reference-order f32 oracle, with the f64
    # sum as arbiter where the reference's own f32 result is further from the truth than 1e-4
    # ... and an absolute floor where the reference itself works in denormals: there `(G*m/d2)/|d|` is quantised
    # to 2^-149 before it multiplies to_body, so n * 2^-149 * max|to_body| is the reference's own error bound
    floor = n * 2.0**-149 * 2.0 * float(np.abs(pos).max()) / ACC_TOL
    err32 = rel_err_rows(app.acc, ref, floor=floor)
    err64 = rel_err_rows(app.acc, ref64, floor=floor)
    ok = (err32 <= ACC_TOL) | (err64 <= ACC_TOL)
    assert ok.all(), f"{name}: {np.count_nonzero(~ok)} rows off, worst {np.minimum(err32, err64).max():.3e}"


@pytest.mark.parametrize("n,ticks", [(16384, 100), (65536, 10)])
def test_drift_bound_at_larger_sizes(nbody_oracle, n, ticks):
    """SURVEY 8d drift gate beyond configs[0]: after `ticks` full ticks, max_i |p_gpu - p_ref| / max(|p_ref|, 1 m) <= 1e-4
    and the same for velocity, against the oracle run with the reference's worker-pool policy on all host threads
    (sizes chosen so that the CPU side stays within seconds)."""
    import os

    from recs_b200 import GpuContext

    with GpuContext(device=0) as ctx:
        app, (pos, mass, vel, acc) = _app(ctx, n, seed=8)
        app.tick(ticks)
        app.download()
    nbody_oracle.run_pool(pos, mass, vel, acc, ticks, max(os.cpu_count() or 1, 1))
    assert rel_err_rows(app.pos, pos, floor=1.0).max() <= 1e-4
    assert rel_err_rows(app.vel, vel, floor=1.0).max() <= 1e-4
