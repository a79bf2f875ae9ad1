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
    start = [a.copy() for a in (pos, mass, vel, acc)]
    nbody_oracle.run_sequential(pos, mass, vel, acc, 100)
    assert rel_err_rows(app.pos, pos, floor=1.0).max() <= 1e-4
    assert rel_err_rows(app.vel, vel, floor=1.0).max() <= 1e-4
    assert rel_err_rows(app.acc, acc).max() <= 1e-3  # accelerations of drifted positions
    # relative total-energy drift, beside the f64 run of the same integration.  The integrator (semi-implicit
    # Euler at dt = 1/20000) does not conserve energy exactly; what is asserted is that f32 on the GPU adds nothing visible
    # to what the f64 run and the reference-order f32 run show.
    e0 = sum(nbody_oracle.total_energy(start[0], start[1], start[2]))
    p64, v64, a64 = (start[i].astype(np.float64) for i in (0, 2, 3))
    nbody_oracle.run_f64(p64, start[1], v64, a64, 100)
    drift = {name: abs(sum(nbody_oracle.total_energy(p, start[1], v)) - e0) / abs(e0)
             for name, (p, v) in {"gpu": (app.pos, app.vel), "ref32": (pos, vel), "f64": (p64, v64)}.items()}
    print("energy drift 1000 x 100:", drift)
    assert abs(drift["gpu"] - drift["f64"]) <= 1e-6 + 2.0 * abs(drift["ref32"] - drift["f64"])


@pytest.mark.parametrize("n,ticks", [(1, 3), (2, 4), (31, 5), (33, 3), (255, 4), (257, 7), (1000, 100), (1000, 101), (1536, 6), (2048, 9), (3000, 5), (4096, 4)])
def test_resident_ticks_equal_single_ticks(monkeypatch, n, ticks):
    """Small single-archetype worlds run `recs_gpu_tick(n)` as one cooperative launch (nbody_resident_ticks_kernel) after
    the first tick.  Same arithmetic, same summation tree: every column must equal the tick-by-tick chain bit for bit,
    and the mirror it leaves behind must be the one the next ordinary tick needs."""
    from recs_b200 import GpuContext

    outs, launches = [], []
    for resident in (True, False):
        if not resident:
            monkeypatch.setenv("RECS_GPU_RESIDENT_MAX_ROWS", "0")
        with GpuContext(device=0, use_cuda_graph=False) as ctx:
            app, host = _app(ctx, n, seed=9)
            if n >= 4:  # a coincident pair and a sub-epsilon neighbour: the masked pairs of the reference
                app.pos[1] = app.pos[0]
                app.pos[3] = app.pos[2] + np.float32(1e-5)
                app.upload()
            app.tick(ticks)
            launches.append(ctx.timing()["kernel_launches"])
            assert resident or launches[-1] >= 2 * ticks
            app.tick(1)       # an ordinary tick on the mirror the resident launch left
            app.tick(ticks)   # and a resident launch on a mirror that needs no pack
            app.download()
            outs.append((app.pos.copy(), app.vel.copy(), app.acc.copy()))
    for a, b in zip(*outs):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    assert launches[0] <= 4, launches  # pack + tick 1 (2 launches) + ONE launch for the rest


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
@pytest.mark.parametrize("n", [700, 3000, 6000])
def test_random_call_sequences_do_not_depend_on_the_tick_path(monkeypatch, seed, n):
    """The same random sequence of calls -- ticks of several lengths, single systems, host writes of positions,
    velocities or masses in between, a changed tick order -- on three contexts: resident ticks + CUDA graphs (the defaults),
    graphs only, neither.  All the bookkeeping that decides when the mirror is re-packed, when a captured graph is still
    valid and when the resident launch may run must be invisible: columns bit-identical at every download."""
    from recs_b200 import GpuContext

    rng = np.random.default_rng(100 * seed + n)
    ops = []
    for _ in range(14):
        kind = rng.choice(["tick", "tick", "tick", "system", "write", "order", "download"])
        if kind == "tick":
            ops.append(("tick", int(rng.choice([1, 2, 3, 4, 7, 20]))))
        elif kind == "system":
            ops.append(("system", str(rng.choice(["gravity", "movement", "acceleration"]))))
        elif kind == "write":
            ops.append(("write", str(rng.choice(["pos", "vel", "mass"])), int(rng.integers(0, n)), float(rng.uniform(0.5, 2.0))))
        elif kind == "order":
            ops.append(("order", tuple(rng.permutation(["gravity", "movement", "acceleration"]))))
        else:
            ops.append(("download",))
    ops.append(("tick", 5))
    ops.append(("download",))

    def run(resident, graph):
        if resident:
            monkeypatch.delenv("RECS_GPU_RESIDENT_MAX_ROWS", raising=False)
        else:
            monkeypatch.setenv("RECS_GPU_RESIDENT_MAX_ROWS", "0")
        snapshots = []
        with GpuContext(device=0, use_cuda_graph=graph) as ctx:
            app, _ = _app(ctx, n, seed=40 + seed)
            for op in ops:
                if op[0] == "tick":
                    app.tick(op[1])
                elif op[0] == "system":
                    app.run(op[1])
                elif op[0] == "write":
                    app.download()
                    col = {"pos": app.pos, "vel": app.vel, "mass": app.mass}[op[1]]
                    col[op[2]] = col[op[2]] * np.float32(op[3])
                    app.upload()
                elif op[0] == "order":
                    app.set_tick_order(op[1])
                else:
                    app.download()
                    snapshots.append((app.pos.copy(), app.vel.copy(), app.acc.copy()))
        return snapshots

    base = run(True, True)
    assert base and all(np.all(np.isfinite(x)) for snap in base for x in snap)
    for other in (run(False, True), run(False, False)):
        assert len(other) == len(base)
        for a, b in zip(base, other):
            for x, y in zip(a, b):
                assert np.array_equal(x.view(np.uint32), y.view(np.uint32))


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


@pytest.mark.parametrize("n", [5120, 5121, 16384, 22528, 22529, 49153])
def test_kernel_shape_switch_sizes(gpu_ctx, nbody_oracle, n):
    """Sizes on both sides of the small / medium / default kernel-shape switches."""
    app, (pos, mass, vel, acc) = _app(gpu_ctx, n, seed=3)
    app.run("gravity")
    app.download()
    rows = np.random.default_rng(1).choice(n, size=64, replace=False)
    ref32, _ = nbody_oracle.gravity_sample(pos, mass, rows, f64=False)
    assert rel_err_rows(app.acc[rows], ref32).max() <= ACC_TOL


# ---- kernel shapes (recs_b200/csrc/gravity.cu, kVariants) --------------------------------------------------------
RESULT_VARIANTS = [0, 1, 2, 3, 4, 5, 6, 7, 8]  # every shape of the shipped library (ablations exist in tools/ builds only)


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


@pytest.mark.parametrize("n,ticks", [(16384, 100)])
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


# ---- the epsilon boundary (crates/n-body/src/lib.rs:116-120) -------------------------------------------------------
def _epsilon_boundary_offsets(per_class=24):
    """Offsets d whose reference d2 = (dx*dx + dy*dy) + dz*dz (separate roundings) is exactly epsilon, one ulp below
    and one ulp above it, preferring those where an FMA-accumulated d2 would decide `d2 <= epsilon` differently."""
    eps = np.float32(1.1920929e-7)
    targets = {"below": np.nextafter(eps, np.float32(0)), "equal": eps, "above": np.nextafter(eps, np.float32(1))}
    picked = {k: [] for k in targets}
    flipped = {k: 0 for k in targets}
    rng = np.random.default_rng(3)
    for _ in range(6000):
        dx, dy = (rng.uniform(0.2, 0.6, size=2) * np.sqrt(eps)).astype(np.float32)
        s = np.float32(np.float32(dx * dx) + np.float32(dy * dy))
        dz0 = np.float32(np.sqrt(np.float64(eps) - np.float64(s)))
        for k in range(-6, 7):
            dz = dz0
            for _ in range(abs(k)):
                dz = np.nextafter(dz, np.float32(1 if k > 0 else 0))
            d2 = np.float32(s + np.float32(dz * dz))
            fused = np.float32(np.float64(dz) * np.float64(dz)
                               + np.float64(np.float32(np.float64(dy) * np.float64(dy) + np.float64(np.float32(dx * dx)))))
            for name, t in targets.items():
                if d2 == t:
                    flips = bool((fused <= eps) != (d2 <= eps))
                    if flips and flipped[name] < per_class // 2:
                        picked[name].insert(0, (dx, dy, dz))
                        flipped[name] += 1
                    elif len(picked[name]) < per_class:
                        picked[name].append((dx, dy, dz))
    out = []
    for name in targets:
        assert len(picked[name]) >= 8, name
        out += picked[name][:per_class]
    return np.array(out, dtype=np.float32), sum(flipped.values())


@pytest.mark.parametrize("variant", RESULT_VARIANTS)
def test_epsilon_boundary_mask_is_bit_exact(monkeypatch, nbody_oracle, variant):
    """Pairs whose d2 is epsilon, epsilon - 1 ulp and epsilon + 1 ulp: every kernel shape must take the reference's
    mask decision (`if distance_squared <= f32::EPSILON { zero }`), which it forms from the UNFUSED d2.  One heavy
    body at the origin, massless probes at -d (so `to_body` is exactly d): a probe's acceleration is exactly zero
    iff the pair is masked."""
    from recs_b200.nbody import NBodyGpuApp

    offsets, n_flipping = _epsilon_boundary_offsets()
    assert n_flipping > 0, "no case where a fused d2 would decide differently: the test would prove nothing"
    n = 1 + len(offsets)
    pos = np.zeros((n, 3), np.float32)
    pos[1:] = -offsets
    mass = np.zeros(n, np.float32)
    mass[0] = 5.0e20
    vel = np.zeros((n, 3), np.float32)
    acc = np.ones((n, 3), np.float32)
    with _forced_variant_ctx(monkeypatch, variant) as ctx:
        app = NBodyGpuApp(ctx, pos.copy(), mass, vel, acc)
        app.upload()
        app.run("gravity")
        app.download()
    ref = nbody_oracle.gravity(pos, mass)
    masked_ref = ~np.any(ref[1:] != 0, axis=1)
    masked_gpu = ~np.any(app.acc[1:] != 0, axis=1)
    assert masked_ref.any() and (~masked_ref).any()
    assert np.array_equal(masked_gpu, masked_ref), f"mask decisions differ on probes {np.nonzero(masked_gpu != masked_ref)[0]}"
    assert rel_err_rows(app.acc, ref).max() <= ACC_TOL


# ---- the summation tree: results do not depend on grid size, launch split or rank count -----------------------------
@pytest.mark.parametrize("n", [3000, 20000, 70000])
def test_result_does_not_depend_on_grid_or_rank_split(monkeypatch, n):
    """Tile sums are added in a fixed tree (chunks of consecutive tiles, then chunks in ascending order; kernels.cuh),
    so neither the number of persistent CTAs the (i-block, tile) units are dealt to, nor the two-phase walk of a rank of a
    sharded run (own slice first, then the other ranks' slices in rotated order, mirror padded to R equal slices), nor
    which of the i-packed kernel shapes computes a row may change a bit of any row."""
    from recs_b200 import GpuContext
    from recs_b200.nbody import NBodyGpuApp

    def run(env):
        for key in ("RECS_GPU_GRAVITY_GRID", "RECS_GPU_GRAVITY_EMULATE_RANK", "RECS_GPU_GRAVITY_VARIANT", "RECS_GPU_GRAVITY_NO_LAST_SPLIT"):
            monkeypatch.delenv(key, raising=False)
        for key, value in env.items():
            monkeypatch.setenv(key, value)
        pos, mass, vel, acc = _bodies_with_close_pairs(n, seed=13)
        with GpuContext(device=0, use_cuda_graph=False) as ctx:
            app = NBodyGpuApp(ctx, pos, mass, vel, acc)
            app.upload()
            app.tick(2)  # the second tick reads the mirror the fused tail refreshed
            app.download()
            return app.pos.copy(), app.vel.copy(), app.acc.copy()

    base = run({})
    assert np.all(np.isfinite(base[2]))
    for env in ({"RECS_GPU_GRAVITY_GRID": "1"}, {"RECS_GPU_GRAVITY_GRID": "7"}, {"RECS_GPU_GRAVITY_GRID": "100"},
                {"RECS_GPU_GRAVITY_GRID": "333"}, {"RECS_GPU_GRAVITY_EMULATE_RANK": "2,0"},
                {"RECS_GPU_GRAVITY_EMULATE_RANK": "2,1"}, {"RECS_GPU_GRAVITY_EMULATE_RANK": "4,2"},
                {"RECS_GPU_GRAVITY_EMULATE_RANK": "8,7", "RECS_GPU_GRAVITY_GRID": "61"},
                {"RECS_GPU_GRAVITY_EMULATE_RANK": "3,1"}) + (
                # the i-packed shapes (1536-row and 256-row i-blocks, unroll 16 and 8) share arithmetic and tree
                ({"RECS_GPU_GRAVITY_VARIANT": "0"}, {"RECS_GPU_GRAVITY_VARIANT": "3", "RECS_GPU_GRAVITY_GRID": "97"},
                 {"RECS_GPU_GRAVITY_VARIANT": "8"},
                 # at 70 000 rows the partly filled last i-block is dealt by itself (phases of its own, chunk sums in a
                 # separate launch); switching that off, alone and inside an emulated rank, must change nothing
                 {"RECS_GPU_GRAVITY_NO_LAST_SPLIT": "1"}, {"RECS_GPU_GRAVITY_NO_LAST_SPLIT": "1", "RECS_GPU_GRAVITY_EMULATE_RANK": "4,1"})
                if n > 5120 else ()):
        got = run(env)
        for a, b, what in zip(got, base, ("pos", "vel", "acc")):
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), f"{what} differs with {env}"


# ---- several body archetypes (the demo has planets and light-emitting suns, n_body_demo.rs:28-52) ------------------
def _two_archetype_system(ctx, parts):
    """Registers the three n-body systems over two body archetypes (2 and 5); returns ids and host columns."""
    from recs_b200 import _lib
    from recs_b200.nbody import ACCELERATION, MASS, POSITION, VELOCITY

    systems = {
        "movement": ctx.create_system(_lib.SYS_NBODY_MOVEMENT, [POSITION, VELOCITY]),
        "acceleration": ctx.create_system(_lib.SYS_NBODY_ACCELERATION, [VELOCITY, ACCELERATION]),
        "gravity": ctx.create_system(_lib.SYS_NBODY_GRAVITY, [POSITION, ACCELERATION, MASS]),
    }
    cols = {}
    for arch, (pos, mass, vel, acc) in zip((2, 5), parts):
        cols[arch] = tuple(np.ascontiguousarray(a.copy()) for a in (pos, mass, vel, acc))
        for comp, arr in zip((POSITION, MASS, VELOCITY, ACCELERATION), cols[arch]):
            ctx.bind_column(arch, comp, arr)
            ctx.upload(arch, comp)
    for sid in systems.values():
        ctx.set_archetypes(sid, [2, 5])
    ctx.set_query_archetypes(systems["gravity"], [2, 5])
    return systems, cols


def _download_two(ctx, cols):
    from recs_b200.nbody import ACCELERATION, POSITION, VELOCITY

    for arch in cols:
        for comp in (POSITION, VELOCITY, ACCELERATION):
            ctx.download(arch, comp)
    return [np.concatenate([cols[2][i], cols[5][i]]) for i in (0, 2, 3)]


def test_fused_tail_with_two_body_archetypes(nbody_oracle):
    """gravity must finish for ALL outer archetypes before movement advances any body (gravity reads Position,
    movement writes it): the fused tick over two body archetypes equals the three systems run one by one, bit for
    bit, and follows the oracle run on the concatenated bodies."""
    from recs_b200 import GpuContext

    pos, mass, vel, acc = scenes.spawn_bodies(scenes.all_heavy_random_cube_with_bodies(5000), 17)
    vel *= np.float32(3.0e4)  # fast bodies: positions move by ~1e-1 per tick, a stale read is far outside tolerance
    split = 1800
    parts = [tuple(a[:split] for a in (pos, mass, vel, acc)), tuple(a[split:] for a in (pos, mass, vel, acc))]
    outs = []
    for mode in ("tick", "separate"):
        with GpuContext(device=0, use_cuda_graph=False) as ctx:
            systems, cols = _two_archetype_system(ctx, parts)
            ctx.set_tick_order([systems["gravity"], systems["movement"], systems["acceleration"]])
            for _ in range(3):
                if mode == "tick":
                    ctx.tick(1)
                else:
                    for name in ("gravity", "movement", "acceleration"):
                        ctx.run_system(systems[name])
            outs.append(_download_two(ctx, cols))
    for a, b in zip(*outs):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    ref = [a.copy() for a in (pos, mass, vel, acc)]
    nbody_oracle.run_sequential(ref[0], ref[1], ref[2], ref[3], 3)
    assert rel_err_rows(outs[0][0], ref[0], floor=1.0).max() <= 1e-4
    assert rel_err_rows(outs[0][1], ref[2], floor=1.0).max() <= 1e-4
    assert rel_err_rows(outs[0][2], ref[3]).max() <= 1e-4


def test_captured_tick_sees_positions_written_between_ticks(nbody_oracle):
    """A captured tick (CUDA graph) that found the gravity mirror current holds no pack kernel.  When something outside
    the chain writes Position between ticks -- here recs_gpu_run_system(movement) and a host upload -- the next
    recs_gpu_tick(n > 1) must not replay it on the stale mirror."""
    from recs_b200 import GpuContext
    from recs_b200.nbody import POSITION, BODY_ARCHETYPE

    outs = []
    for use_graph in (True, False):
        with GpuContext(device=0, use_cuda_graph=use_graph) as ctx:
            app, _ = _app(ctx, 3000, seed=21)
            app.vel *= np.float32(3.0e4)
            app.upload()
            app.tick(4)  # captures (with graphs on): the fused order leaves the mirror current, no pack in the graph
            app.run("movement")  # writes Position outside the chain
            app.tick(3)
            app.download()
            app.pos += np.float32(0.25)  # the host moves every body and uploads
            ctx.upload(BODY_ARCHETYPE, POSITION)
            app.tick(3)
            app.download()
            outs.append((app.pos.copy(), app.vel.copy(), app.acc.copy()))
    for a, b in zip(*outs):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


# ---- the gates of SURVEY 8(d) at the full sizes -------------------------------------------------------------------
def test_gravity_1M_sampled_rows(nbody_oracle):
    """configs[2] size on one GPU: 4 096 sampled rows of 1 048 576 bodies against the oracle (all j).
    The reference's own sequential f32 sum is up to ~1e-4 away from the f64 sum at this size (its rounding noise
    grows with N), so the literal gate `GPU vs reference-order f32 <= 1e-4` can fail on a few rows for which the
    REFERENCE is the inaccurate side.  Arbitration rule (stated in DESIGN.md 7 and reported by bench.py): a row passes
    if it is within 1e-4 of the reference-order f32 result, or the reference-order result is itself further than
    that row's GPU error from the f64 truth and the GPU is within 1e-5 of the truth."""
    from recs_b200 import GpuContext

    n = 1_048_576
    with GpuContext(device=0) as ctx:
        app, (pos, mass, vel, acc) = _app(ctx, n, seed=1)
        app.run("gravity")
        app.download()
        got = app.acc.copy()
    rows = np.sort(np.random.default_rng(11).choice(n, size=4096, replace=False))
    ref32, ref64 = nbody_oracle.gravity_sample(pos, mass, rows)
    err32 = rel_err_rows(got[rows], ref32)
    err64 = rel_err_rows(got[rows], ref64)
    ref_err64 = rel_err_rows(ref32, ref64)
    assert err64.max() <= 1e-5, f"GPU vs f64 truth: {err64.max():.3e}"
    over = err32 > ACC_TOL
    # every row over the literal tolerance must be one where the reference-order sum is the inaccurate side
    assert np.all(ref_err64[over] > err64[over]), f"{over.sum()} rows over 1e-4 not explained by the reference's own rounding"
    assert over.mean() <= 0.01, f"{over.sum()} of {rows.size} rows over 1e-4 vs the reference-order f32 sum"


def test_drift_65536_bodies_100_ticks(nbody_oracle):
    """north_star: "a stated drift bound must hold over 100 steps" -- at configs[1] size, 65 536 bodies x 100 full ticks.

    The reference trajectory comes from the reference's own arithmetic and summation order executed on the device
    (ReferenceOrderNBodyApp: bit-identical to the CPU oracle, asserted here for the first ticks against the oracle's
    worker pool; 100 ticks of the CPU oracle would take minutes).  An N-body system amplifies rounding differences
    through close encounters, for the reference as for anything else, so the bound is stated with the reference's own
    sensitivity as yardstick: the same reference arithmetic with the nested query iterated in REVERSE order (the
    reference's archetype iteration order is itself unspecified, crates/ecs/src/lib.rs:842-865).

    Bound: max_i |p_gpu - p_ref| / max(|p_ref|, 1 m) <= 1e-4, and for velocity <= max(1e-4, 2 x what reversing the
    reference's own summation order does)."""
    import os

    from recs_b200 import GpuContext
    from recs_b200.nbody import NBodyGpuApp, ReferenceOrderNBodyApp

    n, ticks, pin_ticks = 65536, 100, 2
    start = scenes.spawn_bodies(scenes.all_heavy_random_cube_with_bodies(n), 8)

    def run(app_type, cols, n_ticks):
        cols = [np.ascontiguousarray(a.copy()) for a in cols]
        with GpuContext(device=0) as ctx:
            app = app_type(ctx, *cols)
            app.upload()
            app.tick(n_ticks)
            app.download()
        return cols

    # the device-run reference trajectory is the oracle's, bit for bit
    pinned = run(ReferenceOrderNBodyApp, start, pin_ticks)
    want = [a.copy() for a in start]
    threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    nbody_oracle.run_pool(want[0], want[1], want[2], want[3], pin_ticks, threads)
    for got, ref in zip(pinned, want):
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))

    gpu = run(NBodyGpuApp, start, ticks)
    ref = run(ReferenceOrderNBodyApp, start, ticks)
    rev = [a[::-1] for a in run(ReferenceOrderNBodyApp, [a[::-1] for a in start], ticks)]
    pos_err = rel_err_rows(gpu[0], ref[0], floor=1.0).max()
    vel_err = rel_err_rows(gpu[2], ref[2], floor=1.0).max()
    ref_pos_sens = rel_err_rows(rev[0], ref[0], floor=1.0).max()
    ref_vel_sens = rel_err_rows(rev[2], ref[2], floor=1.0).max()
    print(f"drift 65536 x 100: pos {pos_err:.3e} vel {vel_err:.3e}; reference order sensitivity pos {ref_pos_sens:.3e} vel {ref_vel_sens:.3e}")
    assert pos_err <= 1e-4
    assert vel_err <= max(1e-4, 2.0 * ref_vel_sens), f"velocity drift {vel_err:.3e}, reference's own order sensitivity {ref_vel_sens:.3e}"
    # the bulk of the bodies is nowhere near the bound
    assert np.quantile(rel_err_rows(gpu[2], ref[2], floor=1.0), 0.999) <= 1e-5
    # relative total-energy drift over the 100 ticks (f64 from the columns), beside the reference-order run's
    e0 = sum(nbody_oracle.total_energy(start[0], start[1], start[2]))
    e_gpu = sum(nbody_oracle.total_energy(gpu[0], start[1], gpu[2]))
    e_ref = sum(nbody_oracle.total_energy(ref[0], start[1], ref[2]))
    print(f"energy drift 65536 x 100: gpu {abs(e_gpu - e0) / abs(e0):.3e}  reference order {abs(e_ref - e0) / abs(e0):.3e}")
    assert abs(e_gpu - e_ref) / abs(e0) <= 1e-6


# ---- every scene the reference defines (crates/n-body/src/scenes.rs:10-47, 179-206) -----------------------------------
def _reference_scenes():
    from dataclasses import replace

    one_heavy = lambda seed: tuple(np.concatenate([a, b]) for a, b in zip(          # the demo's set-up (n_body_demo.rs:28-52):
        scenes.spawn_bodies(scenes.ONE_HEAVY_BODY, seed), scenes.spawn_bodies(scenes.EVERYTHING_LIGHT, seed + 1)))  # a sun among light bodies
    return {
        "EVERYTHING_HEAVY": lambda seed: scenes.spawn_bodies(scenes.EVERYTHING_HEAVY, seed),
        "EVERYTHING_LIGHT": lambda seed: scenes.spawn_bodies(scenes.EVERYTHING_LIGHT, seed),
        "ONE_HEAVY_BODY+LIGHT": one_heavy,
        "SMALL_HEAVY_CLUSTERS": lambda seed: scenes.spawn_clusters(scenes.SMALL_HEAVY_CLUSTERS, seed),
        "SMALL_LIGHT_CLUSTERS": lambda seed: scenes.spawn_clusters(scenes.SMALL_LIGHT_CLUSTERS, seed),
        "HUGE_HEAVY_CLUSTERS": lambda seed: scenes.spawn_clusters(scenes.HUGE_HEAVY_CLUSTERS, seed),
        "HUGE_LIGHT_CLUSTERS": lambda seed: scenes.spawn_clusters(scenes.HUGE_LIGHT_CLUSTERS, seed),
        "HUGE_HEAVY_CLUSTERS x 6000": lambda seed: scenes.spawn_clusters(
            replace(scenes.HUGE_HEAVY_CLUSTERS, scene=replace(scenes.HUGE_HEAVY_CLUSTERS.scene, body_count=6000)), seed),
    }


@pytest.mark.parametrize("name", list(_reference_scenes()))
def test_every_reference_scene(nbody_oracle, name):
    """The reference's own scenes -- uniform cubes heavy and light, one heavy body among light ones, four clusters 100 m and
    1 km apart -- through (a) the hand-written tick: positions bit-exact after tick 1, accelerations within 1e-4 of the
    reference-order sum, 20 ticks (one resident launch for the small ones) within the drift bound; (b) the generic
    reference-order system: accelerations equal to the oracle bit for bit (its fast fold must cope with the light scenes'
    tiny G*m and the clusters' range of distances)."""
    from recs_b200 import GpuContext
    from recs_b200.nbody import NBodyGpuApp, ReferenceOrderNBodyApp

    start = _reference_scenes()[name](31)
    n = start[0].shape[0]
    want = [a.copy() for a in start]
    nbody_oracle.run_sequential(want[0], want[1], want[2], want[3], 1)
    for app_type in (NBodyGpuApp, ReferenceOrderNBodyApp):
        cols = [a.copy() for a in start]
        with GpuContext(device=0) as ctx:
            app = app_type(ctx, *cols)
            app.upload()
            app.tick(1)
            app.download()
        assert np.array_equal(cols[0].view(np.uint32), want[0].view(np.uint32)), "positions after tick 1"
        if app_type is ReferenceOrderNBodyApp:
            assert np.array_equal(cols[3].view(np.uint32), want[3].view(np.uint32)), "reference-order accelerations"
            assert np.array_equal(cols[2].view(np.uint32), want[2].view(np.uint32))
        else:
            assert rel_err_rows(cols[3], want[3]).max() <= ACC_TOL
    ticks = 20
    later = [a.copy() for a in start]
    nbody_oracle.run_sequential(later[0], later[1], later[2], later[3], ticks)
    cols = [a.copy() for a in start]
    with GpuContext(device=0) as ctx:
        app = NBodyGpuApp(ctx, *cols)
        app.upload()
        app.tick(ticks)
        app.download()
    assert rel_err_rows(cols[0], later[0], floor=1.0).max() <= 1e-4
    assert rel_err_rows(cols[2], later[2], floor=1.0).max() <= 1e-4
    assert n >= 1000
