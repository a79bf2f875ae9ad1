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


def test_fast_fold_is_generated_only_when_the_source_defines_it(tmp_path, monkeypatch):
    """`bool <pair>_fast(...)` in the source switches the generated kernel to the speculative fold (fast body per tile, exact
    redo from the saved accumulator); without it the exact body alone is folded; a fast body with other parameters than
    `pair` is a compile error in the user's own source, not a silent fallback.  RECS_GPU_KERNEL_DUMP keeps source + cubin."""
    monkeypatch.setenv("RECS_GPU_KERNEL_DUMP", str(tmp_path))
    with_fast = GRAVITY + "\n// variant A\n"     # (distinct texts: compiled sources are cached per process)
    i, j = GRAVITY.index("bool pair_fast("), GRAVITY.index("void finish(")
    without = GRAVITY[:i] + GRAVITY[j:] + "\n// variant B\n"
    wrong = GRAVITY.replace("bool pair_fast(Total& total, const Position& position,", "bool pair_fast(Total& total, const Mass& position,") + "\n// variant C\n"
    st, log, cubin = compile_pair_kernel_system(with_fast, "Total", "pair", "finish", GRAVITY_PARAMS)
    assert st == 0 and cubin > 1000, log
    st, log, cubin = compile_pair_kernel_system(without, "Total", "pair", "finish", GRAVITY_PARAMS)
    assert st == 0 and cubin > 1000, log
    st, log, _ = compile_pair_kernel_system(wrong, "Total", "pair", "finish", GRAVITY_PARAMS)
    assert st == 12 and "pair_fast" in log
    dumped = sorted(tmp_path.glob("recs_generated_*.cu"))
    assert len(dumped) == 3 and len(list(tmp_path.glob("recs_generated_*.cubin"))) == 2
    texts = {("A" if "// variant A" in t else "B" if "// variant B" in t else "C"): t for t in (p.read_text() for p in dumped)}
    generated = lambda t: t[t.index("recs_generated_kernel.cu"):]
    assert "recs_ok = recs_ok & pair_fast(recs_acc0" in generated(texts["A"]) and "recs_acc0 = recs_saved" in generated(texts["A"])
    assert "pair_fast" not in generated(texts["B"]) and "pair(recs_acc0" in generated(texts["B"])


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


def _strip_fast(source: str) -> str:
    """The same system without its `pair_fast` body: the generated kernel then folds with the exact body only."""
    i = source.index("bool pair_fast(")
    j = source.index("void finish(")
    return source[:i] + source[j:]


@pytest.mark.gpu
@pytest.mark.parametrize("what", ["denormal_gm", "huge_distances", "tiny_distances", "mixed", "nan_and_inf"])
def test_fast_fold_redoes_tiles_with_out_of_range_pairs(gpu_ctx, nbody_oracle, what):
    """The speculative fast fold (pair_fast: branch-free division / square-root sequences, valid for operands within
    2^-63 .. 2^63) must hand every tile holding an out-of-range pair to the exact body: denormal G*m, distances beyond
    2^31.5 m (d2 > 2^63), d2 below 2^-63 but above zero is masked anyway, quotients that leave the range, NaN / inf
    positions.  Bit for bit against the CPU oracle AND against the kernel generated without the fast body."""
    ctx = gpu_ctx
    n = 1300
    rng = np.random.default_rng(21)
    pos, mass, vel, acc = scenes.spawn_bodies(scenes.all_heavy_random_cube_with_bodies(n), 17)
    if what in ("denormal_gm", "mixed"):
        mass[rng.choice(n, 40, replace=False)] = np.float32(1e-30)   # G*m = 6.67e-41: denormal
        mass[rng.choice(n, 5, replace=False)] = np.float32(0.0)
    if what in ("huge_distances", "mixed"):
        pos[rng.choice(n, 30, replace=False)] *= np.float32(3e7)     # d2 ~ 1e21..1e25 > 2^63
        pos[rng.choice(n, 3, replace=False)] = np.float32(1e19)      # d2 overflows to inf
    if what in ("tiny_distances", "mixed"):
        k = rng.choice(n - 1, 30, replace=False)
        pos[k + 1] = pos[k] + rng.uniform(-4e-4, 4e-4, (30, 3)).astype(np.float32)  # around epsilon: some masked, some huge quotients
        mass[k] = np.float32(3e33)                                   # acceleration beyond 2^63
    if what == "nan_and_inf":
        pos[5, 1] = np.float32(np.nan)
        pos[700] = np.float32(np.inf)
    got = {}
    for name, src in (("fast", GRAVITY), ("exact", _strip_fast(GRAVITY))):
        out = np.zeros_like(acc)
        for comp, arr in ((POS, pos), (MASS, mass), (ACC, out)):
            _bind(ctx, 1, comp, arr)
        sid = ctx.create_pair_kernel_system("gravity_" + name, src, "Total", "pair", "finish", GRAVITY_PARAMS)
        ctx.set_archetypes(sid, [1])
        ctx.set_query_archetypes(sid, [1])
        ctx.run_system(sid)
        ctx.download(1, ACC)
        got[name] = out
    with np.errstate(all="ignore"):
        want = nbody_oracle.gravity(pos, mass)
    same = lambda a, b: np.array_equal(a.view(np.uint32), b.view(np.uint32)) or np.array_equal(  # NaN payloads may differ
        np.where(np.isnan(a), np.float32(0), a).view(np.uint32), np.where(np.isnan(b), np.float32(0), b).view(np.uint32)) and np.array_equal(np.isnan(a), np.isnan(b))
    assert same(got["fast"], got["exact"])
    assert same(got["fast"], want)


DIV_SQRT_PROBE = """
struct Pair { float a, b; };
struct Out { float q, s; int in_range; };
void probe(const Pair& p, Out& o) {
    o.q = recs_div_rn_seq(p.a, p.b);
    o.s = recs_sqrt_rn_seq(p.b);
    o.in_range = (recs_div_in_range(p.a) & recs_div_in_range(p.b) ? 1 : 0) | (recs_sqrt_in_range(p.b) ? 2 : 0);
}
"""


@pytest.mark.gpu
def test_division_and_square_root_sequences_are_ieee_in_range(gpu_ctx):
    """recs_div_rn_seq / recs_sqrt_rn_seq (the generated kernels' branch-free helpers) against numpy's correctly rounded
    float32 division and square root: bit for bit on every operand pair their range tests admit, over four million pairs
    with exponents drawn across (and beyond) the admitted ranges, mantissas random and special (powers of two, all-ones)."""
    ctx = gpu_ctx
    n = 1 << 22
    rng = np.random.default_rng(5)

    def floats(lo_exp, hi_exp):
        e = rng.integers(lo_exp + 127, hi_exp + 128, n).astype(np.uint32)
        m = rng.integers(0, 1 << 23, n).astype(np.uint32)
        special = rng.integers(0, 16, n)
        m = np.where(special == 0, 0, np.where(special == 1, (1 << 23) - 1, np.where(special == 2, 1, m))).astype(np.uint32)
        sign = (rng.integers(0, 2, n).astype(np.uint32) << 31)
        return ((e << 23) | m | sign).view(np.float32)

    pairs = np.zeros((n, 2), np.float32)
    pairs[:, 0] = floats(-70, 70)
    pairs[:, 1] = np.abs(floats(-110, 127))
    out = np.zeros(n, dtype=np.dtype([("q", np.float32), ("s", np.float32), ("in_range", np.int32)]))
    _bind(ctx, 3, 31, pairs)
    _bind(ctx, 3, 32, out)
    sid = ctx.create_kernel_system("probe", DIV_SQRT_PROBE, "probe", [(31, 8, "read", "Pair"), (32, 12, "write", "Out")])
    ctx.set_archetypes(sid, [3])
    ctx.run_system(sid)
    ctx.download(3, 32)
    with np.errstate(all="ignore"):
        q = pairs[:, 0] / pairs[:, 1]
        s = np.sqrt(pairs[:, 1])
    div_ok = (out["in_range"] & 1) != 0
    sqrt_ok = (out["in_range"] & 2) != 0
    assert div_ok.sum() > n // 4 and sqrt_ok.sum() > n // 2
    a, b = np.abs(pairs[:, 0]), pairs[:, 1]
    assert np.array_equal(div_ok, (a >= 2.0 ** -63) & (a <= 2.0 ** 63) & (b >= 2.0 ** -63) & (b <= 2.0 ** 63))
    assert np.array_equal(out["q"][div_ok].view(np.uint32), q[div_ok].view(np.uint32))
    assert np.array_equal(out["s"][sqrt_ok].view(np.uint32), s[sqrt_ok].view(np.uint32))


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
