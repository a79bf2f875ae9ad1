"""Generic GPU systems (recs_gpu_system_create_kernel): any per-entity function over Read / Write / Entity
parameters, iterated over the concatenation of its matched archetypes' columns -- the par-iter query
contract of crates/ecs/src/systems/mod.rs:457-566, 627-739 and iteration.rs:26-57, 151-245.

CPU part: source generation + NVRTC compilation for sm_100a (needs no GPU).  GPU part: the reference's
frag_iter / simple_iter benchmark systems and filter / Entity / uniform cases, bit-exact against numpy
restatements of the Rust bodies (mul-then-add, never fused)."""
import struct

import numpy as np
import pytest

from recs_b200.gpu import compile_kernel_system

VEC3 = "struct Position { float x, y, z; };\nstruct Velocity { float x, y, z; };\n"
# crates/ecs_bench_suite/src/recs/simple_iter.rs:47-49
SIMPLE_ITER = VEC3 + "void movement(Position& p, const Velocity& v) { p.x += v.x; p.y += v.y; p.z += v.z; }\n"
# crates/ecs_bench_suite/src/recs/frag_iter.rs:40-42
FRAG_ITER = "struct Data { float v; };\nvoid doubling(Data& d) { d.v *= 2.0f; }\n"
# crates/n-body/src/lib.rs:87-93 with the time step as a uniform
MOVEMENT_DT = VEC3 + (
    "struct Dt { float dt; };\n"
    "void movement(Position& p, const Velocity& v, const Dt& u) {\n"
    "    p.x += v.x * u.dt; p.y += v.y * u.dt; p.z += v.z * u.dt;\n}\n"
)
POS, VEL, DATA, KEY, TAG_A, TAG_B = 11, 12, 13, 14, 15, 16


# ---- CPU: generation + compilation ---------------------------------------------------------------------------------
@pytest.mark.parametrize("direct", [False, True])
def test_kernel_system_compiles_for_sm100a_without_a_gpu(direct):
    st, log, cubin, chunk = compile_kernel_system(
        SIMPLE_ITER, "movement", [(POS, 12, "write", "Position"), (VEL, 12, "read", "Velocity")], force_direct=direct
    )
    assert st == 0, log
    assert cubin > 1000
    assert chunk == (0 if direct else 1024)  # pipelined form: 1024-entity chunks, 3 stages x 24 KiB


def test_wide_rows_shrink_the_staged_chunk_or_go_direct():
    src = "struct M { float m[16]; };\nstruct Big { float b[1024]; };\nvoid f(M& m) { m.m[0] += 1.0f; }\nvoid g(Big& b) { b.b[0] += 1.0f; }\n"
    st, log, _, chunk = compile_kernel_system(src, "f", [(1, 64, "write", "M")])
    assert st == 0, log
    assert chunk == 256  # 3 stages x 256 x 64 B = 48 KiB fits the 72 KiB budget, 512 entities would not
    st, log, _, chunk = compile_kernel_system(src, "g", [(1, 4096, "write", "Big")])
    assert st == 0, log
    assert chunk == 0  # 3 stages x 32 x 4 KiB does not fit: direct global references


def test_compile_errors_come_back_with_the_nvrtc_log():
    st, log, cubin, _ = compile_kernel_system(
        SIMPLE_ITER.replace("p.z += v.z;", "p.w += v.z;"), "movement", [(POS, 12, "write", "Position"), (VEL, 12, "read", "Velocity")]
    )
    assert st == 12 and cubin == 0  # RECS_ERR_COMPILE
    assert "has no member" in log and "w" in log


def test_element_size_mismatch_is_a_compile_error():
    st, log, _, _ = compile_kernel_system(SIMPLE_ITER, "movement", [(POS, 16, "write", "Position"), (VEL, 12, "read", "Velocity")])
    assert st == 12
    assert "sizeof(Position) differs from the bound column's element size" in log


def test_read_parameters_are_const():
    # a system that writes through a Read<T> must not compile (Read derefs to &T, systems/mod.rs:106-118)
    src = VEC3 + "void bad(Position& p, Velocity& v) { v.x = p.x; }\n"
    st, log, _, _ = compile_kernel_system(src, "bad", [(POS, 12, "write", "Position"), (VEL, 12, "read", "Velocity")])
    assert st == 12


# ---- GPU -----------------------------------------------------------------------------------------------------------
def _bind(ctx, arch, comp, array):
    ctx.bind_column(arch, comp, array)
    ctx.upload(arch, comp)


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 255, 1023, 1024, 1025, 4096, 10000, 300001])
@pytest.mark.parametrize("direct", [False, True])
def test_simple_iter_as_a_generic_system(gpu_ctx, monkeypatch, n, direct):
    """crates/ecs_bench_suite/src/recs/simple_iter.rs:28-36, 47-52 (10 000 entities there)."""
    from recs_b200 import GpuContext

    if direct:
        monkeypatch.setenv("RECS_GPU_KERNEL_DIRECT", "1")
    rng = np.random.default_rng(n)
    pos = rng.standard_normal((n, 3)).astype(np.float32)
    vel = rng.standard_normal((n, 3)).astype(np.float32)
    want = pos.copy()
    with GpuContext(device=0) as ctx:
        _bind(ctx, 1, POS, pos)
        _bind(ctx, 1, VEL, vel)
        sid = ctx.create_kernel_system("movement", SIMPLE_ITER, "movement", [(POS, 12, "write", "Position"), (VEL, 12, "read", "Velocity")])
        assert ctx.describe_system(sid) == ("movement", [(POS, True), (VEL, False)])
        ctx.set_archetypes(sid, [1])
        for _ in range(3):
            ctx.run_system(sid)
            want = want + vel
        ctx.download(1, POS)
        ctx.download(1, VEL)
    assert np.array_equal(pos.view(np.uint32), want.view(np.uint32))


@pytest.mark.gpu
def test_frag_iter_26_archetypes(gpu_ctx):
    """crates/ecs_bench_suite/src/recs/frag_iter.rs:6-43: 26 archetypes x 20 entities share `Data`; the system
    doubles every value exactly once per run.  Ragged counts and an empty archetype are added here."""
    ctx = gpu_ctx
    counts = [20] * 26 + [0, 1, 300, 1024, 0, 2051, 257]
    cols = []
    for a, c in enumerate(counts):
        col = (np.arange(c, dtype=np.float32) + 1000.0 * a + 1.0).reshape(c, 1)
        cols.append(col)
        _bind(ctx, a + 1, DATA, col)
    want = [c.copy() for c in cols]
    sid = ctx.create_kernel_system("doubling", FRAG_ITER, "doubling", [(DATA, 4, "write", "Data")])
    ctx.set_archetypes(sid, list(range(1, len(counts) + 1)))
    ctx.reset_timing()
    for _ in range(2):
        ctx.run_system(sid)
    assert ctx.timing()["kernel_launches"] == 2  # one launch for all archetypes
    for a in range(len(counts)):
        ctx.download(a + 1, DATA)
        assert np.array_equal(cols[a], want[a] * np.float32(4.0)), a


@pytest.mark.gpu
def test_uniform_and_unfused_arithmetic(gpu_ctx):
    """`p += v * dt` must round like Rust (mul, then add): the generated kernels are compiled with --fmad=false."""
    ctx = gpu_ctx
    n = 5000
    rng = np.random.default_rng(3)
    pos = (rng.standard_normal((n, 3)) * 100).astype(np.float32)
    vel = (rng.standard_normal((n, 3)) * 1e4).astype(np.float32)
    dt = np.float32(1.0 / 20000.0)
    want = pos + vel * dt  # numpy: two roundings
    fused = (pos.astype(np.float64) + vel.astype(np.float64) * np.float64(dt)).astype(np.float32)
    assert not np.array_equal(want, fused)  # the data distinguishes the two
    _bind(ctx, 1, POS, pos)
    _bind(ctx, 1, VEL, vel)
    sid = ctx.create_kernel_system(
        "movement", MOVEMENT_DT, "movement", [(POS, 12, "write", "Position"), (VEL, 12, "read", "Velocity")], "Dt", 4
    )
    ctx.set_archetypes(sid, [1])
    ctx.set_system_uniform(sid, struct.pack("f", dt))
    ctx.run_system(sid)
    ctx.download(1, POS)
    assert np.array_equal(pos.view(np.uint32), want.view(np.uint32))


@pytest.mark.gpu
def test_generic_system_in_a_captured_tick_and_uniform_change(gpu_ctx):
    ctx = gpu_ctx
    n = 1000
    pos = np.zeros((n, 3), np.float32)
    vel = np.ones((n, 3), np.float32)
    _bind(ctx, 1, POS, pos)
    _bind(ctx, 1, VEL, vel)
    sid = ctx.create_kernel_system(
        "movement", MOVEMENT_DT, "movement", [(POS, 12, "write", "Position"), (VEL, 12, "read", "Velocity")], "Dt", 4
    )
    ctx.set_archetypes(sid, [1])
    ctx.set_tick_order([sid])
    ctx.set_system_uniform(sid, struct.pack("f", 0.5))
    ctx.tick(4)  # eager tick + captured graph
    ctx.set_system_uniform(sid, struct.pack("f", 0.25))
    ctx.tick(4)  # must not replay the old uniform
    ctx.download(1, POS)
    assert np.array_equal(pos, np.full((n, 3), 3.0, np.float32))


@pytest.mark.gpu
def test_entity_parameter(gpu_ctx):
    """`Entity` parameter (crates/ecs/src/systems/mod.rs:868-912): the body sees generation << 32 | id of its row."""
    from recs_b200 import _lib

    ctx = gpu_ctx
    src = "struct Key { unsigned long long k; };\nvoid stamp(const unsigned long long& e, Key& out) { out.k = e * 2ull + 1ull; }\n"
    keys = [np.arange(5, dtype=np.uint64) + (7 << 32), np.arange(400, dtype=np.uint64) + 100]
    outs = [np.zeros(5, np.uint64), np.zeros(400, np.uint64)]
    for a in range(2):
        _bind(ctx, a + 1, _lib.ENTITY_COMPONENT, keys[a])
        _bind(ctx, a + 1, KEY, outs[a])
    sid = ctx.create_kernel_system("stamp", src, "stamp", [(0, 8, "entity", "unsigned long long"), (KEY, 8, "write", "Key")])
    assert ctx.describe_system(sid) == ("stamp", [(KEY, True)])  # Entity is no component access (mod.rs:893-895)
    ctx.set_archetypes(sid, [1, 2])
    ctx.run_system(sid)
    for a in range(2):
        ctx.download(a + 1, KEY)
        assert np.array_equal(outs[a], keys[a] * 2 + 1)


@pytest.mark.gpu
def test_missing_column_and_size_mismatch_are_reported(gpu_ctx):
    from recs_b200 import RecsGpuError

    ctx = gpu_ctx
    pos = np.zeros((10, 3), np.float32)
    _bind(ctx, 1, POS, pos)
    sid = ctx.create_kernel_system("movement", SIMPLE_ITER, "movement", [(POS, 12, "write", "Position"), (VEL, 12, "read", "Velocity")])
    ctx.set_archetypes(sid, [1])
    with pytest.raises(RecsGpuError) as e:
        ctx.run_system(sid)
    assert e.value.code == 5  # RECS_ERR_MISSING_PARAMETER
    ctx.clear_error()
    with pytest.raises(RecsGpuError) as e:
        ctx.create_kernel_system("broken", "void f(int& x) { x = ; }", "f", [(POS, 4, "write", "int")])
    assert e.value.code == 12 and "expected an expression" in str(e.value)
    ctx.clear_error()


@pytest.mark.gpu
def test_application_runs_generic_systems_with_filters(gpu_ctx):
    """Through the Application: archetype membership with filters is computed by the host layer
    (crates/ecs/src/filter.rs, systems/mod.rs:986-1000), the schedule sees the component accesses, and a
    structural change re-binds.  Checked against the Python ECS oracle's membership."""
    from oracle import ecs_oracle
    from recs_b200.ecs import Application

    app = Application(gpu_ctx)
    w = app.world
    w.register_component(DATA, np.dtype(np.float32), "Data")
    w.register_component(TAG_A, np.dtype(np.float32), "A")
    w.register_component(TAG_B, np.dtype(np.float32), "B")
    ow = ecs_oracle.World()
    layouts = [(DATA,), (DATA, TAG_A), (DATA, TAG_B), (DATA, TAG_A, TAG_B), (TAG_A,)]
    ents = []
    for k, comps in enumerate(layouts):
        for i in range(3 + 50 * k):
            vals = {c: np.float32(100 * k + i + 1) for c in comps}
            ents.append(w.create_entity(vals))
            ow.create_entity({c: float(v) for c, v in vals.items()})
    params = [("write", DATA), ("with", TAG_A), ("not", ("with", TAG_B))]
    app.add_gpu_kernel_system("doubling", FRAG_ITER, "doubling", params, ["Data"])
    matched = sorted(w.query_archetypes(params))
    assert matched == sorted(ecs_oracle.get_archetype_indices(ow, params)) and len(matched) == 1
    app.run(2)
    app.sync_to_host()
    for arch in range(w.archetype_count()):
        comps = w.archetype_components(arch)
        if DATA not in comps:
            continue
        col = w.column(arch, DATA)
        n = len(w.archetype_entities(arch))
        k = [set(l) for l in layouts].index(set(comps))
        base = np.array([100 * k + i + 1 for i in range(n)], np.float32)
        factor = np.float32(4.0) if arch in matched else np.float32(1.0)
        assert np.array_equal(col[:n].reshape(-1), base * factor), (arch, comps)
    # structural change: a DATA+A entity gains B and leaves the match; the rest keeps doubling
    moved = ents[3 + 1]  # second entity of layout 1 (DATA, TAG_A)
    w.add_components(moved, {TAG_B: np.float32(0.0)})
    app.run(1)
    app.sync_to_host()
    arch = matched[0]
    col = w.column(arch, DATA)
    n = len(w.archetype_entities(arch))
    assert n == 53 - 1
    assert np.all(col[:n].reshape(-1) % np.float32(8.0) == 0)
    app.close()


@pytest.mark.gpu
@pytest.mark.parametrize("floats,expect_chunk", [(16, 256), (1024, 0)])
def test_wide_components_pipelined_and_direct(gpu_ctx, floats, expect_chunk):
    """A Matrix4-sized component (64 B: 256-entity chunks) and a 4 KiB one (direct form), two archetypes, ragged
    counts; body = an in-place transform that reads the whole element (heavy_compute's access pattern,
    crates/ecs_bench_suite/src/recs/heavy_compute.rs:47-52, without cgmath's invert)."""
    ctx = gpu_ctx
    src = (f"struct W {{ float m[{floats}]; }};\nstruct P {{ float x, y, z; }};\n"
           "void f(W& w, P& p) {\n"
           f"    float s = 0.0f; for (int i = 0; i < {floats}; ++i) s += w.m[i];\n"
           f"    for (int i = 0; i < {floats}; ++i) w.m[i] = w.m[i] * 2.0f + (float)i;\n"
           "    p.x = s; p.y += 1.0f;\n}\n")
    st, log, _, chunk = compile_kernel_system(src, "f", [(1, floats * 4, "write", "W"), (2, 12, "write", "P")])
    assert st == 0 and chunk == expect_chunk, log
    rng = np.random.default_rng(floats)
    data = {}
    for arch, count in ((1, 700), (2, 3)):
        w = rng.integers(-8, 8, size=(count, floats)).astype(np.float32)  # small integers: sums are exact in any order
        p = np.zeros((count, 3), np.float32)
        data[arch] = (w, p, w.copy())
        _bind(ctx, arch, 1, w)
        _bind(ctx, arch, 2, p)
    sid = ctx.create_kernel_system("f", src, "f", [(1, floats * 4, "write", "W"), (2, 12, "write", "P")])
    ctx.set_archetypes(sid, [1, 2])
    ctx.run_system(sid)
    for arch, (w, p, w0) in data.items():
        ctx.download(arch, 1)
        ctx.download(arch, 2)
        assert np.array_equal(w, w0 * np.float32(2.0) + np.arange(floats, dtype=np.float32))
        assert np.array_equal(p[:, 0], w0.sum(axis=1)) and np.all(p[:, 1] == 1.0) and np.all(p[:, 2] == 0.0)


GROUND = ("struct Position { float x, y, z; };\n"
          # crates/rain-simulation/src/lib.rs:152-156
          "void ground_collision(const unsigned long long& entity, const Position& p, const recs_commands& commands) {\n"
          "    if (p.y <= 0.0f) commands.remove();\n}\n")


@pytest.mark.gpu
@pytest.mark.parametrize("direct", [False, True])
def test_commands_remove_in_a_generic_system(monkeypatch, direct):
    """`ground_collision(Entity, Read<Position>, Commands)` written as a generic system: the rows it removes must be the
    rows the built-in RECS_SYS_RAIN_GROUND_COLLISION kind flags, over several archetypes incl. an empty one."""
    from recs_b200 import GpuContext, _lib

    if direct:
        monkeypatch.setenv("RECS_GPU_KERNEL_DIRECT", "1")
    rng = np.random.default_rng(9)
    with GpuContext(device=0) as ctx:
        cols = {}
        for arch, count in ((1, 3000), (2, 0), (3, 77), (4, 1025)):
            pos = (rng.standard_normal((count, 3)) * 2).astype(np.float32)
            pos[::5, 1] = 0.0  # exactly on the ground: removed (`<=`)
            cols[arch] = pos
            _bind(ctx, arch, POS, pos)
            _bind(ctx, arch, _lib.ENTITY_COMPONENT, np.arange(count, dtype=np.uint64))
        gen = ctx.create_kernel_system(
            "ground_collision", GROUND, "ground_collision",
            [(0, 8, "entity", "unsigned long long"), (POS, 12, "read", "Position"), (0, 0, "commands", None)])
        assert ctx.describe_system(gen) == ("ground_collision", [(POS, False)])
        builtin = ctx.create_system(_lib.SYS_RAIN_GROUND_COLLISION, [POS])
        for s in (gen, builtin):
            ctx.set_archetypes(s, [1, 2, 3, 4])
        for _ in range(2):  # flags are cleared per run
            ctx.run_system(gen)
        ctx.run_system(builtin)
        for arch, pos in cols.items():
            want = np.flatnonzero(pos[:, 1] <= 0.0).astype(np.uint32)
            assert np.array_equal(ctx.system_removals(gen, arch), want), arch
            assert np.array_equal(ctx.system_removals(builtin, arch), want), arch


@pytest.mark.gpu
def test_generic_removal_system_feeds_command_playback(gpu_ctx):
    """Through the Application: a generic system with a Commands parameter removes entities at the end of the tick
    (playback on the host World, resident columns follow by row moves) while another generic system keeps writing
    the same archetype."""
    from recs_b200.ecs import Application

    app = Application(gpu_ctx)
    w = app.world
    F3 = np.dtype((np.float32, 3))
    w.register_component(POS, F3, "Position")
    n = 500
    pos = np.zeros((n, 3), np.float32)
    pos[:, 1] = np.arange(n, dtype=np.float32) * 0.01 + 0.005  # heights 0.005 .. 4.995
    w.create_entities(n, {POS: pos})
    fall = "struct Position { float x, y, z; };\nvoid fall(Position& p) { p.y -= 1.0f; }\n"
    app.add_gpu_kernel_system("fall", fall, "fall", [("write", POS)], ["Position"])
    app.add_gpu_kernel_system("ground_collision", GROUND, "ground_collision",
                              [("entity",), ("read", POS), ("commands",)], ["unsigned long long", "Position"])
    alive = n
    for tick in range(1, 5):
        app.run(1)
        order, _ = app.last_tick_order()
        names = [app.system_names[i] for i in order]
        # the reader of Position precedes its writer (precedence.rs:39-66): collisions are decided on last tick's heights
        assert names == ["ground_collision", "fall"]
        heights_before = pos[:, 1] - np.float32(tick - 1)
        expected_alive = int(np.count_nonzero(heights_before > 0))
        created, removed = app.last_playback()
        assert created == 0 and alive - removed == expected_alive, tick
        alive = expected_alive
    app.sync_to_host()
    left = w.column(1, POS)[:alive]
    assert alive == len(w.archetype_entities(1)) and alive > 0
    assert np.all(left[:, 1] + np.float32(1.0) > 0)  # every survivor was above ground when last checked
    h2d, d2h = gpu_ctx.transfer_bytes()
    assert d2h == n * 12 * 0 + alive * 12  # only the final sync came back
    app.close()


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(10))
def test_random_layouts_against_numpy(gpu_ctx, seed):
    """Fuzz of the generated kernels: random element sizes (1 ... 40 bytes, odd ones included), 1-4 parameters, 1-5
    archetypes with random row counts around the chunk boundaries (0, 1, 1023, 1024, 1025, ...).  Body: every Write
    parameter gets its first and last byte bumped by a value that depends on the first byte of every Read parameter."""
    rng = np.random.default_rng(1000 + seed)
    n_params = int(rng.integers(1, 5))
    sizes = [int(rng.choice([1, 2, 3, 4, 6, 8, 12, 16, 20, 36, 40])) for _ in range(n_params)]
    kinds = ["write"] + [str(rng.choice(["read", "write"])) for _ in range(n_params - 1)]
    counts = [int(rng.choice([0, 1, 31, 255, 1023, 1024, 1025, 2048, 3000])) for _ in range(int(rng.integers(1, 6)))]
    src = "".join(f"struct T{i} {{ unsigned char b[{s}]; }};\n" for i, s in enumerate(sizes))
    args = ", ".join(f"{'const ' if k == 'read' else ''}T{i}& p{i}" for i, k in enumerate(kinds))
    reads = " + ".join([f"p{i}.b[0]" for i, k in enumerate(kinds) if k == "read"] or ["0"])
    body = "".join(f"    p{i}.b[0] += (unsigned char)(1 + k); p{i}.b[{s - 1}] += (unsigned char)(2 + k);\n"
                   for i, (s, k) in enumerate(zip(sizes, kinds)) if k == "write")
    src += f"void f({args}) {{\n    const unsigned k = {reads};\n{body}}}\n"
    ctx = gpu_ctx
    data = {}
    for a, c in enumerate(counts, start=1):
        cols = []
        for i, s in enumerate(sizes):
            col = rng.integers(0, 256, size=(c, s), dtype=np.uint8)
            cols.append(col)
            ctx.bind_column(a, 50 + i, col, elem_size=s)
            ctx.upload(a, 50 + i)
        data[a] = (cols, [c_.copy() for c_ in cols])
    params = [(50 + i, s, k, f"T{i}") for i, (s, k) in enumerate(zip(sizes, kinds))]
    sid = ctx.create_kernel_system("fuzz", src, "f", params)
    ctx.set_archetypes(sid, list(data))
    ctx.run_system(sid)
    for a, (cols, before) in data.items():
        k = np.zeros(len(before[0]), np.uint32)
        for i, kind in enumerate(kinds):
            if kind == "read":
                k += before[i][:, 0]
        for i, (s, kind) in enumerate(zip(sizes, kinds)):
            ctx.download(a, 50 + i)
            want = before[i].copy()
            if kind == "write":
                if s == 1:
                    want[:, 0] = (want[:, 0].astype(np.uint32) + 1 + k + 2 + k).astype(np.uint8)
                else:
                    want[:, 0] = (want[:, 0].astype(np.uint32) + 1 + k).astype(np.uint8)
                    want[:, s - 1] = (want[:, s - 1].astype(np.uint32) + 2 + k).astype(np.uint8)
            assert np.array_equal(cols[i], want), (seed, a, i, sizes, kinds, counts)


# ---- Commands::create in generic systems (crates/ecs/src/systems/command_buffers.rs:53-66) ---------------------------
SPAWNER = """
struct Budget { int n; };
struct Tag { float v; };
struct Child { float parent; float index; float tag; };
void spawner(Budget& budget, const Tag& tag, const recs_commands& commands) {
    for (int k = 0; k < budget.n; ++k) commands.create(Child{tag.v, (float)k, tag.v * 10.0f + (float)k});
    budget.n = 0;
}
"""
SPAWNER_PARAMS = [(1, 4, "write", "Budget"), (2, 4, "read", "Tag"), (0, 12, "commands", "Child")]


def test_creating_system_compiles_without_a_gpu():
    st, log, cubin, _ = compile_kernel_system(SPAWNER, "spawner", SPAWNER_PARAMS)
    assert st == 0 and cubin > 1000, log
    # the registered record size is checked against the struct in the source
    st, log, _, _ = compile_kernel_system(SPAWNER, "spawner", SPAWNER_PARAMS[:2] + [(0, 16, "commands", "Child")])
    assert st == 12 and "Commands::create record" in log


@pytest.mark.gpu
@pytest.mark.parametrize("counts", [(3000, 0), (700, 1300), (1, 1)])
def test_create_records_come_back_in_sequential_run_order(gpu_ctx, counts):
    """Every entity of two archetypes spawns 0..3 children; thousands of threads race for slots of the append buffer,
    yet the records come back in the order a sequential run over (archetype, row) would have recorded them -- the order
    that fixes the entity ids handed out at playback."""
    ctx = gpu_ctx
    rng = np.random.default_rng(5)
    cols, want = {}, []
    for arch, n in zip((1, 2), counts):
        budget = rng.integers(0, 4, size=n).astype(np.int32)
        tag = (np.arange(n, dtype=np.float32) + np.float32(arch * 100000))
        cols[arch] = (budget, tag)
        ctx.bind_column(arch, 1, budget)
        ctx.bind_column(arch, 2, tag)
        ctx.upload(arch, 1)
        ctx.upload(arch, 2)
        for i in range(n):
            want += [(tag[i], float(k), tag[i] * np.float32(10.0) + np.float32(k)) for k in range(budget[i])]
    sid = ctx.create_kernel_system("spawner", SPAWNER, "spawner", SPAWNER_PARAMS)
    ctx.set_archetypes(sid, [1, 2])
    ctx.run_system(sid)
    child = np.dtype([("parent", np.float32), ("index", np.float32), ("tag", np.float32)])
    got = ctx.system_creates(sid, child)
    assert len(got) == len(want)
    assert np.array_equal(got.view(np.float32).reshape(-1, 3), np.array(want, np.float32).reshape(-1, 3))
    ctx.download(1, 1)
    assert not cols[1][0].any()  # budgets spent
    # a second run records nothing
    ctx.run_system(sid)
    assert len(ctx.system_creates(sid, child)) == 0


@pytest.mark.gpu
def test_create_buffer_overflow_is_an_error_not_a_silent_drop(gpu_ctx):
    from recs_b200 import RecsGpuError

    ctx = gpu_ctx
    n = 1000
    budget = np.full(n, 3, np.int32)
    tag = np.arange(n, dtype=np.float32)
    ctx.bind_column(1, 1, budget)
    ctx.bind_column(1, 2, tag)
    ctx.upload(1, 1)
    ctx.upload(1, 2)
    sid = ctx.create_kernel_system("spawner", SPAWNER, "spawner", SPAWNER_PARAMS)
    ctx.set_archetypes(sid, [1])
    ctx.set_create_capacity(sid, 100)
    ctx.run_system(sid)
    child = np.dtype([("parent", np.float32), ("index", np.float32), ("tag", np.float32)])
    with pytest.raises(RecsGpuError) as e:
        ctx.system_creates(sid, child)
    assert e.value.code == 7 and "3000 records" in str(e.value)
    ctx.clear_error()
    ctx.set_create_capacity(sid, 4096)
    ctx.upload(1, 1)
    ctx.run_system(sid)
    assert len(ctx.system_creates(sid, child)) == 3000


# ---- heavy_compute (crates/ecs_bench_suite/src/recs/heavy_compute.rs:47-52) as a generic system ----------------------
# fn heavy_computation_system(mut position: Write<Position>, mut affine: Write<Affine>) {
#     for _ in 0..100 { affine.0 = affine.0.invert().unwrap(); }
#     position.0 = affine.0.transform_vector(position.0);
# }
# Matrix4::invert / transform_vector are cgmath 0.18.0 (Cargo.lock:1025-1026, not vendored): restated from its published
# source (matrix.rs: `det_sub_proc_unsafe`, `SquareMatrix::invert`, `Mul<Vector4> for Matrix4` as the sum of the columns
# scaled by the vector's components) -- the unverifiable third-party assumption of SURVEY 8(c).  The device body and the
# numpy restatement below are written independently from that description, one operation per line.
HEAVY_COMPUTE = """
struct Position { float x, y, z; };
struct Affine { float m[16]; };                       // Matrix4<f32>, column major: m[4 * column + row]
struct V4 { float x, y, z, w; };
static V4 det_sub_proc(const float* s, int x, int y, int z) {
    const V4 a{s[4 + x], s[12 + x], s[x], s[8 + x]};
    const V4 b{s[8 + y], s[8 + y], s[4 + y], s[4 + y]};
    const V4 c{s[12 + z], s[z], s[12 + z], s[z]};
    const V4 d{s[8 + x], s[8 + x], s[4 + x], s[4 + x]};
    const V4 e{s[12 + y], s[y], s[12 + y], s[y]};
    const V4 f{s[4 + z], s[12 + z], s[z], s[8 + z]};
    const V4 g{s[12 + x], s[x], s[12 + x], s[x]};
    const V4 h{s[4 + y], s[12 + y], s[y], s[8 + y]};
    const V4 i{s[8 + z], s[8 + z], s[4 + z], s[4 + z]};
    V4 t{a.x * (b.x * c.x), a.y * (b.y * c.y), a.z * (b.z * c.z), a.w * (b.w * c.w)};
    t.x += d.x * (e.x * f.x); t.y += d.y * (e.y * f.y); t.z += d.z * (e.z * f.z); t.w += d.w * (e.w * f.w);
    t.x += g.x * (h.x * i.x); t.y += g.y * (h.y * i.y); t.z += g.z * (h.z * i.z); t.w += g.w * (h.w * i.w);
    t.x -= a.x * (e.x * i.x); t.y -= a.y * (e.y * i.y); t.z -= a.z * (e.z * i.z); t.w -= a.w * (e.w * i.w);
    t.x -= d.x * (h.x * c.x); t.y -= d.y * (h.y * c.y); t.z -= d.z * (h.z * c.z); t.w -= d.w * (h.w * c.w);
    t.x -= g.x * (b.x * f.x); t.y -= g.y * (b.y * f.y); t.z -= g.z * (b.z * f.z); t.w -= g.w * (b.w * f.w);
    return t;
}
static void invert(float* s) {
    const V4 t0 = det_sub_proc(s, 1, 2, 3);
    const float det = ((t0.x * s[0] + t0.y * s[4]) + t0.z * s[8]) + t0.w * s[12];    // tmp0.dot(row 0)
    const float inv_det = 1.0f / det;                                                 // (.unwrap(): det != 0)
    const V4 t1 = det_sub_proc(s, 0, 3, 2), t2 = det_sub_proc(s, 0, 1, 3), t3 = det_sub_proc(s, 0, 2, 1);
    const V4 cols[4] = {t0, t1, t2, t3};
    for (int c = 0; c < 4; ++c) {
        s[4 * c + 0] = cols[c].x * inv_det; s[4 * c + 1] = cols[c].y * inv_det;
        s[4 * c + 2] = cols[c].z * inv_det; s[4 * c + 3] = cols[c].w * inv_det;
    }
}
void heavy_computation_system(Position& position, Affine& affine) {
    for (int k = 0; k < 100; ++k) invert(affine.m);
    const float* m = affine.m;   // transform_vector: (m * v.extend(0)).truncate(), columns scaled and summed
    const float vx = position.x, vy = position.y, vz = position.z;
    position.x = ((m[0] * vx + m[4] * vy) + m[8] * vz) + m[12] * 0.0f;
    position.y = ((m[1] * vx + m[5] * vy) + m[9] * vz) + m[13] * 0.0f;
    position.z = ((m[2] * vx + m[6] * vy) + m[10] * vz) + m[14] * 0.0f;
}
"""


def _heavy_compute_numpy(pos, aff, inversions=100):
    """The same operations on whole columns in numpy float32 (every * and + rounds on its own, like rustc's code)."""
    s = aff.copy()  # (n, 16)

    def dsp(x, y, z):
        col = lambda k: s[:, k]
        a = np.stack([col(4 + x), col(12 + x), col(x), col(8 + x)], 1)
        b = np.stack([col(8 + y), col(8 + y), col(4 + y), col(4 + y)], 1)
        c = np.stack([col(12 + z), col(z), col(12 + z), col(z)], 1)
        d = np.stack([col(8 + x), col(8 + x), col(4 + x), col(4 + x)], 1)
        e = np.stack([col(12 + y), col(y), col(12 + y), col(y)], 1)
        f = np.stack([col(4 + z), col(12 + z), col(z), col(8 + z)], 1)
        g = np.stack([col(12 + x), col(x), col(12 + x), col(x)], 1)
        h = np.stack([col(4 + y), col(12 + y), col(y), col(8 + y)], 1)
        i = np.stack([col(8 + z), col(8 + z), col(4 + z), col(4 + z)], 1)
        t = a * (b * c)
        t = t + d * (e * f)
        t = t + g * (h * i)
        t = t - a * (e * i)
        t = t - d * (h * c)
        t = t - g * (b * f)
        return t

    for _ in range(inversions):
        t0 = dsp(1, 2, 3)
        det = ((t0[:, 0] * s[:, 0] + t0[:, 1] * s[:, 4]) + t0[:, 2] * s[:, 8]) + t0[:, 3] * s[:, 12]
        inv_det = (np.float32(1.0) / det)[:, None]
        t1, t2, t3 = dsp(0, 3, 2), dsp(0, 1, 3), dsp(0, 2, 1)
        s = np.concatenate([t0 * inv_det, t1 * inv_det, t2 * inv_det, t3 * inv_det], 1).astype(np.float32)
    v = pos
    zero = np.float32(0.0)
    out = np.stack([((s[:, r] * v[:, 0] + s[:, 4 + r] * v[:, 1]) + s[:, 8 + r] * v[:, 2]) + s[:, 12 + r] * zero for r in range(3)], 1)
    return out.astype(np.float32), s


def test_heavy_compute_compiles_without_a_gpu():
    st, log, cubin, chunk = compile_kernel_system(HEAVY_COMPUTE, "heavy_computation_system",
                                                  [(1, 12, "write", "Position"), (2, 64, "write", "Affine")])
    assert st == 0 and cubin > 1000 and chunk > 0, log


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1000, 4097])
def test_heavy_compute_as_generic_system_bit_exact(gpu_ctx, n):
    """The benchmark's 1 000 entities (Affine = from_angle_x(1.2 rad), Position = unit_x), and a larger set of random
    well-conditioned matrices: 100 inversions + transform_vector on the device equal the numpy restatement bit for bit,
    and the matrix really is an inverse pair (A^-1 A = I), which pins the index pattern of det_sub_proc."""
    ctx = gpu_ctx
    rng = np.random.default_rng(n)
    aff = np.zeros((n, 16), np.float32)
    c, sn = np.float32(np.cos(np.float32(1.2))), np.float32(np.sin(np.float32(1.2)))
    aff[:] = np.array([1, 0, 0, 0, 0, c, sn, 0, 0, -sn, c, 0, 0, 0, 0, 1], np.float32)  # Matrix4::from_angle_x, column major
    pos = np.zeros((n, 3), np.float32)
    pos[:, 0] = 1.0
    if n != 1000:
        aff += (rng.standard_normal((n, 16)) * 0.2).astype(np.float32)
        pos = rng.standard_normal((n, 3)).astype(np.float32)
    start = aff.copy()
    want_pos, want_aff = _heavy_compute_numpy(pos, aff)
    ctx.bind_column(1, 1, pos)
    ctx.bind_column(1, 2, aff)
    ctx.upload(1, 1)
    ctx.upload(1, 2)
    sid = ctx.create_kernel_system("heavy_compute", HEAVY_COMPUTE, "heavy_computation_system",
                                   [(1, 12, "write", "Position"), (2, 64, "write", "Affine")])
    ctx.set_archetypes(sid, [1])
    ctx.run_system(sid)
    ctx.download(1, 1)
    ctx.download(1, 2)
    assert np.array_equal(aff.view(np.uint32), want_aff.view(np.uint32))
    assert np.array_equal(pos.view(np.uint32), want_pos.view(np.uint32))
    # 100 inversions = 50 round trips: back at the start up to rounding; and one inversion really inverts
    assert np.allclose(aff, start, rtol=0, atol=5e-3)
    one_step = _heavy_compute_numpy(pos[:64], start[:64], inversions=1)[1]
    as_matrix = lambda flat: flat.astype(np.float64).reshape(-1, 4, 4).transpose(0, 2, 1)  # column major -> matrices
    assert np.allclose(as_matrix(one_step) @ as_matrix(start[:64]), np.eye(4), atol=1e-4)
