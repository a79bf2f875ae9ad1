"""Application run loop (crates/ecs/src/lib.rs:314-341) with CPU systems (no GPU needed) and, on
the GPU box, the n-body benchmark assembled exactly like
crates/ecs_bench_suite/src/recs/n_body.rs:54-63 behind the World / System / Schedule contract."""
import ctypes as C

import numpy as np
import pytest

from recs_b200 import _lib, ecs, scenes
from tests.util import rel_err_rows

POS, MASS, VEL, ACC = 1, 2, 3, 4
F3 = np.dtype((np.float32, 3))


def _f32(ptr, count, width=1):
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_float)), shape=(count * width,))


def test_cpu_application_runs_systems_in_schedule_order():
    app = ecs.Application(gpu=None)
    w = app.world
    w.register_component(POS, F3, "Position")
    w.register_component(VEL, F3, "Velocity")
    n = 10
    w.create_entities(n, {POS: np.zeros((n, 3), np.float32), VEL: np.ones((n, 3), np.float32)})
    log = []

    def movement(arch, count, cols):  # (Write<Position>, Read<Velocity>)
        log.append("movement")
        p, v = _f32(cols[0], count, 3), _f32(cols[1], count, 3)
        p += v * np.float32(0.5)

    def double_velocity(arch, count, cols):  # (Write<Velocity>,)
        log.append("double_velocity")
        _f32(cols[0], count, 3)[:] *= np.float32(2.0)

    def tick_counter(arch, count, cols):  # no parameters: runs once per tick
        log.append("tick_counter")

    app.add_cpu_system("movement", [("write", POS), ("read", VEL)], movement)
    app.add_cpu_system("double_velocity", [("write", VEL)], double_velocity)
    app.add_cpu_system("tick_counter", [], tick_counter)
    app.run(2)
    # readers of Velocity run before its writer within a tick (schedule/mod.rs:1208-1224)
    assert [x for x in log if x != "tick_counter"] == ["movement", "double_velocity"] * 2
    assert log.count("tick_counter") == 2
    assert np.array_equal(w.column(1, POS), np.full((n, 3), 0.5 + 1.0, np.float32))
    assert np.array_equal(w.column(1, VEL), np.full((n, 3), 4.0, np.float32))
    order, batches = app.last_tick_order()
    assert [app.system_names[i] for i in order if app.system_names[i] != "tick_counter"] == ["movement", "double_velocity"]


def test_gpu_system_without_gpu_context_is_refused():
    app = ecs.Application(gpu=None)
    with pytest.raises(ecs.EcsError) as e:
        app.add_gpu_system(_lib.SYS_NBODY_MOVEMENT, [POS, VEL])
    assert e.value.code == 2  # RECS_ERR_NO_DEVICE: no CPU fallback for GPU systems


def _nbody_app(gpu, n, seed, ticks_counter):
    app = ecs.Application(gpu=gpu)
    w = app.world
    w.register_component(POS, F3, "Position")
    w.register_component(MASS, np.float32, "Mass")
    w.register_component(VEL, F3, "Velocity")
    w.register_component(ACC, F3, "Acceleration")
    # registration order of the benchmark (n_body.rs:54-59)
    app.add_gpu_system(_lib.SYS_NBODY_MOVEMENT, [POS, VEL], "movement")
    app.add_gpu_system(_lib.SYS_NBODY_ACCELERATION, [VEL, ACC], "acceleration")
    app.add_gpu_system(_lib.SYS_NBODY_GRAVITY, [POS, ACC, MASS], "gravity")
    app.add_cpu_system("benchmark_system", [], lambda a, c, cols: ticks_counter.append(1))
    pos, mass, vel, acc = scenes.spawn_bodies(scenes.all_heavy_random_cube_with_bodies(n), seed)
    # create_planet_entity: app.create_entity((position, mass, velocity, acceleration)) (scenes.rs:108-117)
    w.create_entities(n, {POS: pos, MASS: mass, VEL: vel, ACC: acc})
    return app, (pos, mass, vel, acc)


@pytest.mark.gpu
def test_nbody_benchmark_through_application(gpu_ctx, nbody_oracle):
    ticks = []
    n = 1000
    app, (pos, mass, vel, acc) = _nbody_app(gpu_ctx, n, 1, ticks)
    w = app.world
    assert w.archetype_count() == 2 and w.archetype_entities(1) == [(k, 0) for k in range(n)]
    app.run(1)
    order, batches = app.last_tick_order()
    names = [app.system_names[i] for i in order]
    assert [x for x in names if x != "benchmark_system"] == ["gravity", "movement", "acceleration"]
    app.sync_to_host()
    nbody_oracle.run_sequential(pos, mass, vel, acc, 1)
    assert np.array_equal(w.column(1, POS).view(np.uint32), pos.view(np.uint32))
    assert rel_err_rows(w.column(1, ACC), acc).max() <= 1e-4
    assert rel_err_rows(w.column(1, VEL), vel).max() <= 1e-4
    app.run(9)
    assert len(ticks) == 10
    app.sync_to_host()
    nbody_oracle.run_sequential(pos, mass, vel, acc, 9)
    assert rel_err_rows(w.column(1, POS), pos, floor=1.0).max() <= 1e-4


@pytest.mark.gpu
def test_structural_change_rebinds_device_columns(gpu_ctx, nbody_oracle):
    """Entities created between ticks (command playback, lib.rs:321-323) move the host columns; the
    device mirror must re-bind and see the new rows."""
    ticks = []
    app, (pos, mass, vel, acc) = _nbody_app(gpu_ctx, 300, 2, ticks)
    w = app.world
    app.run(1)
    app.sync_to_host()
    extra = scenes.spawn_bodies(scenes.all_heavy_random_cube_with_bodies(50), 9)
    w.create_entities(50, {POS: extra[0], MASS: extra[1], VEL: extra[2], ACC: extra[3]})
    assert w.archetype_entities(1)[-1] == (349, 0)
    ref = [np.ascontiguousarray(w.column(1, c).copy()) for c in (POS, MASS, VEL, ACC)]
    app.run(1)
    app.sync_to_host()
    nbody_oracle.run_sequential(ref[0], ref[1], ref[2], ref[3], 1)
    assert np.array_equal(w.column(1, POS).view(np.uint32), ref[0].view(np.uint32))
    assert rel_err_rows(w.column(1, ACC), ref[3]).max() <= 1e-4


@pytest.mark.gpu
def test_cpu_system_sees_device_results(gpu_ctx):
    """A CPU system reading a column a GPU system wrote gets the downloaded values (the executor
    orders the copy like the reference orders the channel drop)."""
    seen = []
    app = ecs.Application(gpu=gpu_ctx)
    w = app.world
    w.register_component(POS, F3, "Position")
    w.register_component(VEL, F3, "Velocity")
    n = 64
    w.create_entities(n, {POS: np.zeros((n, 3), np.float32), VEL: np.full((n, 3), 2.0, np.float32)})
    app.add_gpu_system(_lib.SYS_QUERY_ADD, [POS, VEL], "add")  # pos += vel
    app.add_cpu_system("observer", [("read", POS)], lambda a, c, cols: seen.append(_f32(cols[0], c, 3).copy()))
    app.run(3)
    # observer reads Position, add writes it: readers first within a tick -> sees 0, 2, 4
    assert [float(s[0]) for s in seen] == [0.0, 2.0, 4.0]


# ---- WorkerPool executor (crates/scheduler/src/executor/mod.rs:101-291) -------------------------------------------------
def _segment_app(threads):
    app = ecs.Application(gpu=None)
    w = app.world
    DATA, TAG = 7, 8
    w.register_component(DATA, np.dtype(np.float32), "Data")
    w.register_component(TAG, np.dtype(np.float32), "Tag")
    counts = {1: 1000, 2: 37}  # archetype 1: {Data}, archetype 2: {Data, Tag}
    w.create_entities(counts[1], {DATA: np.arange(counts[1], dtype=np.float32)})
    w.create_entities(counts[2], {DATA: np.arange(counts[2], dtype=np.float32) + 5000, TAG: np.zeros(counts[2], np.float32)})
    app.set_executor(threads)
    return app, w, DATA, TAG, counts


@pytest.mark.parametrize("threads", [0, 1, 4])
def test_segment_system_visits_every_entity_exactly_once(threads):
    """The par-iter contract (crates/ecs/src/systems/iteration.rs:271-324: segmented == sequential visitation): under
    the WorkerPool executor the matched archetypes are cut into SegmentSize::Auto segments (systems/mod.rs:394-411),
    segments may span archetypes, and the union of all calls covers every row once."""
    import threading

    app, w, DATA, TAG, counts = _segment_app(threads)
    calls, lock, thread_ids = [], threading.Lock(), set()

    def doubling(arch, row_begin, row_count, cols):
        d = _f32(cols[0], counts[arch])
        d[row_begin : row_begin + row_count] *= np.float32(2.0)
        with lock:
            calls.append((arch, row_begin, row_count))
            thread_ids.add(threading.get_ident())

    app.add_cpu_segment_system("doubling", [("write", DATA)], doubling)
    app.run(1)
    assert np.array_equal(w.column(1, DATA).reshape(-1), np.arange(1000, dtype=np.float32) * 2)
    assert np.array_equal(w.column(2, DATA).reshape(-1), (np.arange(37, dtype=np.float32) + 5000) * 2)
    covered = {1: np.zeros(1000, int), 2: np.zeros(37, int)}
    for arch, b, c in calls:
        covered[arch][b : b + c] += 1
    assert all((v == 1).all() for v in covered.values())
    total = 1037
    if threads == 0:
        assert sorted(calls) == [(1, 0, 1000), (2, 0, 37)]  # Sequential executor: whole columns
    else:
        size = ecs.auto_segment_size(total, threads)  # max(16, N / (threads * 4))
        want = ecs.split_segments([1000, 37], size)
        want_calls = sorted((1 + col, begin, ln) for seg in want for (col, begin, ln) in seg if ln)
        assert sorted(calls) == want_calls
        assert len(want) == ecs.segment_count(total, size)
    if threads == 4:
        assert threading.get_ident() not in thread_ids  # bodies ran on worker threads, not on the caller


def test_pool_runs_non_conflicting_systems_of_a_batch_and_orders_conflicting_ones():
    app, w, DATA, TAG, counts = _segment_app(3)
    log = []

    def read_data(arch, row_begin, row_count, cols):
        log.append("read_data")

    def write_tag(arch, row_begin, row_count, cols):  # does not conflict with read_data: same batch
        _f32(cols[0], counts[arch])[row_begin : row_begin + row_count] += np.float32(1.0)
        log.append("write_tag")

    def write_data(arch, count, cols):  # plain system: one task, after its reader
        log.append("write_data")
        _f32(cols[0], count)[:] += np.float32(0.5)

    app.add_cpu_segment_system("read_data", [("read", DATA)], read_data)
    app.add_cpu_segment_system("write_tag", [("write", TAG)], write_tag)
    app.add_cpu_system("write_data", [("write", DATA)], write_data)
    app.run(3)
    order, batches = app.last_tick_order()
    names = [app.system_names[i] for i in order]
    assert batches[names.index("read_data")] < batches[names.index("write_data")]
    last_reader = max(i for i, x in enumerate(log) if x == "read_data")
    per_tick = len(log) // 3
    for t in range(3):  # within every tick all read_data segments precede write_data
        chunk = log[t * per_tick : (t + 1) * per_tick]
        assert max(i for i, x in enumerate(chunk) if x == "read_data") < min(i for i, x in enumerate(chunk) if x == "write_data")
    assert last_reader >= 0
    assert np.array_equal(w.column(2, TAG).reshape(-1), np.full(37, 3.0, np.float32))
    assert np.array_equal(w.column(1, DATA).reshape(-1), np.arange(1000, dtype=np.float32) + 1.5)


def test_commands_recorded_from_worker_threads_are_played_back():
    app, w, DATA, TAG, counts = _segment_app(4)

    def spawn(arch, row_begin, row_count, cols):  # every segment records one create
        app.command_create({DATA: np.float32(-1.0)})

    app.add_cpu_segment_system("spawn", [("read", DATA), ("commands",)], spawn)
    app.run(1)
    size = ecs.auto_segment_size(1037, 4)
    n_calls = sum(1 for seg in ecs.split_segments([1000, 37], size) for (_, _, ln) in seg if ln)
    assert app.last_playback() == (n_calls, 0)
    assert len(w.archetype_entities(1)) == 1000 + n_calls


def test_segment_systems_need_parallelizable_parameters():
    """`try_as_segment_iterable` is None unless every parameter supports parallelization; a nested Query does only
    if all of its accesses are reads (crates/ecs/src/systems/mod.rs:210-212, 1101-1105)."""
    app = ecs.Application(gpu=None)
    app.world.register_component(7, np.dtype(np.float32), "Data")
    app.add_cpu_segment_system("ok", [("read", 7), ("query", [("read", 7)])], lambda *a: None)
    with pytest.raises(ecs.EcsError) as e:
        app.add_cpu_segment_system("bad", [("read", 7), ("query", [("write", 7)])], lambda *a: None)
    assert e.value.code == 11  # RECS_ERR_UNSUPPORTED


@pytest.mark.gpu
def test_application_hands_gpu_chains_to_the_context_as_one(gpu_ctx, nbody_oracle):
    """Consecutive GPU systems of a tick reach the context as one chain, so its fusions apply behind the Application
    too: the n-body tick is one launch for a small world (two for a large one: interaction kernel, reduce with the
    movement / acceleration tail) even with a parameterless CPU system (`benchmark_system`, crates/ecs_bench_suite/src/recs/n_body.rs:59) in the schedule --
    and a CPU system that touches a column splits the chain without changing the results."""
    n, ticks = 2000, 6
    for with_cpu_reader in (False, True):
        gpu_ctx.reset_timing()
        app = ecs.Application(gpu=gpu_ctx)
        w = app.world
        for c, dt, name in ((POS, F3, "Position"), (MASS, np.float32, "Mass"), (VEL, F3, "Velocity"), (ACC, F3, "Acceleration")):
            w.register_component(c, dt, name)
        pos, mass, vel, acc = scenes.spawn_bodies(scenes.all_heavy_random_cube_with_bodies(n), 3)
        ref = [a.copy() for a in (pos, mass, vel, acc)]
        app.add_gpu_system(_lib.SYS_NBODY_MOVEMENT, [POS, VEL], "movement")
        app.add_gpu_system(_lib.SYS_NBODY_ACCELERATION, [VEL, ACC], "acceleration")
        app.add_gpu_system(_lib.SYS_NBODY_GRAVITY, [POS, ACC, MASS], "gravity")
        seen = []
        app.add_cpu_system("benchmark_system", [], lambda a, c, cols: seen.append(1))
        if with_cpu_reader:  # reads Acceleration on the host every tick: forces a flush + download mid-tick
            app.add_cpu_system("probe", [("read", ACC)], lambda a, c, cols: seen.append(float(_f32(cols[0], c, 3)[0])))
        w.create_entities(n, {POS: pos, MASS: mass, VEL: vel, ACC: acc})
        app.run(1)
        before = gpu_ctx.timing()["kernel_launches"]
        app.run(ticks - 1)
        per_tick = (gpu_ctx.timing()["kernel_launches"] - before) / (ticks - 1)
        if not with_cpu_reader:
            assert per_tick == 1  # (2 000 bodies: the whole fused tick is one cooperative launch; larger worlds take two)
        app.sync_to_host()
        # the extra reader of Acceleration changes the DAG: follow whatever order the schedule produced
        from oracle.nbody import SYS_ACCELERATION, SYS_GRAVITY, SYS_MOVEMENT

        code = {"gravity": SYS_GRAVITY, "movement": SYS_MOVEMENT, "acceleration": SYS_ACCELERATION}
        order, _ = app.last_tick_order()
        names = [app.system_names[i] for i in order if app.system_names[i] in code]
        if not with_cpu_reader:
            assert names == ["gravity", "movement", "acceleration"]
        nbody_oracle.run_sequential(ref[0], ref[1], ref[2], ref[3], ticks, order=tuple(code[x] for x in names))
        assert rel_err_rows(w.column(1, POS)[:n], ref[0], floor=1.0).max() <= 1e-4, names
        assert rel_err_rows(w.column(1, VEL)[:n], ref[2], floor=1.0).max() <= 1e-4, names
        assert rel_err_rows(w.column(1, ACC)[:n], ref[3]).max() <= 1e-3, names
        app.close()


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1000, 6000])
def test_application_of_gpu_systems_only_hands_the_remaining_ticks_over_in_one_call(gpu_ctx, nbody_oracle, n):
    """An application whose every system is a GPU system of one chain and cannot record commands repeats the same chain
    every tick: recs_app_run gives ticks 2..n to the context in one recs_gpu_tick call (CUDA graph, or for a small n-body
    world ONE cooperative launch).  Columns equal the tick-by-tick run bit for bit."""
    ticks = 12
    outs, launches = [], []
    for one_by_one in (False, True):
        gpu_ctx.reset_timing()
        app = ecs.Application(gpu=gpu_ctx)
        w = app.world
        for c, dt, name in ((POS, F3, "Position"), (MASS, np.float32, "Mass"), (VEL, F3, "Velocity"), (ACC, F3, "Acceleration")):
            w.register_component(c, dt, name)
        pos, mass, vel, acc = scenes.spawn_bodies(scenes.all_heavy_random_cube_with_bodies(n), 5)
        app.add_gpu_system(_lib.SYS_NBODY_MOVEMENT, [POS, VEL], "movement")
        app.add_gpu_system(_lib.SYS_NBODY_ACCELERATION, [VEL, ACC], "acceleration")
        app.add_gpu_system(_lib.SYS_NBODY_GRAVITY, [POS, ACC, MASS], "gravity")
        w.create_entities(n, {POS: pos, MASS: mass, VEL: vel, ACC: acc})
        if one_by_one:
            for _ in range(ticks):
                app.run(1)
        else:
            app.run(ticks)
        launches.append(gpu_ctx.timing()["kernel_launches"])
        order, _ = app.last_tick_order()
        assert [app.system_names[i] for i in order] == ["gravity", "movement", "acceleration"]
        app.sync_to_host()
        outs.append([w.column(1, c)[:n].copy() for c in (POS, VEL, ACC)])
        app.close()
    for a, b in zip(*outs):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    assert launches[1] >= (ticks if n <= 4096 else 2 * ticks)  # tick by tick: one launch per tick for a small world, two otherwise
    if n <= 4096:
        assert launches[0] <= 6, launches  # pack, tick 1, tick 2 (first of the handed-over call), one launch for the rest


@pytest.mark.gpu
def test_hand_over_survives_structural_changes_between_runs(gpu_ctx, nbody_oracle):
    """run(k) hands ticks 2..k to the context in one call; entities created and removed between run calls change the
    archetype under it (columns move, the mirror grows and shrinks, the resident launch's second mirror buffer with it).
    The same program tick by tick must give the same columns bit for bit."""
    outs = []
    for one_by_one in (False, True):
        app = ecs.Application(gpu=gpu_ctx)
        w = app.world
        for c, dt, name in ((POS, F3, "Position"), (MASS, np.float32, "Mass"), (VEL, F3, "Velocity"), (ACC, F3, "Acceleration")):
            w.register_component(c, dt, name)
        app.add_gpu_system(_lib.SYS_NBODY_MOVEMENT, [POS, VEL], "movement")
        app.add_gpu_system(_lib.SYS_NBODY_ACCELERATION, [VEL, ACC], "acceleration")
        app.add_gpu_system(_lib.SYS_NBODY_GRAVITY, [POS, ACC, MASS], "gravity")
        run = (lambda k: [app.run(1) for _ in range(k)]) if one_by_one else app.run
        first = scenes.spawn_bodies(scenes.all_heavy_random_cube_with_bodies(900), 6)
        w.create_entities(900, {POS: first[0], MASS: first[1], VEL: first[2], ACC: first[3]})
        run(7)
        more = scenes.spawn_bodies(scenes.all_heavy_random_cube_with_bodies(400), 7)   # 1 300 bodies: one more mirror tile
        app.sync_to_host()
        w.create_entities(400, {POS: more[0], MASS: more[1], VEL: more[2], ACC: more[3]})
        run(6)
        app.sync_to_host()
        for k in (5, 17, 1200, 640):
            app.command_remove(ecs.Entity(*w.archetype_entities(1)[k]))
        run(1)   # playback removes them after this tick
        run(9)
        big = scenes.spawn_bodies(scenes.all_heavy_random_cube_with_bodies(4000), 8)   # 5 296 bodies: beyond the resident limit
        app.sync_to_host()
        w.create_entities(4000, {POS: big[0], MASS: big[1], VEL: big[2], ACC: big[3]})
        run(4)
        app.sync_to_host()
        n = len(w.archetype_entities(1))
        assert n == 900 + 400 - 4 + 4000
        outs.append([w.column(1, c)[:n].copy() for c in (POS, VEL, ACC)])
        app.close()
    for a, b in zip(*outs):
        assert np.all(np.isfinite(a)) and np.array_equal(a.view(np.uint32), b.view(np.uint32))
