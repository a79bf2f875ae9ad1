"""The boundary's threading contract: the reference runs a system's task on an arbitrary worker thread,
different each tick (crates/scheduler/src/executor/worker.rs:143-154), so every entry point must be callable from
any host thread -- cudaSetDevice inside each call, no thread-local CUDA state, one mutex per context."""
import threading

import numpy as np
import pytest

from recs_b200 import scenes
from tests.util import rel_err_rows

pytestmark = pytest.mark.gpu


def test_ticks_issued_from_a_different_thread_each_time(gpu_ctx, nbody_oracle):
    from recs_b200.nbody import NBodyGpuApp

    pos, mass, vel, acc = scenes.spawn_bodies(scenes.all_heavy_random_cube_with_bodies(3000), 4)
    ref = [a.copy() for a in (pos, mass, vel, acc)]
    app = NBodyGpuApp(gpu_ctx, pos, mass, vel, acc)
    errors = []

    def step(what):
        try:
            what()
        except Exception as exc:  # pragma: no cover - reported below
            errors.append(exc)

    ops = [app.upload] + [lambda name=name: app.run(name) for _ in range(4) for name in ("gravity", "movement", "acceleration")] + [app.download]
    for op in ops:  # one fresh thread per call, strictly one after the other like the schedule's batches
        t = threading.Thread(target=step, args=(op,))
        t.start()
        t.join()
    assert not errors, errors
    nbody_oracle.run_sequential(ref[0], ref[1], ref[2], ref[3], 4)
    assert rel_err_rows(pos, ref[0], floor=1.0).max() <= 1e-4
    assert rel_err_rows(vel, ref[2], floor=1.0).max() <= 1e-4


def test_concurrent_callers_are_serialised_per_context(gpu_ctx):
    """Eight threads hammer the same context with independent systems (disjoint columns): the context mutex
    serialises them, every system runs exactly as often as it was called."""
    from recs_b200 import _lib

    ctx = gpu_ctx
    n, calls = 20000, 25
    cols, sids = [], []
    for k in range(8):
        p = np.zeros((n, 3), np.float32)
        v = np.full((n, 3), float(k + 1), np.float32)
        cols.append((p, v))
        ctx.bind_column(10 + k, 1, p)
        ctx.bind_column(10 + k, 2, v)
        ctx.upload(10 + k, 1)
        ctx.upload(10 + k, 2)
        sid = ctx.create_system(_lib.SYS_QUERY_ADD, [1, 2])
        ctx.set_archetypes(sid, [10 + k])
        sids.append(sid)
    errors = []

    def worker(k):
        try:
            for _ in range(calls):
                ctx.run_system(sids[k], sync=False)
        except Exception as exc:  # pragma: no cover
            errors.append(exc)

    threads = [threading.Thread(target=worker, args=(k,)) for k in range(8)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    ctx.synchronize()
    for k in range(8):
        ctx.download(10 + k, 1)
        assert np.array_equal(cols[k][0], np.full((n, 3), float((k + 1) * calls), np.float32))


def test_two_contexts_on_one_device_from_two_threads(nbody_oracle):
    from recs_b200 import GpuContext
    from recs_b200.nbody import NBodyGpuApp

    results, errors = {}, []

    def run(seed):
        try:
            pos, mass, vel, acc = scenes.spawn_bodies(scenes.all_heavy_random_cube_with_bodies(2000), seed)
            host = [a.copy() for a in (pos, mass, vel, acc)]
            with GpuContext(device=0) as ctx:
                app = NBodyGpuApp(ctx, pos, mass, vel, acc)
                app.upload()
                app.tick(5)
                app.download()
            results[seed] = (pos, vel, host)
        except Exception as exc:  # pragma: no cover
            errors.append(exc)

    threads = [threading.Thread(target=run, args=(s,)) for s in (21, 22)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for seed, (pos, vel, host) in results.items():
        nbody_oracle.run_sequential(host[0], host[1], host[2], host[3], 5)
        assert rel_err_rows(pos, host[0], floor=1.0).max() <= 1e-4
        assert rel_err_rows(vel, host[2], floor=1.0).max() <= 1e-4
