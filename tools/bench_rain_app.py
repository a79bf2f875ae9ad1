"""The reference's rain_simulation benchmark (crates/ecs_bench_suite/src/recs/rain_simulation.rs: all seven systems, clouds
from create_evenly_interspersed_clouds, >= 100 ticks) with the WHOLE tick on the device (recs_b200/rain.py); the host World
plays the commands back every tick and the resident columns follow by appends / row moves.  Wall-clock per tick (host +
device).  python tools/bench_rain_app.py [clouds,...] [ticks]      (RECS_APP_PROFILE=1: host phases)"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from recs_b200 import GpuContext  # noqa: E402
from recs_b200.rain import rain_application  # noqa: E402

sizes = [int(a) for a in sys.argv[1].split(",")] if len(sys.argv) > 1 else [1024, 16384]
ticks = int(sys.argv[2]) if len(sys.argv) > 2 else 100
for n_clouds in sizes:
    with GpuContext(device=0) as ctx:
        app = rain_application(ctx, n_clouds)
        w = app.world
        app.run(100)  # reach the steady state: drops are landing
        ctx.synchronize()
        t0 = time.perf_counter()
        created = removed = 0
        for _ in range(ticks):
            app.run(1)
            c, r = app.last_playback()
            created += c
            removed += r
        ctx.synchronize()
        sec = time.perf_counter() - t0
        drops = len(w.archetype_entities(2)) if w.archetype_count() > 2 else 0
        print(f"clouds={len(w.archetype_entities(1)):6d}: {sec / ticks * 1e3:8.3f} ms per tick over {ticks} ticks (wall clock, host playback included); "
              f"{drops} drops alive, {created / ticks:.0f} created and {removed / ticks:.0f} removed per tick", flush=True)
        app.close()
