"""The n-body case study (crates/n-body) on the GPU hot path: component ids, system
registration and the archetype the benchmark builds.

`NBodyGpuApp` is the slice of `ecs_bench_suite::recs::n_body::Benchmark::run`
(crates/ecs_bench_suite/src/recs/n_body.rs:29-75) that touches the hot path: it registers the
three systems in the benchmark's order, creates the bodies (entity k -> archetype 1, component
index k) and ticks them in the order the schedule produced.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from .gpu import GpuContext

# TypeId stand-ins for the component types of crates/n-body/src/lib.rs:6,43,53,70
POSITION, MASS, VELOCITY, ACCELERATION = 1, 2, 3, 4
BODY_ARCHETYPE = 1  # archetype 0 is the empty-entity archetype (crates/ecs/src/archetypes.rs:185)


class NBodyGpuApp:
    def __init__(self, ctx: GpuContext, pos, mass, vel, acc, tick_order=("gravity", "movement", "acceleration")):
        self.ctx = ctx
        n = pos.shape[0]
        assert pos.shape == (n, 3) and vel.shape == (n, 3) and acc.shape == (n, 3) and mass.shape == (n,)
        self.pos, self.mass, self.vel, self.acc = pos, mass, vel, acc
        self.n = n
        # registration order of the benchmark: movement, acceleration, gravity (n_body.rs:54-58)
        self.systems = {
            "movement": ctx.create_system(_lib.SYS_NBODY_MOVEMENT, [POSITION, VELOCITY]),
            "acceleration": ctx.create_system(_lib.SYS_NBODY_ACCELERATION, [VELOCITY, ACCELERATION]),
            "gravity": ctx.create_system(_lib.SYS_NBODY_GRAVITY, [POSITION, ACCELERATION, MASS]),
        }
        for comp, arr in ((POSITION, pos), (MASS, mass), (VELOCITY, vel), (ACCELERATION, acc)):
            ctx.bind_column(BODY_ARCHETYPE, comp, arr)
        for sid in self.systems.values():
            ctx.set_archetypes(sid, [BODY_ARCHETYPE])
        ctx.set_query_archetypes(self.systems["gravity"], [BODY_ARCHETYPE])
        self.set_tick_order(tick_order)

    def set_tick_order(self, names) -> None:
        self.tick_order = tuple(names)
        self.ctx.set_tick_order([self.systems[name] for name in names])

    def upload(self, rows=None) -> None:
        """All four columns host -> device; `rows=(begin, end)` moves only that row range (a rank's own shard)."""
        for comp in (POSITION, MASS, VELOCITY, ACCELERATION):
            if rows is None:
                self.ctx.upload(BODY_ARCHETYPE, comp)
            else:
                self.ctx.upload_rows(BODY_ARCHETYPE, comp, rows[0], rows[1])

    def download(self, comps=(POSITION, VELOCITY, ACCELERATION), rows=None) -> None:
        for comp in comps:
            if rows is None:
                self.ctx.download(BODY_ARCHETYPE, comp)
            else:
                self.ctx.download_rows(BODY_ARCHETYPE, comp, rows[0], rows[1])

    def run(self, name: str, sync: bool = True) -> None:
        self.ctx.run_system(self.systems[name], sync=sync)

    def tick(self, n: int = 1) -> None:
        self.ctx.tick(n)
