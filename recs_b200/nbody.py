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

    # what a tick reads before writing it: gravity overwrites Acceleration (`*acceleration = total_acceleration`,
    # crates/n-body/src/lib.rs:134), so it never has to go up
    INPUT_COMPONENTS = (POSITION, MASS, VELOCITY)

    def upload(self, rows=None, comps=(POSITION, MASS, VELOCITY, ACCELERATION)) -> None:
        """Columns host -> device (default all four); `rows=(begin, end)` moves only that row range (a rank's own shard)."""
        for comp in comps:
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


# `gravity` as a GENERIC nested-query system (recs_gpu_system_create_pair_kernel): crates/n-body/src/lib.rs:104-135 one
# expression per line.  Folded in query order with IEEE division / square root and no FMA contraction, it reproduces
# the reference's sequential f32 sum bit for bit -- the exactness cross-check of the hand-written kernel (about 4.5x
# slower than it).
REFERENCE_ORDER_GRAVITY_SOURCE = """
struct Position { float x, y, z; };
struct Mass { float m; };
struct Acceleration { float x, y, z; };
struct Total { float x, y, z; };                                   // Vector3::sum() starts from zero
void pair(Total& total, const Position& position, const Position& body_position, const Mass& body_mass) {
    const float tx = body_position.x - position.x;                 // let to_body = body_position.point - position.point;
    const float ty = body_position.y - position.y;
    const float tz = body_position.z - position.z;
    const float distance_squared = (tx * tx + ty * ty) + tz * tz;  // to_body.magnitude2()
    float cx = 0.0f, cy = 0.0f, cz = 0.0f;                         // if distance_squared <= f32::EPSILON { return zero }
    if (!(distance_squared <= 1.1920929e-7f)) {
        const float acceleration = (6.67e-11f * body_mass.m) / distance_squared;
        const float scale = acceleration / sqrtf(distance_squared);  // normalize_to(m) = self * (m / self.magnitude())
        cx = tx * scale; cy = ty * scale; cz = tz * scale;
    }
    total.x += cx; total.y += cy; total.z += cz;
}
// The same pair for the generated kernel's speculative fast fold: branch-free, the divisions and the square root as the
// in-range instruction sequences (recs_div_rn_seq / recs_sqrt_rn_seq: bit-identical to `/` and sqrtf in range); returns
// false when this pair needs the exact body above (then the whole tile is folded again with it).
bool pair_fast(Total& total, const Position& position, const Position& body_position, const Mass& body_mass) {
    const float tx = body_position.x - position.x;
    const float ty = body_position.y - position.y;
    const float tz = body_position.z - position.z;
    const float distance_squared = (tx * tx + ty * ty) + tz * tz;
    const bool masked = distance_squared <= 1.1920929e-7f;
    const float gm = 6.67e-11f * body_mass.m;
    const float acceleration = recs_div_rn_seq(gm, distance_squared);
    const float scale = recs_div_rn_seq(acceleration, recs_sqrt_rn_seq(distance_squared));
    total.x += masked ? 0.0f : tx * scale;
    total.y += masked ? 0.0f : ty * scale;
    total.z += masked ? 0.0f : tz * scale;
    // (distance_squared within 2^-63 .. 2^63 puts its square root inside both the square root's and the division's range)
    const bool in_range = recs_div_in_range(gm) & recs_div_in_range(distance_squared) & recs_div_in_range(acceleration);
    return in_range | masked;  // a masked pair contributes zero whatever the discarded quotient was
}
void finish(const Position& position, Acceleration& acceleration, const Total& total) {
    acceleration.x = total.x; acceleration.y = total.y; acceleration.z = total.z;   // *acceleration = total_acceleration
}
"""
REFERENCE_ORDER_GRAVITY_PARAMS = [
    (POSITION, 12, "read", "Position"),
    (ACCELERATION, 12, "write", "Acceleration"),
    (POSITION, 12, "query_read", "Position"),
    (MASS, 4, "query_read", "Mass"),
]


def reference_order_gravity(ctx: GpuContext, outer_pos: np.ndarray, qpos: np.ndarray, qmass: np.ndarray,
                            outer_archetype: int = 1001, query_archetype: int = 1002) -> np.ndarray:
    """Accelerations of `outer_pos` rows against the bodies (qpos, qmass), computed on the device by the generic
    reference-order system above.  Single-GPU contexts."""
    acc = np.zeros_like(outer_pos)
    for arch, comp, arr in ((outer_archetype, POSITION, outer_pos), (outer_archetype, ACCELERATION, acc),
                            (query_archetype, POSITION, qpos), (query_archetype, MASS, qmass)):
        ctx.bind_column(arch, comp, arr)
        ctx.upload(arch, comp)
    sid = ctx.create_pair_kernel_system("gravity_reference_order", REFERENCE_ORDER_GRAVITY_SOURCE, "Total", "pair", "finish",
                                        REFERENCE_ORDER_GRAVITY_PARAMS)
    ctx.set_archetypes(sid, [outer_archetype])
    ctx.set_query_archetypes(sid, [query_archetype])
    ctx.run_system(sid)
    ctx.download(outer_archetype, ACCELERATION)
    return acc


class ReferenceOrderNBodyApp(NBodyGpuApp):
    """The same three systems with `gravity` replaced by the generic reference-order system above: the reference's own
    arithmetic and summation order, executed on the device (bit-identical to the CPU oracle; about 4.5x slower than the
    hand-written kernel).  Used where the reference trajectory is needed at sizes the CPU oracle takes minutes for."""

    def __init__(self, ctx: GpuContext, pos, mass, vel, acc, tick_order=("gravity", "movement", "acceleration")):
        super().__init__(ctx, pos, mass, vel, acc, tick_order)
        self.systems["gravity"] = ctx.create_pair_kernel_system(
            "gravity_reference_order", REFERENCE_ORDER_GRAVITY_SOURCE, "Total", "pair", "finish", REFERENCE_ORDER_GRAVITY_PARAMS)
        ctx.set_archetypes(self.systems["gravity"], [BODY_ARCHETYPE])
        ctx.set_query_archetypes(self.systems["gravity"], [BODY_ARCHETYPE])
        self.set_tick_order(tick_order)
