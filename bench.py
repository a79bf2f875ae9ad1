#!/usr/bin/env python
"""Benchmark of the recs n-body hot path on B200 (driver contract: one JSON line on rank 0).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--bodies B] [--impl reference]

A "step" is one tick of the benchmark's three systems in schedule order (gravity -> movement ->
acceleration) over B synthetic bodies drawn from the reference's EVERYTHING_HEAVY
distributions (crates/n-body/src/scenes.rs:10-21) with a seeded generator.  Metric: body
interactions per second, one interaction = one ordered pair (i, j) including j == i
(SURVEY 8d), i.e. B*B per step.

  value      device-resident: columns already in HBM, K ticks, CUDA-event time per tick summed,
             L2 flushed between ticks outside the events; max over ranks.
  e2e        the same tick through the C ABI with HOST (pinned) columns: upload of all four
             columns, tick, download of Position/Velocity/Acceleration inside the timed region.
  roofline   the interaction kernel alone (CUDA events around its launches) against the FP32 FMA
             rate measured live on the same GPU with a register-resident FFMA loop.
  cpu_baseline / --impl reference
             oracle/nbody_oracle.c (C restatement of the reference's CPU systems driven with its
             worker-pool policy) on the box's host cores, on a bounded row sample of the same
             workload.  The Rust reference itself cannot be built here (no cargo/rustc).

N > 1 (torchrun, one rank per GPU): the B bodies are index-sharded, each rank computes its rows
and positions are exchanged with an NCCL all-gather every tick ("scaling": "strong").
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOPS_PER_INTERACTION = 20.0  # all-pairs convention; the kernel issues 12 FMA-pipe lane-ops
FMA_LANE_OPS_PER_INTERACTION = 12.0
# dram__bytes_read.sum + dram__bytes_write.sum of one interaction-kernel launch on one GPU, from ncu
# (profiles/r01_gravity_traffic_1M.csv, profiles/r01_gravity_default_ncu_summary.txt); the kernel is
# FP32-bound, the figure only shows there are no wasted re-reads (posm + outer positions = 28 MB at 1 M)
KERNEL_DRAM_TRAFFIC_BYTES = {1_048_576: 29_457_664 + 49_920, 65_536: 1_854_208}
DEFAULT_BODIES = 1_048_576
SEED = 1


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--bodies", type=int, default=DEFAULT_BODIES)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    return ap.parse_args()


# --------------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    QUERY = (
        "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
        "clocks_event_reasons.sw_power_cap"
    )

    def __init__(self, device: int):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE,
                stderr=subprocess.DEVNULL,
                text=True,
            )
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
                power.append(float(parts[2]))
            except ValueError:
                continue
            for name, flag in zip(names, parts[3:7]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {
            "sm_mhz": float(np.median(sm)),
            "sm_max_mhz": float(max(mx)),
            "power_w_max": float(max(power)) if power else None,
            "samples": len(sm),
            "reasons": sorted(reasons),
        }


# --------------------------------------------------------------------------------------------------
# CPU side (oracle) -- reported baseline and the --impl reference arm
# --------------------------------------------------------------------------------------------------
def cpu_sample_rate(n_bodies: int, budget_s: float, threads: int):
    """Times one tick of the oracle's worker pool on rows [0, S) x all n_bodies j (a bounded row
    sample of the workload).  Returns (interactions/s, S, seconds)."""
    import oracle
    from recs_b200 import scenes

    orc = oracle.load_nbody()
    pos, mass, vel, acc = scenes.spawn_bodies(scenes.all_heavy_random_cube_with_bodies(n_bodies), SEED)
    # calibrate on a small row block
    cal_rows = min(n_bodies, max(64, threads * 16))
    t0 = time.perf_counter()
    _pool_rows(orc, pos, mass, vel, acc, cal_rows, threads)
    cal = time.perf_counter() - t0
    rate = cal_rows * n_bodies / max(cal, 1e-9)
    rows = int(min(n_bodies, max(cal_rows, rate * budget_s / n_bodies)))
    rows = max(threads * 4 * 16, rows // (threads * 4) * (threads * 4)) if rows < n_bodies else rows
    rows = min(rows, n_bodies)
    t0 = time.perf_counter()
    _pool_rows(orc, pos, mass, vel, acc, rows, threads)
    sec = time.perf_counter() - t0
    return rows * n_bodies / sec, rows, sec, (orc, pos, mass, vel, acc)


def _pool_rows(orc, pos, mass, vel, acc, rows, threads):
    """One tick (gravity -> movement -> acceleration) of the reference's executor policy on the
    first `rows` outer rows; the nested query still iterates all bodies."""
    import ctypes as C

    n = pos.shape[0]
    if rows >= n:
        orc.run_pool(pos, mass, vel, acc, 1, threads)
        return
    # outer iteration restricted to `rows` rows: columns of the sampled archetype slice + full query
    lib = orc.lib
    if not hasattr(lib, "_sample_ready"):
        z = C.c_size_t
        f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
        lib.oracle_nbody_run_pool_rows.argtypes = [f32p, f32p, f32p, f32p, z, z, C.c_int, C.c_int, C.c_int * 3]
        lib.oracle_nbody_run_pool_rows.restype = C.c_double
        lib._sample_ready = True
    lib.oracle_nbody_run_pool_rows(pos, mass, vel, acc, n, rows, 1, threads, (C.c_int * 3)(0, 1, 2))


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def run_reference_arm(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = host_threads()
    steps = args.steps or 3
    per_step_budget = max(1.0, min(20.0, 90.0 / (steps + args.warmup)))
    rate, rows, sec, state = cpu_sample_rate(args.bodies, per_step_budget, threads)
    orc, pos, mass, vel, acc = state
    for _ in range(max(args.warmup - 1, 0)):
        _pool_rows(orc, pos, mass, vel, acc, rows, threads)
    t0 = time.perf_counter()
    for _ in range(steps):
        _pool_rows(orc, pos, mass, vel, acc, rows, threads)
    total = time.perf_counter() - t0
    value = rows * args.bodies * steps / total / 1e9
    sample = f"{rows} of {args.bodies} outer rows x all {args.bodies} bodies per step (gravity+movement+acceleration on the sampled rows)"
    line = {
        "impl": "reference",
        "metric": "nbody_body_interactions_per_second",
        "value": value,
        "unit": "G interactions/s",
        "n_gpus": args.gpus,
        "steps": steps,
        "warmup": args.warmup,
        "ms_per_step": total / steps * 1e3,
        "higher_is_better": True,
        "scaling": "strong",
        "vs_baseline": None,
        "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"nbody_{args.bodies}", "bodies": args.bodies, "sample": sample, "host_threads": threads},
        "cpu_baseline": {
            "value": value,
            "unit": "G interactions/s",
            "cores": threads,
            "kind": "port",
            "sample": sample,
            "note": "C restatement of recs' CPU systems + worker-pool segment policy (oracle/nbody_oracle.c); the Rust reference cannot be built here",
        },
        "e2e": {"value": value, "unit": "G interactions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
def sampled_parity(app, host_state, rb, re, n_rows=32):
    """Parity gate of the run: accelerations of sampled rows of this rank's shard after
    the first tick against the oracle, all j.  Returns max norm-wise relative errors
    (gpu vs reference-order f32, gpu vs f64, reference-order f32 vs f64)."""
    import oracle

    orc = oracle.load_nbody()
    pos0, mass0 = host_state
    rows = np.random.default_rng(11).choice(np.arange(rb, re), size=min(n_rows, re - rb), replace=False)
    ref32, ref64 = orc.gravity_sample(pos0, mass0, rows)
    got = app.acc[rows].astype(np.float64)

    def rel(a, b):
        return float((np.linalg.norm(a - b, axis=1) / np.maximum(np.linalg.norm(b, axis=1), 1e-30)).max())

    # the same rows through the generic reference-order system (recs_b200/nbody.py): sequential fold in query order,
    # IEEE division and square root -- must equal the oracle's f32 result bit for bit, at full problem size
    exact = None
    try:
        from recs_b200 import GpuContext
        from recs_b200.nbody import reference_order_gravity

        with GpuContext(device=int(os.environ.get("LOCAL_RANK", "0")), use_cuda_graph=False) as check_ctx:
            got_exact = reference_order_gravity(check_ctx, np.ascontiguousarray(pos0[rows]), pos0, mass0)
        exact = bool(np.array_equal(got_exact.view(np.uint32), ref32.view(np.uint32)))
    except Exception as exc:  # the check never invalidates the headline line
        exact = repr(exc)
    return rel(got, ref32.astype(np.float64)), rel(got, ref64), rel(ref32.astype(np.float64), ref64), exact


def run_gpu_arm(args) -> None:
    from recs_b200 import GpuContext, scenes
    from recs_b200.nbody import ACCELERATION, BODY_ARCHETYPE, MASS, POSITION, VELOCITY, NBodyGpuApp

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n = args.bodies
    steps = args.steps or (10 if n >= 500_000 else 50)
    warmup = max(args.warmup, 3)

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod

        dist = dist_mod
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    ctx = GpuContext(device=local_rank, rank=rank, world_size=world, use_cuda_graph=True)
    if world > 1:
        uid = [ctx.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(uid[0])
    rb, re = ctx.shard_rows(n)

    # synthetic bodies in pinned host columns (the host side of the e2e path)
    pos0, mass0, vel0, acc0 = scenes.spawn_bodies(scenes.all_heavy_random_cube_with_bodies(n), SEED)
    pos, mass, vel, acc = (ctx.pinned_empty(a.shape) for a in (pos0, mass0, vel0, acc0))
    for dst, src in ((pos, pos0), (mass, mass0), (vel, vel0), (acc, acc0)):
        dst[...] = src
    app = NBodyGpuApp(ctx, pos, mass, vel, acc)

    def barrier():
        ctx.synchronize()
        if dist is not None:
            import torch

            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if dist is None:
            return x
        import torch

        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- parity gate: first tick against the oracle on sampled rows -----------------------------
    app.upload()
    app.tick(1)
    app.download()
    perr = sampled_parity(app, (pos0, mass0), rb, re)
    parity_err, parity_f64, oracle_f64 = (max_over_ranks(x) for x in perr[:3])
    reference_order_exact = perr[3]
    # positions after tick 1 must be bit-exact (movement precedes acceleration)
    import oracle

    orc = oracle.load_nbody()
    pref = pos0.copy()
    orc.axpy(pref, vel0, orc.dt, rows=(rb, re))
    pos_exact = bool(np.array_equal(app.pos[rb:re].view(np.uint32), pref[rb:re].view(np.uint32)))

    # ---- device-resident timed region ------------------------------------------------------------
    for _ in range(warmup):
        app.tick(1)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ctx.reset_timing()
    step_ms = []
    wall0 = time.perf_counter()
    for _ in range(steps):
        ctx.flush_l2()
        ctx.timer_start()
        app.tick(1)
        step_ms.append(ctx.timer_stop())
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop()
    total_ms = max_over_ranks(float(np.sum(step_ms)))
    per_rank_ms = [float(np.mean(step_ms))]
    if dist is not None:
        gathered = [None] * world
        dist.all_gather_object(gathered, float(np.mean(step_ms)))
        per_rank_ms = gathered
    launches_per_step = ctx.timing()["kernel_launches"] / steps
    value = n * n * steps / (total_ms * 1e-3) / 1e9

    # ---- end to end: host columns in, host columns out, every step --------------------------------
    # A rank moves the rows it owns (all rows on one GPU): the other ranks' positions and masses reach it through the
    # all-gather, and it never touches their velocities or accelerations.
    e2e_steps = steps
    own = None if world == 1 else (rb, re)
    for _ in range(2):
        app.upload(rows=own)
        app.tick(1)
        app.download(rows=own)
    barrier()
    e0 = time.perf_counter()
    for _ in range(e2e_steps):
        app.upload(rows=own)
        app.tick(1)
        app.download(rows=own)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - e0)
    e2e_value = n * n * e2e_steps / e2e_s / 1e9
    # whole-job bytes per step, counted from what is copied: 40 B up and 36 B down per body
    h2d = int(pos.nbytes + mass.nbytes + vel.nbytes + acc.nbytes)
    d2h = int(pos.nbytes + vel.nbytes + acc.nbytes)

    # ---- roofline of the interaction kernel (events around its launches only) --------------------
    tf_peak, mhz = ctx.measure_fp32_peak(packed=True)
    ctx.enable_timing(True)
    ctx.reset_timing()
    k_steps = max(3, min(steps, 10))
    for _ in range(k_steps):
        app.tick(1)
    t = ctx.timing()
    ctx.enable_timing(False)
    grav_ms = t["gravity_ms"] / k_steps  # all interaction-kernel launches of one tick on this rank
    rows_here = re - rb
    kernel_flops = FLOPS_PER_INTERACTION * rows_here * n
    achieved_tf = kernel_flops / (grav_ms * 1e-3) / 1e12
    fma_lane_rate = FMA_LANE_OPS_PER_INTERACTION * rows_here * n / (grav_ms * 1e-3)
    roofline = {
        "bound": "fp32",
        "kernel": "nbody_gravity_kernel",
        "achieved": achieved_tf,
        "peak": tf_peak,
        "unit": "TFLOP/s",
        "frac": achieved_tf / tf_peak,
        "peak_source": "measured live on this GPU: register-resident FFMA2 loop (MEASURED_PEAKS.json has no FP32 figure)",
        "flops_per_interaction": FLOPS_PER_INTERACTION,
        "fma_pipe_frac": fma_lane_rate / (tf_peak * 1e12 / 2.0),
        "kernel_ms_per_step": grav_ms,
        "kernel_share_of_step": grav_ms / (t["last_tick_chain_ms"] / k_steps) if t["last_tick_chain_ms"] else None,
        "traffic": KERNEL_DRAM_TRAFFIC_BYTES.get(n) if world == 1 else None,
        "traffic_unit": "bytes of DRAM traffic per launch (ncu, single GPU)",
        "sm_mhz_during_peak_probe": mhz,
    }

    comm = None
    if world > 1:
        comm = {
            "collective": "ncclAllGather of the pair-interleaved position/mass mirror, in place, once per tick",
            "bytes_per_rank_per_step": int((re - rb) * 16),
            "bytes_total_per_step": int(n * 16),
            "gather_ms_per_step_incl_peer_wait": t["gather_ms"] / k_steps,
            "overlap": "enqueued after the local-tile launch; never on the critical path before it",
        }

    # ---- secondary configs (streaming systems), reported as extras ---------------------------------
    extras = {}
    if world > 1:
        extras = {"skipped": "the streaming configs are reported by the N=1 run (a sharded context streams only its own slice)"}
    elif not args.no_extras and rank == 0:
        try:
            extras = streaming_extras(ctx)
        except Exception as exc:  # extras never invalidate the headline line
            extras = {"error": repr(exc)}

    # ---- CPU baseline (rank 0, bounded sample) ---------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:  # the contract: rank 0 at N=1 only
        threads = host_threads()
        rate, rows, sec, _ = cpu_sample_rate(n, 12.0, threads)
        cpu = {
            "value": rate / 1e9,
            "unit": "G interactions/s",
            "cores": threads,
            "kind": "port",
            "sample": f"{rows} of {n} outer rows x all {n} bodies, one tick, {sec:.1f} s",
            "note": "oracle/nbody_oracle.c: C restatement of recs' CPU systems + worker-pool segment policy",
        }

    if rank == 0:
        line = {
            "metric": "nbody_body_interactions_per_second",
            "value": value,
            "unit": "G interactions/s",
            "n_gpus": world,
            "steps": steps,
            "warmup": warmup,
            "ms_per_step": total_ms / steps,
            "higher_is_better": True,
            "scaling": "strong",
            "vs_baseline": None,
            "dtype": "f32",
            "data": "synthetic",
            "config": {
                "workload": f"nbody_{n}",
                "bodies": n,
                "systems": "gravity->movement->acceleration",
                "parallelism": f"index-sharded x{world}, NCCL all-gather of positions per tick" if world > 1 else "single GPU",
                "l2": "flushed between timed ticks (256 MiB write + 256 MiB read sweep outside the event bracket)",
                "seed": SEED,
            },
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "G interactions/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_s / e2e_steps * 1e3},
            "gpu_launches": int(round(launches_per_step * steps)),
            "roofline": roofline,
            "cpu_baseline": cpu,
            "parity": {"acc_max_rel_err_vs_oracle": parity_err, "tolerance": 1e-4, "pos_bit_exact_after_tick1": pos_exact,
                       "acc_max_rel_err_vs_f64": parity_f64, "oracle_f32_order_vs_f64": oracle_f64, "rows_per_rank": 32,
                       "reference_order_generic_kernel_bit_exact": reference_order_exact},
            "step_ms": [round(x, 3) for x in step_ms],
            "ms_per_step_by_rank": [round(x, 3) for x in per_rank_ms],
            "wall_ms_per_step_incl_flush": wall / steps * 1e3,
            "comm": comm,
            "extras": extras,
        }
        print(json.dumps(line), flush=True)
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


def streaming_extras(ctx) -> dict:
    """HBM-bound systems of configs[3] / configs[4]: achieved GB/s on algorithmic bytes."""
    from recs_b200 import _lib, scenes

    out = {}
    peak = None
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6650.0  # B200_PROFILING.md fallback
    # simple_iter: 10 M entities, pos += vel, 36 B / entity / pass
    n = 10_000_000
    A, B, ARCH = 101, 102, 7
    p = np.ones((n, 3), np.float32)
    v = np.ones((n, 3), np.float32)
    ctx.bind_column(ARCH, A, p)
    ctx.bind_column(ARCH, B, v)
    ctx.upload(ARCH, A)
    ctx.upload(ARCH, B)
    sid = ctx.create_system(_lib.SYS_QUERY_ADD, [A, B])
    ctx.set_archetypes(sid, [ARCH])
    for _ in range(3):
        ctx.run_system(sid, sync=False)
    # (a) steady state: 20 passes back to back in one event bracket.  One pass touches 240 MB, more than the
    #     126 MB L2, so every pass is cold and also pays for the write-backs the previous one left in L2.
    passes = 20
    ctx.synchronize()
    ctx.timer_start()
    for _ in range(passes):
        ctx.run_system(sid, sync=False)
    steady_ms = ctx.timer_stop() / passes
    # (b) single passes after an L2 flush that leaves clean lines
    ms = []
    for _ in range(20):
        ctx.flush_l2()
        ctx.timer_start()
        ctx.run_system(sid, sync=False)
        ms.append(ctx.timer_stop())
    gbs = 36.0 * n / (steady_ms * 1e-3) / 1e9
    out["simple_iter_10M"] = {
        "GBps": gbs,
        "frac_of_measured_hbm": gbs / peak,
        "ms": float(steady_ms),
        "bytes_per_entity": 36,
        "how": "20 passes back to back in one CUDA-event bracket; inputs (240 MB) larger than L2",
        "single_pass_after_flush_ms": float(np.median(ms)),
        "single_pass_after_flush_GBps": 36.0 * n / (np.median(ms) * 1e-3) / 1e9,
    }
    # the same system as a GENERIC GPU system: its body handed over as source, compiled with NVRTC at registration,
    # run through the TMA-pipelined generated kernel (recs_gpu_system_create_kernel)
    body = ("struct Position { float x, y, z; };\nstruct Velocity { float x, y, z; };\n"
            "void movement(Position& p, const Velocity& v) { p.x += v.x; p.y += v.y; p.z += v.z; }\n")
    gid = ctx.create_kernel_system("simple_iter", body, "movement", [(A, 12, "write", "Position"), (B, 12, "read", "Velocity")])
    ctx.set_archetypes(gid, [ARCH])
    for _ in range(3):
        ctx.run_system(gid, sync=False)
    ctx.synchronize()
    ctx.timer_start()
    for _ in range(passes):
        ctx.run_system(gid, sync=False)
    generic_ms = ctx.timer_stop() / passes
    ggbs = 36.0 * n / (generic_ms * 1e-3) / 1e9
    out["simple_iter_10M_generic_system"] = {
        "GBps": ggbs,
        "frac_of_measured_hbm": ggbs / peak,
        "ms": float(generic_ms),
        "bytes_per_entity": 36,
        "how": "same 20-pass bracket; NVRTC-compiled body in the generated TMA-pipelined kernel",
    }
    # rain: 4 Mi drops, fused wind -> gravity -> movement, 48 B / entity / tick
    n = 4 * 1024 * 1024
    VEL, POS, ARCH2 = 201, 202, 8
    vel, mass, pos = scenes.rain_drops(n, 1)
    ctx.bind_column(ARCH2, VEL, vel)
    ctx.bind_column(ARCH2, POS, pos)
    ctx.upload(ARCH2, VEL)
    ctx.upload(ARCH2, POS)
    w = ctx.create_system(_lib.SYS_RAIN_WIND, [VEL])
    g = ctx.create_system(_lib.SYS_RAIN_GRAVITY, [VEL])
    m = ctx.create_system(_lib.SYS_RAIN_MOVEMENT, [POS, VEL])
    for s in (w, g, m):
        ctx.set_archetypes(s, [ARCH2])
    ctx.set_tick_order([w, g, m])
    for _ in range(3):
        ctx.tick(1)
    ms = []
    for _ in range(20):
        ctx.flush_l2()
        ctx.timer_start()
        ctx.tick(1)
        ms.append(ctx.timer_stop())
    gbs = 48.0 * n / (np.median(ms) * 1e-3) / 1e9
    ctx.synchronize()
    ctx.timer_start()
    for _ in range(20):
        ctx.tick(1)
    resident_ms = ctx.timer_stop() / 20
    out["rain_4Mi_fused"] = {
        "GBps": gbs,
        "frac_of_measured_hbm": gbs / peak,
        "ms": float(np.median(ms)),
        "bytes_per_entity": 48,
        "how": "single ticks, L2 flushed before each (write 256 MiB, then stream 256 MiB of loads so no dirty lines remain)",
        "back_to_back_ms": float(resident_ms),
        "back_to_back_note": "the two 50 MB columns fit the 126 MB L2, so consecutive ticks run L2-resident and even a flushed tick "
                             "keeps its writes in L2: see rain_16Mi_fused for the HBM-bound rate",
    }
    # the same tick HBM-bound: 16 Mi drops (2 x 200 MB columns, larger than L2), 20 ticks back to back
    n = 16 * 1024 * 1024
    vel, mass, pos = scenes.rain_drops(n, 1)
    ctx.bind_column(ARCH2, VEL, vel)
    ctx.bind_column(ARCH2, POS, pos)
    ctx.upload(ARCH2, VEL)
    ctx.upload(ARCH2, POS)
    for _ in range(3):
        ctx.tick(1)
    ctx.synchronize()
    ctx.timer_start()
    for _ in range(20):
        ctx.tick(1)
    big_ms = ctx.timer_stop() / 20
    bgbs = 48.0 * n / (big_ms * 1e-3) / 1e9
    out["rain_16Mi_fused"] = {
        "GBps": bgbs,
        "frac_of_measured_hbm": bgbs / peak,
        "ms": float(big_ms),
        "bytes_per_entity": 48,
        "how": "20 ticks back to back in one CUDA-event bracket; columns (400 MB) larger than L2: the HBM-bound rate",
    }
    out["hbm_peak_GBps"] = peak
    return out


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
