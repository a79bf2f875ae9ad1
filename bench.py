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
  e2e        the same tick through the C ABI with HOST (pinned) columns: upload of the tick's
             input columns (Position, Mass, Velocity), tick, download of Position / Velocity /
             Acceleration inside the timed region.
  roofline   the interaction kernel alone (CUDA events around its launches) against the FP32 FMA
             rate measured live on the same GPU with a register-resident FFMA loop.
  parity     4 096 sampled rows of the first tick against the oracle: counts of rows
             over 1e-4 vs the reference-order f32 sum and vs the f64 sum, and the arbitration of
             rows where the reference's own f32 rounding is the larger error.
  configs    (N = 1) sub-records for configs[0] (1 000 bodies x 100 ticks) and configs[1]
             (65 536 bodies), each with value, roofline, parity, clocks and its cpu_baseline.
  cpu_baseline / --impl reference
             oracle/nbody_oracle.c (C restatement of the reference's CPU systems driven with its
             worker-pool policy) on the box's host cores, on a bounded row sample of the same
             workload, built on the box with -march=native (the reference's .cargo/config.toml
             asks for target-cpu=native).  The Rust reference itself cannot be built here.

N > 1 (torchrun, one rank per GPU): the B bodies are index-sharded, each rank computes its rows;
refreshed positions are pushed into every rank's mirror over NVLink by the tick's integration tail
(peer push over CUDA IPC; `RECS_GPU_EXCHANGE=nccl` selects the ncclAllGather exchange), "scaling":
"strong".
"""
from __future__ import annotations

import argparse
import importlib.util
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
DEFAULT_BODIES = 1_048_576
SEED = 1
PARITY_ROWS = 4096  # "At 1 M: 4 096 sampled rows"
ACC_TOL = 1e-4


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--bodies", type=int, default=DEFAULT_BODIES)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the configs[0] / configs[1] sub-records")
    return ap.parse_args()


def load_scenes():
    """recs_b200/scenes.py as a stand-alone module: the reference arm must not map librecs_gpu.so, which importing
    the recs_b200 package does."""
    spec = importlib.util.spec_from_file_location("recs_scenes", os.path.join(ROOT, "recs_b200", "scenes.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["recs_scenes"] = mod
    spec.loader.exec_module(mod)
    return mod


def measured_traffic() -> dict:
    """dram bytes per launch of the dominant kernels, from the kept ncu captures (profiles/r02_traffic.json names the
    CSV each figure was read from)."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
    except Exception:
        return {}


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

    def __init__(self, device: int, period_ms: int = 100):
        self.device = device
        self.period_ms = period_ms
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms",
                 str(self.period_ms)],
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
def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def load_cpu_baseline_oracle():
    """The oracle built ON THIS BOX with -march=native (the reference's .cargo/config.toml asks for
    target-cpu=native), still -ffp-contract=off.  Falls back to the portable build when no compiler is present.
    Returns (NBodyOracle, note)."""
    import oracle
    from oracle import nbody as onb

    try:
        orc = onb.load_native()
        return orc, "built on this box: gcc -O2 -march=native -ffp-contract=off (oracle/liboracle.so rebuilt in place)"
    except Exception as exc:  # no gcc on the box: the portable build that travelled with the repo
        return oracle.load_nbody(), f"portable build (no -march=native: {exc!r})"


def _pool_rows(orc, pos, mass, vel, acc, rows, threads):
    """One tick (gravity -> movement -> acceleration) of the reference's executor policy on the
    first `rows` outer rows; the nested query still iterates all bodies."""
    import ctypes as C

    n = pos.shape[0]
    if rows >= n:
        orc.run_pool(pos, mass, vel, acc, 1, threads)
        return
    lib = orc.lib
    if not hasattr(lib, "_sample_ready"):
        z = C.c_size_t
        f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
        lib.oracle_nbody_run_pool_rows.argtypes = [f32p, f32p, f32p, f32p, z, z, C.c_int, C.c_int, C.c_int * 3]
        lib.oracle_nbody_run_pool_rows.restype = C.c_double
        lib._sample_ready = True
    lib.oracle_nbody_run_pool_rows(pos, mass, vel, acc, n, rows, 1, threads, (C.c_int * 3)(0, 1, 2))


def cpu_sample_rate(orc, scenes, n_bodies: int, budget_s: float, threads: int):
    """Times one tick of the oracle's worker pool on rows [0, S) x all n_bodies j (a bounded row
    sample of the workload).  Returns (interactions/s, S, seconds, state)."""
    pos, mass, vel, acc = scenes.spawn_bodies(scenes.all_heavy_random_cube_with_bodies(n_bodies), SEED)
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
    return rows * n_bodies / sec, rows, sec, (pos, mass, vel, acc)


def run_reference_arm(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    scenes = load_scenes()
    orc, build_note = load_cpu_baseline_oracle()
    threads = host_threads()
    steps = args.steps or 3
    per_step_budget = max(1.0, min(20.0, 90.0 / (steps + args.warmup)))
    rate, rows, sec, state = cpu_sample_rate(orc, scenes, args.bodies, per_step_budget, threads)
    pos, mass, vel, acc = state
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
            "note": "C restatement of recs' CPU systems + worker-pool segment policy (oracle/nbody_oracle.c); the Rust "
                    "reference cannot be built here; " + build_note,
        },
        "e2e": {"value": value, "unit": "G interactions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
def _rel(a, b):
    return np.linalg.norm(a - b, axis=1) / np.maximum(np.linalg.norm(b, axis=1), 1e-30)


def sampled_parity(app, host_state, rb, re, n_rows, device, check_exact=True):
    """Parity gate of the run: accelerations of `n_rows` sampled rows of this rank's shard after the
    first tick against the oracle, all j.  Returns per-row error arrays (gpu vs reference-order f32, gpu vs f64,
    reference-order f32 vs f64) and whether the generic reference-order kernel reproduces the oracle bit for bit."""
    import oracle

    orc = oracle.load_nbody()
    pos0, mass0 = host_state
    rows = np.sort(np.random.default_rng(11).choice(np.arange(rb, re), size=min(n_rows, re - rb), replace=False))
    ref32, ref64 = orc.gravity_sample(pos0, mass0, rows)
    got = app.acc[rows].astype(np.float64)
    # a subset of the rows through the generic reference-order system (recs_b200/nbody.py): sequential fold in query
    # order, IEEE division and square root -- must equal the oracle's f32 result bit for bit, at full problem size
    exact = None
    if check_exact:
        try:
            from recs_b200 import GpuContext
            from recs_b200.nbody import reference_order_gravity

            sub = rows[:: max(1, len(rows) // 128)]
            with GpuContext(device=device, use_cuda_graph=False) as check_ctx:
                got_exact = reference_order_gravity(check_ctx, np.ascontiguousarray(pos0[sub]), pos0, mass0)
            want = ref32[:: max(1, len(rows) // 128)]
            exact = bool(np.array_equal(got_exact.view(np.uint32), want.view(np.uint32)))
        except Exception as exc:  # the check never invalidates the headline line
            exact = repr(exc)
    return _rel(got, ref32.astype(np.float64)), _rel(got, ref64), _rel(ref32.astype(np.float64), ref64), exact, rows


def parity_record(err32, err64, ref_err64, exact, pos_exact, reduce_max, reduce_sum):
    """The parity block of a bench line.  Arbitration rule: a row over the literal 1e-4 tolerance against
    the reference-order f32 sum is attributed to the reference's own sequential-sum rounding when that sum is further
    from the f64 sum than the GPU result is AND the GPU result is within 1e-5 of the f64 sum; any other row over the
    tolerance is a failure of this implementation."""
    over32 = err32 > ACC_TOL
    explained = over32 & (ref_err64 > err64) & (err64 <= 1e-5)
    return {
        "rows": int(reduce_sum(len(err32))),
        "tolerance": ACC_TOL,
        "acc_max_rel_err_vs_oracle": reduce_max(float(err32.max())),
        "acc_max_rel_err_vs_f64": reduce_max(float(err64.max())),
        "oracle_f32_order_vs_f64_max": reduce_max(float(ref_err64.max())),
        "rows_over_tol_vs_ref32": int(reduce_sum(int(over32.sum()))),
        "rows_over_tol_vs_f64": int(reduce_sum(int((err64 > ACC_TOL).sum()))),
        "rows_over_tol_where_reference_f32_is_the_inaccurate_side": int(reduce_sum(int(explained.sum()))),
        "rows_failing_after_arbitration": int(reduce_sum(int((over32 & ~explained).sum()))),
        "arbitration": "row passes if <= 1e-4 vs reference-order f32, or if the reference-order f32 sum is further from the f64 "
                       "sum than the GPU result and the GPU result is <= 1e-5 from the f64 sum",
        "pos_bit_exact_after_tick1": pos_exact,
        "reference_order_generic_kernel_bit_exact": exact,
    }


def kernel_roofline(ctx, app, n, rows_here, k_steps, traffic):
    tf_peak, mhz = ctx.measure_fp32_peak(packed=True)
    ctx.enable_timing(True)
    ctx.reset_timing()
    for _ in range(k_steps):
        app.tick(1)
    t = ctx.timing()
    ctx.enable_timing(False)
    grav_ms = t["gravity_ms"] / k_steps  # all interaction-kernel launches of one tick on this rank
    kernel_flops = FLOPS_PER_INTERACTION * rows_here * n
    achieved_tf = kernel_flops / (grav_ms * 1e-3) / 1e12
    fma_lane_rate = FMA_LANE_OPS_PER_INTERACTION * rows_here * n / (grav_ms * 1e-3)
    return {
        "bound": "fp32",
        "kernel": "nbody_gravity_kernel",
        "achieved": achieved_tf,
        "peak": tf_peak,
        "unit": "TFLOP/s",
        "frac": achieved_tf / tf_peak,
        "peak_source": "measured live on this GPU: register-resident FFMA2 loop (MEASURED_PEAKS.json has no FP32 figure)",
        "peak_derived": 148 * 128 * 2 * mhz * 1e6 / 1e12,
        "peak_derived_source": "148 SMs x 128 lanes x 2 FLOP x SM clock seen by the probe (SURVEY 8d formula)",
        "flops_per_interaction": FLOPS_PER_INTERACTION,
        "fma_pipe_frac": fma_lane_rate / (tf_peak * 1e12 / 2.0),
        "kernel_ms_per_step": grav_ms,
        "kernel_launches_per_step": t["gravity_launches"] / k_steps,
        "kernel_share_of_step": grav_ms / (t["last_tick_chain_ms"] / k_steps) if t["last_tick_chain_ms"] else None,
        "traffic": traffic,
        "traffic_unit": "bytes of DRAM traffic per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum, single GPU, "
                        "profiles/r02_traffic.json)",
        "sm_mhz_during_peak_probe": mhz,
    }, t


def energy_drift(orc, initial, finals: dict) -> dict:
    """|E_end - E_0| / |E_0| with E = kinetic + potential in f64 (oracle.total_energy) for each named final state."""
    k0, u0 = orc.total_energy(initial[0], initial[1], initial[2])
    out = {"e0": k0 + u0, "kinetic0": k0, "potential0": u0, "norm": "|E_end - E_0| / |E_0|, E in f64 from the columns"}
    for name, (p, v) in finals.items():
        k, u = orc.total_energy(p, initial[1], v)
        out[name] = abs((k + u) - (k0 + u0)) / abs(k0 + u0)
    return out


def config_subrecord(n, ticks_per_step, steps, device, traffic, scenes, cpu_orc, cpu_note):
    """One secondary config on a fresh single-GPU context: value, roofline, parity, clocks, cpu_baseline."""
    import oracle
    from recs_b200 import GpuContext
    from recs_b200.nbody import NBodyGpuApp

    scene = scenes.EVERYTHING_HEAVY if n == 1000 else scenes.all_heavy_random_cube_with_bodies(n)
    pos0, mass0, vel0, acc0 = scenes.spawn_bodies(scene, SEED)
    orc = oracle.load_nbody()
    with GpuContext(device=device, use_cuda_graph=True) as ctx:
        cols = [ctx.pinned_empty(a.shape) for a in (pos0, mass0, vel0, acc0)]
        for dst, src in zip(cols, (pos0, mass0, vel0, acc0)):
            dst[...] = src
        app = NBodyGpuApp(ctx, *cols)
        app.upload()
        app.tick(1)
        app.download()
        pref = pos0.copy()
        orc.axpy(pref, vel0, orc.dt)
        pos_exact = bool(np.array_equal(app.pos.view(np.uint32), pref.view(np.uint32)))
        e32, e64, r64, exact, _ = sampled_parity(app, (pos0, mass0), 0, n, min(n, 1024), device, check_exact=False)
        ident = lambda x: x
        parity = parity_record(e32, e64, r64, exact, pos_exact, ident, ident)
        if ticks_per_step > 1:
            # the whole config: ticks_per_step ticks from the initial state against the oracle (drift)
            for dst, src in zip(cols, (pos0, mass0, vel0, acc0)):
                dst[...] = src
            app.upload()
            app.tick(ticks_per_step)
            app.download()
            ref = [a.copy() for a in (pos0, mass0, vel0, acc0)]
            orc.run_sequential(ref[0], ref[1], ref[2], ref[3], ticks_per_step)
            den = lambda r: np.maximum(np.linalg.norm(r, axis=1), 1.0)
            parity[f"drift_after_{ticks_per_step}_ticks"] = {
                "pos_max": float((np.linalg.norm(app.pos.astype(np.float64) - ref[0], axis=1) / den(ref[0])).max()),
                "vel_max": float((np.linalg.norm(app.vel.astype(np.float64) - ref[2], axis=1) / den(ref[2])).max()),
                "bound": 1e-4,
                "norm": "max_i |x_gpu - x_ref| / max(|x_ref|, 1)",
            }
            # relative total-energy drift, reported beside the f64 run of the same integration
            p64, v64, a64 = pos0.astype(np.float64), vel0.astype(np.float64), acc0.astype(np.float64)
            orc.run_f64(p64, mass0, v64, a64, ticks_per_step)
            parity[f"energy_drift_after_{ticks_per_step}_ticks"] = energy_drift(
                orc, (pos0, mass0, vel0), {"gpu": (app.pos, app.vel), "reference_order_f32": (ref[0], ref[2]), "f64_oracle": (p64, v64)})
        elif n <= 65536:
            # configs[1]: the north star's 100 steps; the energy of the GPU run alone (the f64 run of 100 ticks at this
            # size is minutes of host time; tests/test_gpu_nbody.py compares the columns with the reference order)
            for dst, src in zip(cols, (pos0, mass0, vel0, acc0)):
                dst[...] = src
            app.upload()
            app.tick(100)
            app.download()
            parity["energy_drift_after_100_ticks"] = energy_drift(orc, (pos0, mass0, vel0), {"gpu": (app.pos, app.vel)})
            parity["energy_drift_after_100_ticks"]["note"] = (
                "the integrator's own drift (semi-implicit Euler at dt = 1/20000 through close encounters), not arithmetic: "
                "tests/test_gpu_nbody.py::test_drift_65536_bodies_100_ticks runs the reference summation order beside it and "
                "asserts the two energies agree to 1e-6 of E_0")
        for _ in range(3):
            app.tick(ticks_per_step)
        ctx.synchronize()
        sampler = ClockSampler(device, period_ms=20)
        sampler.start()
        ms = []
        t_end = time.perf_counter() + 0.6  # at least this long under load, so the clock sampler sees it
        while len(ms) < steps or time.perf_counter() < t_end:
            ctx.flush_l2()
            ctx.timer_start()
            app.tick(ticks_per_step)
            ms.append(ctx.timer_stop())
            if len(ms) >= 20 * steps:
                break
        clocks = sampler.stop()
        ms_step = float(np.mean(ms))
        value = n * n * ticks_per_step / (ms_step * 1e-3) / 1e9
        roof, _ = kernel_roofline(ctx, app, n, n, 10, traffic)
        if ticks_per_step > 1:
            roof["note"] = ("latency-bound at this size: ticks 2..n of a recs_gpu_tick(n) call run in ONE cooperative launch "
                            "(nbody_resident_ticks_kernel, one grid barrier per tick; bit-identical to the chain); this block times "
                            "the interaction kernel of the tick-by-tick chain (two launches per tick), which is what timing mode runs")
        # end to end at this size: upload inputs, tick(s), download results
        e0 = time.perf_counter()
        e_steps = max(3, min(steps, 20))
        for _ in range(e_steps):
            app.upload(comps=app.INPUT_COMPONENTS)
            app.tick(ticks_per_step)
            app.download()
        e2e_s = (time.perf_counter() - e0) / e_steps
    threads = host_threads()
    cpu = None
    if cpu_orc is not None:
        if n * n * ticks_per_step <= 2e9:  # measured directly: the whole config on the host cores
            ref = [a.copy() for a in (pos0, mass0, vel0, acc0)]
            t0 = time.perf_counter()
            cpu_orc.run_pool(ref[0], ref[1], ref[2], ref[3], ticks_per_step, threads)
            sec = time.perf_counter() - t0
            cpu = {"value": n * n * ticks_per_step / sec / 1e9, "unit": "G interactions/s", "cores": threads, "kind": "port",
                   "sample": f"the whole config: {n} bodies x {ticks_per_step} ticks, {sec:.2f} s", "note": cpu_note}
        else:
            rate, rows, sec, _ = cpu_sample_rate(cpu_orc, scenes, n, 6.0, threads)
            cpu = {"value": rate / 1e9, "unit": "G interactions/s", "cores": threads, "kind": "port",
                   "sample": f"{rows} of {n} outer rows x all {n} bodies, one tick, {sec:.1f} s", "note": cpu_note}
    return {
        "workload": f"nbody_{n}" + (f"_x{ticks_per_step}_ticks" if ticks_per_step > 1 else ""),
        "value": value,
        "unit": "G interactions/s",
        "ms_per_step": ms_step,
        "ticks_per_step": ticks_per_step,
        "steps": len(ms),
        "e2e": {"value": n * n * ticks_per_step / e2e_s / 1e9, "unit": "G interactions/s", "ms_per_step": e2e_s * 1e3,
                "h2d_bytes_per_step": int(28 * n), "d2h_bytes_per_step": int(36 * n)},
        "roofline": roof,
        "parity": parity,
        "clocks": clocks,
        "cpu_baseline": cpu,
    }


def sharded_small_check(rank, world, local_rank, dist, scenes):
    """Every SCALE line carries it: 16 384 bodies, 2 ticks, sharded over the same ranks with the same exchange; every rank
    checks ALL the rows it owns against the oracle and against the single-GPU result of the same library (bit for bit)."""
    import oracle
    from recs_b200 import GpuContext
    from recs_b200.nbody import NBodyGpuApp

    n, ticks = 16384, 2
    start = scenes.spawn_bodies(scenes.all_heavy_random_cube_with_bodies(n), 5)
    cols = [a.copy() for a in start]
    ctx = GpuContext(device=local_rank, rank=rank, world_size=world, use_cuda_graph=False)
    mode = setup_exchange(ctx, dist, rank, world, n)
    rb, re = ctx.shard_rows(n)
    app = NBodyGpuApp(ctx, *cols)
    app.upload(rows=(rb, re))
    app.tick(ticks)
    app.download(rows=(rb, re))
    ctx.synchronize()
    dist.barrier()
    ctx.close()
    single = [a.copy() for a in start]
    with GpuContext(device=local_rank, use_cuda_graph=False) as one:
        app1 = NBodyGpuApp(one, *single)
        app1.upload()
        app1.tick(ticks)
        app1.download()
    same = all(np.array_equal(a[rb:re].view(np.uint32), b[rb:re].view(np.uint32))
               for a, b in ((cols[0], single[0]), (cols[2], single[2]), (cols[3], single[3])))
    ref = [a.copy() for a in start]
    orc = oracle.load_nbody()
    orc.run_pool(ref[0], ref[1], ref[2], ref[3], ticks, host_threads() // max(world, 1) or 1)
    den = lambda r, f: np.maximum(np.linalg.norm(r[rb:re].astype(np.float64), axis=1), f)
    err = lambda g, r, f: float((np.linalg.norm(g[rb:re].astype(np.float64) - r[rb:re], axis=1) / den(r, f)).max()) if re > rb else 0.0
    mine = {"pos": err(cols[0], ref[0], 1.0), "vel": err(cols[2], ref[2], 1.0), "acc": err(cols[3], ref[3], 1e-30), "same": bool(same)}
    everyone = [None] * world
    dist.all_gather_object(everyone, mine)
    return {
        "bodies": n, "ticks": ticks, "rows_checked": "every row of every rank's shard", "exchange": mode,
        "pos_max_rel_err": max(e["pos"] for e in everyone), "vel_max_rel_err": max(e["vel"] for e in everyone),
        "acc_max_rel_err": max(e["acc"] for e in everyone),
        "bit_identical_to_single_gpu_run": all(e["same"] for e in everyone),
        "ok": all(e["same"] and e["pos"] <= 1e-4 and e["vel"] <= 1e-4 and e["acc"] <= 1e-3 for e in everyone),
    }


def setup_exchange(ctx, dist, rank, world, n):
    """Peer push (CUDA IPC handles exchanged through torch.distributed) unless RECS_GPU_EXCHANGE=nccl."""
    if os.environ.get("RECS_GPU_EXCHANGE") != "nccl":
        try:
            handles = [None] * world
            dist.all_gather_object(handles, ctx.exchange_export(n))
            ctx.exchange_import(handles)
            return ctx.exchange_mode()
        except Exception as exc:  # IPC mapping refused on this box: the NCCL exchange is the other GPU path
            ctx.clear_error()
            sys.stderr.write(f"[bench] peer-push exchange unavailable ({exc!r}); using the NCCL all-gather exchange\n")
    uid = [ctx.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx.comm_init(uid[0])
    return ctx.exchange_mode()


def run_gpu_arm(args) -> None:
    from recs_b200 import GpuContext, scenes
    from recs_b200.nbody import NBodyGpuApp

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n = args.bodies
    steps = args.steps or (10 if n >= 500_000 else 50)
    warmup = max(args.warmup, 3)
    traffic = measured_traffic()

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod

        dist = dist_mod
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    small = sharded_small_check(rank, world, local_rank, dist, scenes) if world > 1 else None

    ctx = GpuContext(device=local_rank, rank=rank, world_size=world, use_cuda_graph=True)
    exchange = setup_exchange(ctx, dist, rank, world, n) if world > 1 else "single"
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

    def reduce_over_ranks(x: float, op: str) -> float:
        if dist is None:
            return x
        import torch

        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
        return float(t.item())

    max_over_ranks = lambda x: reduce_over_ranks(x, "max")
    sum_over_ranks = lambda x: reduce_over_ranks(float(x), "sum")

    # ---- parity gate: first tick against the oracle on sampled rows -----------------------------
    app.upload()
    app.tick(1)
    app.download()
    e32, e64, r64, exact, _ = sampled_parity(app, (pos0, mass0), rb, re, max(PARITY_ROWS // world, 1), local_rank)
    # positions after tick 1 must be bit-exact (movement precedes acceleration)
    import oracle

    orc = oracle.load_nbody()
    pref = pos0.copy()
    orc.axpy(pref, vel0, orc.dt, rows=(rb, re))
    pos_exact = bool(np.array_equal(app.pos[rb:re].view(np.uint32), pref[rb:re].view(np.uint32)))
    pos_exact = bool(max_over_ranks(0.0 if pos_exact else 1.0) == 0.0)
    if isinstance(exact, bool):
        exact = bool(max_over_ranks(0.0 if exact else 1.0) == 0.0)
    parity = parity_record(e32, e64, r64, exact, pos_exact, max_over_ranks, sum_over_ranks)
    parity["rows_per_rank"] = int(len(e32))
    if small is not None:
        parity["sharded_small_check"] = small

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
    # exchange, and it never touches their velocities or accelerations.  Inputs of a tick are Position, Mass and
    # Velocity; Acceleration is overwritten by gravity (crates/n-body/src/lib.rs:134) and only comes back.
    e2e_steps = steps
    own = None if world == 1 else (rb, re)
    for _ in range(2):
        app.upload(rows=own, comps=app.INPUT_COMPONENTS)
        app.tick(1)
        app.download(rows=own)
    barrier()
    e0 = time.perf_counter()
    for _ in range(e2e_steps):
        app.upload(rows=own, comps=app.INPUT_COMPONENTS)
        app.tick(1)
        app.download(rows=own)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - e0)
    e2e_value = n * n * e2e_steps / e2e_s / 1e9
    # whole-job bytes per step, counted from what is copied: 28 B up and 36 B down per body
    h2d = int(pos.nbytes + mass.nbytes + vel.nbytes)
    d2h = int(pos.nbytes + vel.nbytes + acc.nbytes)

    # ---- roofline of the interaction kernel (events around its launches only) --------------------
    k_steps = max(3, min(steps, 10))
    roofline, t = kernel_roofline(ctx, app, n, re - rb, k_steps, traffic.get(f"nbody_gravity_kernel_{n}") if world == 1 else None)

    comm = None
    if world > 1:
        kernel_by_rank = [None] * world
        dist.all_gather_object(kernel_by_rank, float(roofline["kernel_ms_per_step"]))
        tail_by_rank = [None] * world
        dist.all_gather_object(tail_by_rank, float(t["gravity_reduce_ms"]) / k_steps)
        comm = {
            "kernel_ms_per_step_by_rank": [round(x, 3) for x in kernel_by_rank],
            # chunk sums + reduce + integration tail (with the peer push), CUDA events around those launches
            "tail_ms_per_step_by_rank": [round(x, 3) for x in tail_by_rank],
            "exchange": exchange,
            "bytes_per_rank_per_step": int((re - rb) * 16 * (world if exchange == "peer-push" else 1)),
            "bytes_total_per_step": int(n * 16 * (world if exchange == "peer-push" else 1)),
            "interaction_launches_per_tick": t["gravity_launches"] / k_steps,
        }
        if exchange == "peer-push":
            comm["how"] = ("the integration tail stores every refreshed {position, G*m} entry into all ranks' next mirror buffer "
                           "(P2P stores over NVLink) and raises a generation flag; the next tick is ONE persistent launch that starts on "
                           "its own tiles and acquires a peer's flag before the first tile of that peer's slice")
            comm["residual"] = ("ms_per_step (slowest rank) minus the mean interaction-kernel time: the reduce + integration tail with its "
                                "P2P stores, and the skew between ranks (a rank's producer warp waits inside the kernel for the slowest "
                                "peer's flag, which the kernel time of the waiting rank already contains)")
            comm["residual_ms_per_step"] = total_ms / steps - float(np.mean(kernel_by_rank))
        else:
            comm["how"] = "ncclAllGather of the mirror, in place, between the own-tile launch and the launch over the other ranks' tiles"
            comm["gather_ms_per_step_incl_peer_wait"] = t["gather_ms"] / k_steps

    # ---- secondary configs: configs[0], configs[1] (N = 1) and the streaming systems ---------------------
    extras = {}
    configs = None
    cpu = None
    if world > 1:
        extras = {"skipped": "the streaming configs are reported by the N=1 run (a sharded context streams only its own slice)"}
    cpu_orc, cpu_note = (None, None)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:  # the contract: rank 0 at N=1 only
        cpu_orc, cpu_note = load_cpu_baseline_oracle()
    if world == 1 and rank == 0 and not args.no_extras:
        try:
            extras = streaming_extras(ctx, traffic)
        except Exception as exc:  # extras never invalidate the headline line
            extras = {"error": repr(exc)}
    ctx.close()
    if world == 1 and rank == 0 and not args.no_configs and n == DEFAULT_BODIES:
        configs = {}
        for key, (cn, tps, csteps) in {"configs[0]": (1000, 100, 20), "configs[1]": (65536, 1, 50)}.items():
            try:
                configs[key] = config_subrecord(cn, tps, csteps, local_rank, traffic.get(f"nbody_gravity_kernel_{cn}"), scenes,
                                                cpu_orc, cpu_note)
            except Exception as exc:
                configs[key] = {"error": repr(exc)}

    # ---- CPU baseline (rank 0, bounded sample) ---------------------------------------------------------
    if cpu_orc is not None:
        threads = host_threads()
        rate, rows, sec, _ = cpu_sample_rate(cpu_orc, scenes, n, 12.0, threads)
        cpu = {
            "value": rate / 1e9,
            "unit": "G interactions/s",
            "cores": threads,
            "kind": "port",
            "sample": f"{rows} of {n} outer rows x all {n} bodies, one tick, {sec:.1f} s",
            "note": "oracle/nbody_oracle.c: C restatement of recs' CPU systems + worker-pool segment policy; " + cpu_note,
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
                "parallelism": f"index-sharded x{world}, {exchange} exchange of positions per tick" if world > 1 else "single GPU",
                "l2": "flushed between timed ticks (256 MiB write + 256 MiB read sweep outside the event bracket)",
                "seed": SEED,
            },
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "G interactions/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_s / e2e_steps * 1e3},
            "gpu_launches": int(round(launches_per_step * steps)),
            "roofline": roofline,
            "cpu_baseline": cpu,
            "parity": parity,
            "step_ms": [round(x, 3) for x in step_ms],
            "ms_per_step_by_rank": [round(x, 3) for x in per_rank_ms],
            "wall_ms_per_step_incl_flush": wall / steps * 1e3,
            "comm": comm,
            "configs": configs,
            "extras": extras,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def hbm_roofline(kernel, gbs, peak, ms, traffic):
    return {"bound": "hbm", "kernel": kernel, "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
            "ms_per_launch": ms, "traffic": traffic,
            "traffic_unit": "bytes of DRAM traffic per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum, profiles/r02_traffic.json)"}


def rain_parity(ctx, arch, vel_id, pos_id, vel, pos, ticks, rows=1 << 20) -> dict:
    """The fused rain tick against the oracle's three systems in schedule order (wind -> gravity -> movement) on the first
    and last `rows` drops (drops are independent of each other), after the `ticks` ticks this section ran: bit for bit."""
    import oracle

    orc = oracle.load_nbody()
    n = vel.shape[0]
    v0, p0 = vel.copy(), pos.copy()  # bound host arrays still hold the initial state: nothing was downloaded yet
    vel = ctx.download(arch, vel_id)
    pos = ctx.download(arch, pos_id)
    direction = np.ascontiguousarray(ctx.get_wind_direction(), dtype=np.float32)
    ok, checked = True, 0
    for b, e in ((0, min(rows, n)), (max(n - rows, 0), n)):
        rv, rp = np.ascontiguousarray(v0[b:e]), np.ascontiguousarray(p0[b:e])
        for _ in range(ticks):
            orc.rain_wind(rv, direction)
            orc.rain_gravity(rv)
            orc.rain_movement(rp, rv)
        ok = ok and np.array_equal(rv.view(np.uint32), vel[b:e].view(np.uint32)) and np.array_equal(rp.view(np.uint32), pos[b:e].view(np.uint32))
        checked += e - b
    return {"bit_exact": bool(ok), "rows_checked": int(checked), "ticks": int(ticks),
            "how": "velocity and position columns against the oracle's wind, gravity, movement run in schedule order"}


def streaming_extras(ctx, traffic) -> dict:
    """HBM-bound systems of configs[3] / configs[4]: achieved GB/s on algorithmic bytes, each with its own roofline block."""
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
    sampler = ClockSampler(ctx.config.device, period_ms=20)
    sampler.start()
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
        "roofline": hbm_roofline("axpy_rn_vec_kernel", gbs, peak, float(steady_ms), traffic.get("axpy_rn_vec_kernel_10M")),
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
        "roofline": hbm_roofline("recs_generic_system (generated)", ggbs, peak, float(generic_ms), traffic.get("recs_generic_system_simple_iter_10M")),
    }
    # parity of both simple_iter systems: every pass adds (1, 1, 1) to a position that started at (1, 1, 1), exactly
    got = ctx.download(ARCH, A)
    n_passes = (3 + passes + 20) + (3 + passes)
    ok = bool(np.all(got == np.float32(1 + n_passes)))
    for key in ("simple_iter_10M", "simple_iter_10M_generic_system"):
        out[key]["parity"] = {"bit_exact": ok, "rows_checked": n, "how": f"all rows after the {n_passes} passes of this section (both systems), against 1 + passes"}
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
        "roofline": dict(hbm_roofline("rain_tma_kernel<2>", gbs, peak, float(np.median(ms)), traffic.get("rain_tma_kernel_4Mi")),
                         note="NOT an HBM figure: at this size the written half of the columns stays in L2 (traffic < algorithmic bytes)"),
    }
    out["rain_4Mi_fused"]["parity"] = rain_parity(ctx, ARCH2, VEL, POS, vel, pos, 3 + 20 + 20)
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
        "roofline": hbm_roofline("rain_tma_kernel<2>", bgbs, peak, float(big_ms), traffic.get("rain_tma_kernel_16Mi")),
    }
    out["clocks"] = sampler.stop()
    out["rain_16Mi_fused"]["parity"] = rain_parity(ctx, ARCH2, VEL, POS, vel, pos, 3 + 20)
    # the reference's rain_simulation benchmark as it registers it -- all seven systems, drops spawned (Commands::create) and
    # removed (ground_collision) every tick -- with the whole tick on the device and the host World playing the commands back
    from recs_b200.rain import rain_application

    app = rain_application(ctx, 16384)
    app.run(100)  # steady state: drops are landing
    ctx.synchronize()
    t0 = time.perf_counter()
    created = removed = 0
    for _ in range(50):
        app.run(1)
        c, r = app.last_playback()
        created, removed = created + c, removed + r
    ctx.synchronize()
    sec = time.perf_counter() - t0
    out["rain_benchmark_full_tick_16384_clouds"] = {
        "ms_per_tick": sec / 50 * 1e3,
        "drops_alive": len(app.world.archetype_entities(2)),
        "created_per_tick": created / 50,
        "removed_per_tick": removed / 50,
        "how": "wall clock per tick (host + device) over 50 ticks after 100 warm-up ticks; crates/ecs_bench_suite/src/recs/rain_simulation.rs "
               "system set, create_evenly_interspersed_clouds(16384); entity ids and columns of this set-up are checked against the oracle "
               "tick by tick in tests/test_gpu_playback.py::test_full_rain_tick_through_application",
    }
    app.close()
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
