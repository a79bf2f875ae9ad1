#!/usr/bin/env python
"""Turns the raw ncu output of tools/run_round2_profiles.sh (gpurun_out/r02_*) into the committed summaries under
profiles/: launch lists, per-kernel ncu counters, DRAM traffic per launch (profiles/r02_traffic.json, read by bench.py)
and the SASS instruction mix of the dominant kernels (from the built objects, no GPU needed)."""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
KERNEL = r"(nbody_gravity_kernel|nbody_gravity_reduce_kernel|pack_posm_kernel|fp32_peak_kernel|fill_kernel|read_sweep_kernel|recs_generic_system|collision_flags_kernel|\w+_kernel)"


def launch_list(n, path, extra=""):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    seq = []
    for r in rows[1:]:
        m = re.search(KERNEL, r[ki])
        seq.append((m.group(1) if m else r[ki][:40], float(r[vi].replace(",", "")) / 1e3))
    lines = [f"# ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv python bench.py {extra}--steps 2 --warmup 3 --no-extras --no-configs --no-cpu-baseline",
             f"# round 2, one B200, {n} bodies; per-launch device times are cold-cache and serialised (compare shares, not absolutes)",
             "# launch  kernel  us"]
    lines += [f"{i:3d}  {k:30s} {us:12.2f}" for i, (k, us) in enumerate(seq)]
    agg = collections.OrderedDict()
    for k, us in seq:
        agg.setdefault(k, []).append(us)
    mean = lambda k: sum(agg[k]) / len(agg[k]) if k in agg else 0.0
    g, r, cs = mean("nbody_gravity_kernel"), mean("nbody_gravity_reduce_kernel"), mean("nbody_gravity_chunksum_kernel")
    lines.append(f"# per tick: nbody_gravity_kernel {g:.2f} us + nbody_gravity_chunksum_kernel (chunk sums of a separately dealt last i-block) "
                 f"{cs:.2f} us + nbody_gravity_reduce_kernel (fragment sum + movement + acceleration + mirror refresh) {r:.2f} us "
                 f"-> interaction kernel share of the tick {g / (g + r + cs):.4f}")
    lines.append("# (recs_generic_system = the reference-order generic gravity of the parity block; fp32_peak_kernel = the live FFMA2 probe; "
                 "fill / read_sweep = L2 flush; pack_posm_kernel = mirror pack after an upload)")
    open(os.path.join(P, f"r02_launch_list_{n}.txt"), "w").write("\n".join(lines) + "\n")


def raw_page(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]


def to_bytes(value, unit):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
    return float(value.replace(",", "")) * scale


def main():
    traffic = {"_source": "ncu --set full --clock-control none (tools/run_round2_profiles.sh), per launch: dram__bytes_read.sum + dram__bytes_write.sum",
               "_files": {}}
    if os.path.exists(os.path.join(G, "r02_launches_1M.csv")):
        launch_list(1048576, os.path.join(G, "r02_launches_1M.csv"))
        launch_list(65536, os.path.join(G, "r02_launches_65536.csv"), "--bodies 65536 ")
    lines = ["# ncu --set full --clock-control none --import-source on -k regex:nbody_gravity_kernel -c 1 python tools/prof_gravity.py <bodies> <reps>",
             "# round 2, one B200; the default shape nbody_gravity_kernel<4, 1, 16, 1, 12, 1, 16> (i-packed, overflow detector, unroll 16)"]
    for n, rep in ((1048576, "r02_prof_gravity_1M.ncu-rep"), (65536, "r02_prof_gravity_65536.ncu-rep")):
        path = os.path.join(G, rep)
        if not os.path.exists(path):
            continue
        hdr, units, rows = raw_page(path)
        idx = {h: i for i, h in enumerate(hdr)}
        r = rows[0]
        lines.append(f"\n== {n} bodies: {r[idx['Kernel Name']]}")
        for w in WANT:
            if w in idx:
                lines.append(f"  {w:70s} {r[idx[w]]} {units[idx[w]]}")
        rd, wr = to_bytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]]), to_bytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
        traffic[f"nbody_gravity_kernel_{n}"] = int(rd + wr)
        traffic[f"nbody_gravity_kernel_{n}_read"] = int(rd)
        traffic[f"nbody_gravity_kernel_{n}_write"] = int(wr)
        traffic["_files"][f"nbody_gravity_kernel_{n}"] = "profiles/r02_gravity_ncu_summary.txt"
        alg = 16 * n + 12 * n
        lines.append(f"  algorithmic input bytes: mirror {16 * n} + outer positions {12 * n} = {alg}; DRAM read {int(rd)} ({rd / alg:.2f} x): no re-reads")
        lines.append(f"  DRAM written {int(wr)}: the chunk / tile fragments of the summation tree (results independent of grid, launch split and rank "
                     f"count); {wr / 6.5e12 * 1e3:.3f} ms of HBM time in a {float(r[idx['gpu__time_duration.sum']]):.1f} ms FP32-bound launch")
    open(os.path.join(P, "r02_gravity_ncu_summary.txt"), "w").write("\n".join(lines) + "\n")
    path = os.path.join(G, "r02_prof_streaming.ncu-rep")
    if os.path.exists(path):
        hdr, units, rows = raw_page(path)
        idx = {h: i for i, h in enumerate(hdr)}
        names = ["axpy_rn_vec_kernel_10M"] * 3 + ["recs_generic_system_simple_iter_10M"] * 3 + ["rain_tma_kernel_4Mi"] * 3 + ["rain_tma_kernel_16Mi"] * 3
        alg = {"axpy_rn_vec_kernel_10M": 36 * 10_000_000, "recs_generic_system_simple_iter_10M": 36 * 10_000_000,
               "rain_tma_kernel_4Mi": 48 * 4 * 1024 * 1024, "rain_tma_kernel_16Mi": 48 * 16 * 1024 * 1024}
        lines = ["# ncu --set full --clock-control none -k regex:axpy_rn_vec_kernel|rain_tma_kernel|recs_generic_system -c 12 python tools/prof_streaming_once.py",
                 "# round 2, one B200.  per launch: duration (cold, under ncu), DRAM bytes read / written INSIDE the launch, DRAM throughput %",
                 "# reads equal the algorithmic read bytes (no re-reads); writes inside the launch are below the algorithmic write bytes because up to",
                 "# ~60 MB of dirty lines are still in the 126 MB L2 when the launch ends (they are written back during the next one)"]
        acc = collections.OrderedDict()
        for name, r in zip(names, rows):
            rd = to_bytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]])
            wr = to_bytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
            acc.setdefault(name, []).append((rd, wr))
            lines.append(f"{name:40s} grid {r[idx['launch__grid_size']]:>6s}  {float(r[idx['gpu__time_duration.sum']]):8.2f} us  read {rd / 1e6:8.2f} MB  "
                         f"write {wr / 1e6:8.2f} MB  dram {float(r[idx['gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']]):5.1f} %  (algorithmic {alg[name] / 1e6:.1f} MB)")
        for name, v in acc.items():
            traffic[name] = int(sum(a + b for a, b in v) / len(v))
            traffic[name + "_read"] = int(sum(a for a, _ in v) / len(v))
            traffic[name + "_write"] = int(sum(b for _, b in v) / len(v))
            traffic[name + "_algorithmic"] = alg[name]
            traffic["_files"][name] = "profiles/r02_streaming_ncu_summary.txt"
        open(os.path.join(P, "r02_streaming_ncu_summary.txt"), "w").write("\n".join(lines) + "\n")
    if len(traffic) > 2:
        json.dump(traffic, open(os.path.join(P, "r02_traffic.json"), "w"), indent=1)
    # SASS instruction mix (from the built objects)
    sass = ["# cuobjdump -sass of the built objects (sm_100a), instruction mix per loop (tools/loops.py); round 2",
            "# UBLKCP = 1-D TMA bulk copy, SYNCS = mbarrier; FFMA2 / FMUL2 / FADD2 = packed FP32; no UTC*MMA / LDTM: not a contraction"]
    for obj, key, title in (("recs_b200/csrc/gravity.o", "nbody_gravity_kernelILi4ELi1ELi16ELb1ELi12ELb1ELi16E", "nbody_gravity_kernel<4,1,16,true,12,true,16> (default)"),
                            ("recs_b200/csrc/streaming.o", "rain_tma_kernelILi2E", "rain_tma_kernel<2> (fused wind -> gravity -> movement)"),
                            ("recs_b200/csrc/select.o", "collision_flags_kernel", "collision_flags_kernel")):
        listing = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, obj)], capture_output=True, text=True).stdout
        tmp = os.path.join(G, "_sass.tmp")
        open(tmp, "w").write(listing)
        res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "loops.py"), tmp, key], capture_output=True, text=True).stdout
        counts = collections.Counter()
        on = False
        for line in listing.splitlines():
            if "Function :" in line:
                on = key in line
                continue
            m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
            if on and m:
                counts[m.group(1)] += 1
        sass.append(f"\n== {title}")
        sass.append("whole kernel: " + ", ".join(f"{k} {v}" for k, v in counts.most_common(18)))
        sass.append("TMA / mbarrier / tensor: " + ", ".join(f"{k} {counts.get(k, 0)}" for k in ("UBLKCP", "SYNCS", "LDG", "STG", "LDS", "STS", "UTCHMMA", "LDTM")))
        sass += ["  " + l for l in res.splitlines()]
    open(os.path.join(P, "r02_sass_dominant_kernels.txt"), "w").write("\n".join(sass) + "\n")
    print("profiles written")


main()
