"""Multi-rank check of the row-sharded streaming paths (no exchange step): rain tick, simple_iter and a generic
system over two archetypes.  Every rank must have updated exactly its own slice of every archetype, bit-exact
against the oracle / numpy, and left the other rows alone (run under torchrun, one rank per GPU).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node R --master-addr 127.0.0.1 --master-port P \
        tests/tools/multi_gpu_stream_check.py [entities]
"""
import os
import sys

import numpy as np

sys.path.insert(0, ".")
import torch.distributed as dist  # noqa: E402

import oracle  # noqa: E402  (checker)
from recs_b200 import GpuContext, _lib, scenes  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_003
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dist.init_process_group("gloo")
ctx = GpuContext(device=local, rank=rank, world_size=world, use_cuda_graph=False)
orc = oracle.load_nbody()
ok = True
report = {}


def own(count):
    return ctx.shard_rows(count)


# ---- rain: wind -> gravity -> movement fused, 3 ticks -------------------------------------------------------------
VEL, POS, ARCH = 201, 202, 8
vel, mass, pos = scenes.rain_drops(n, 1)
vel[:] = np.random.default_rng(3).standard_normal(vel.shape).astype(np.float32) * np.float32(4.0)
vel0, pos0 = vel.copy(), pos.copy()
ctx.bind_column(ARCH, VEL, vel)
ctx.bind_column(ARCH, POS, pos)
ctx.upload(ARCH, VEL)
ctx.upload(ARCH, POS)
w = ctx.create_system(_lib.SYS_RAIN_WIND, [VEL])
g = ctx.create_system(_lib.SYS_RAIN_GRAVITY, [VEL])
m = ctx.create_system(_lib.SYS_RAIN_MOVEMENT, [POS, VEL])
for s in (w, g, m):
    ctx.set_archetypes(s, [ARCH])
ctx.set_tick_order([w, g, m])
direction = np.array([1.0, 0.0, 0.0], np.float32)
ctx.tick(3)
ctx.download(ARCH, VEL)
ctx.download(ARCH, POS)
rv, rp = vel0.copy(), pos0.copy()
for _ in range(3):
    orc.rain_wind(rv, direction)
    orc.rain_gravity(rv)
    orc.rain_movement(rp, rv)
rb, re = own(n)
inside = np.array_equal(vel[rb:re].view(np.uint32), rv[rb:re].view(np.uint32)) and np.array_equal(
    pos[rb:re].view(np.uint32), rp[rb:re].view(np.uint32))
outside = np.array_equal(np.delete(vel, np.s_[rb:re], 0), np.delete(vel0, np.s_[rb:re], 0)) and np.array_equal(
    np.delete(pos, np.s_[rb:re], 0), np.delete(pos0, np.s_[rb:re], 0))
report["rain"] = (rb, re, inside, outside)
ok = ok and inside and outside

# ---- simple_iter (built-in) and a generic system over two archetypes ------------------------------------------------
A, B = 101, 102
cols = {}
for arch, count in ((11, n), (12, n // 3 + 7)):
    p = np.random.default_rng(arch).standard_normal((count, 3)).astype(np.float32)
    v = np.random.default_rng(arch + 100).standard_normal((count, 3)).astype(np.float32)
    cols[arch] = (p, v, p.copy())
    ctx.bind_column(arch, A, p)
    ctx.bind_column(arch, B, v)
    ctx.upload(arch, A)
    ctx.upload(arch, B)
hand = ctx.create_system(_lib.SYS_QUERY_ADD, [A, B])
src = ("struct P { float x, y, z; };\nstruct V { float x, y, z; };\n"
       "void twice(P& p, const V& v) { p.x += v.x * 2.0f; p.y += v.y * 2.0f; p.z += v.z * 2.0f; }\n")
gen = ctx.create_kernel_system("twice", src, "twice", [(A, 12, "write", "P"), (B, 12, "read", "V")])
for s in (hand, gen):
    ctx.set_archetypes(s, [11, 12])
ctx.run_system(hand)
ctx.run_system(gen)
for arch, (p, v, p0) in cols.items():
    ctx.download(arch, A)
    rb, re = own(len(p))
    want = p0.copy()
    want[rb:re] = (p0[rb:re] + v[rb:re]) + v[rb:re] * np.float32(2.0)
    good = np.array_equal(p.view(np.uint32), want.view(np.uint32))
    report[f"query_arch{arch}"] = (rb, re, good)
    ok = ok and good

results = [None] * world
dist.all_gather_object(results, (rank, report, bool(ok)))
if rank == 0:
    for r in results:
        print(r)
    covered = sorted((r[1]["rain"][0], r[1]["rain"][1]) for r in results)
    contiguous = covered[0][0] == 0 and covered[-1][1] == n and all(a[1] == b[0] for a, b in zip(covered, covered[1:]))
    print("MULTI_GPU_STREAM_CHECK", "OK" if all(r[2] for r in results) and contiguous else "FAILED", f"entities={n} world={world}")
ctx.close()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
