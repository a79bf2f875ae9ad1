"""Multi-rank parity check of the index-sharded n-body path (run under torchrun, ONE RANK PER GPU: the ranks' kernels
wait on one another, so they must never share a device).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node R --master-addr 127.0.0.1 --master-port P \
        tests/tools/multi_gpu_check.py [bodies] [ticks] [full|own] [p2p|nccl] [noref]

Checks, per rank, on the rows it owns: the oracle tolerances (positions bit-exact after one tick), and that the
sharded result equals the SINGLE-GPU result of the same library bit for bit (the summation tree does not depend on
the rank count, kernels.cuh).
"""
import os
import sys

import numpy as np

sys.path.insert(0, ".")
import torch.distributed as dist  # noqa: E402

import oracle  # noqa: E402  (checker)
from recs_b200 import GpuContext, scenes  # noqa: E402
from recs_b200.nbody import NBodyGpuApp  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
ticks = int(sys.argv[2]) if len(sys.argv) > 2 else 3
own_rows_only = len(sys.argv) > 3 and sys.argv[3] == "own"  # every rank's host holds (and moves) only its own rows
exchange = sys.argv[4] if len(sys.argv) > 4 else "p2p"
skip_oracle = len(sys.argv) > 5 and sys.argv[5] == "noref"  # sizes the CPU oracle takes minutes for: bit-identity with the single-GPU run only
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
device = local
dist.init_process_group("gloo")
ctx = GpuContext(device=device, rank=rank, world_size=world, use_cuda_graph=False)
if exchange == "nccl":
    uid = [ctx.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx.comm_init(uid[0])
else:
    handles = [None] * world
    dist.all_gather_object(handles, ctx.exchange_export(n))
    ctx.exchange_import(handles)
mode = ctx.exchange_mode()
assert mode == ("nccl-allgather" if exchange == "nccl" else "peer-push"), mode
rb, re = ctx.shard_rows(n)
pos, mass, vel, acc = scenes.spawn_bodies(scenes.all_heavy_random_cube_with_bodies(n), 5)
if n >= 64:  # coincident twins and sub-epsilon neighbours spread over the shards: which rows take a tile's exact result
    k = n // 16  # must not depend on the rank count either
    pos[k : 2 * k] = pos[:k]
    pos[2 * k : 3 * k] = pos[:k] + np.array([1e-4, 0, 0], np.float32)
    pos = np.ascontiguousarray(pos)
start = [a.copy() for a in (pos, mass, vel, acc)]
if own_rows_only:  # poison what this rank does not own: it must never be read
    for a in (pos, mass, vel, acc):
        a[:rb] = np.nan
        a[re:] = np.nan
app = NBodyGpuApp(ctx, pos, mass, vel, acc)
app.upload(rows=(rb, re) if own_rows_only else None)
app.tick(ticks)
app.download(rows=(rb, re) if own_rows_only else None)
ctx.synchronize()
dist.barrier()  # every rank is done with the exchange before anyone tears its buffers down

# the same ticks on ONE GPU with the same library: own rows must agree bit for bit
single = [a.copy() for a in start]
with GpuContext(device=device, use_cuda_graph=False) as one:
    app1 = NBodyGpuApp(one, *single)
    app1.upload()
    app1.tick(ticks)
    app1.download()
same_as_single_gpu = all(
    np.array_equal(a[rb:re].view(np.uint32), b[rb:re].view(np.uint32)) for a, b in ((pos, single[0]), (vel, single[2]), (acc, single[3])))

ref = [a.copy() for a in (single if skip_oracle else start)]
if not skip_oracle:
    orc = oracle.load_nbody()
    orc.run_sequential(ref[0], ref[1], ref[2], ref[3], ticks)


def rel(got, want, floor=0.0):
    num = np.linalg.norm(got.astype(np.float64) - want, axis=1)
    den = np.maximum(np.linalg.norm(want.astype(np.float64), axis=1), floor)
    return float((num / np.where(den == 0, 1, den)).max()) if len(got) else 0.0


errs = {
    "pos": rel(pos[rb:re], ref[0][rb:re], 1.0),
    "vel": rel(vel[rb:re], ref[2][rb:re], 1.0),
    "acc": rel(acc[rb:re], ref[3][rb:re]),
}
ok = errs["pos"] <= 1e-4 and errs["vel"] <= 1e-4 and errs["acc"] <= (1e-4 if ticks == 1 else 1e-3)
if ticks == 1:
    ok = ok and np.array_equal(pos[rb:re].view(np.uint32), ref[0][rb:re].view(np.uint32))
ok = ok and same_as_single_gpu
results = [None] * world
dist.all_gather_object(results, (rank, rb, re, errs, {"bit_identical_to_single_gpu": bool(same_as_single_gpu)}, bool(ok)))
if rank == 0:
    for r in results:
        print(r)
    print("MULTI_GPU_CHECK", "OK" if all(r[-1] for r in results) else "FAILED",
          f"bodies={n} ticks={ticks} world={world} own_rows_only={own_rows_only} exchange={mode}")
ctx.close()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
