"""Multi-rank check of GENERIC nested-query systems on a sharded context (run under torchrun, one rank per GPU; needs
NCCL for recs_gpu_column_allgather).

The n-body tick with `gravity` as the generic reference-order system (recs_b200/nbody.py): every rank folds the whole
nested query for the outer rows it owns, movement / acceleration stream the own rows, and the Position column is
gathered before the next tick's nested query.  The reference-order fold is bit-exact, so after several ticks every
rank's own rows must equal the CPU oracle's sequential run BIT FOR BIT.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node R --master-addr 127.0.0.1 --master-port P \
        tests/tools/multi_gpu_pair_check.py [bodies] [ticks]
"""
import os
import sys

import numpy as np

sys.path.insert(0, ".")
import torch.distributed as dist  # noqa: E402

import oracle  # noqa: E402  (checker)
from recs_b200 import GpuContext, scenes  # noqa: E402
from recs_b200.nbody import BODY_ARCHETYPE, POSITION, ReferenceOrderNBodyApp  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 6000
ticks = int(sys.argv[2]) if len(sys.argv) > 2 else 3
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dist.init_process_group("gloo")
ctx = GpuContext(device=local, rank=rank, world_size=world, use_cuda_graph=False)
uid = [ctx.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
ctx.comm_init(uid[0])
rb, re = ctx.shard_rows(n)
pos, mass, vel, acc = scenes.spawn_bodies(scenes.all_heavy_random_cube_with_bodies(n), 9)
ref = [a.copy() for a in (pos, mass, vel, acc)]
app = ReferenceOrderNBodyApp(ctx, pos, mass, vel, acc)
app.upload()
for _ in range(ticks):
    app.tick(1)                                      # gravity (own rows x all bodies), movement, acceleration (own rows)
    ctx.column_allgather(BODY_ARCHETYPE, POSITION)   # the nested query of the next tick reads every body's position
app.download()
ctx.synchronize()
orc = oracle.load_nbody()
orc.run_sequential(ref[0], ref[1], ref[2], ref[3], ticks)
own_exact = all(np.array_equal(a[rb:re].view(np.uint32), b[rb:re].view(np.uint32)) for a, b in ((pos, ref[0]), (vel, ref[2]), (acc, ref[3])))
all_pos = np.array_equal(pos.view(np.uint32), ref[0].view(np.uint32))  # gathered: every row of Position is current
results = [None] * world
dist.all_gather_object(results, (rank, rb, re, bool(own_exact), bool(all_pos)))
if rank == 0:
    for r in results:
        print(r)
    print("MULTI_GPU_PAIR_CHECK", "OK" if all(r[3] and r[4] for r in results) else "FAILED", f"bodies={n} ticks={ticks} world={world}")
dist.barrier()
ctx.close()
dist.destroy_process_group()
sys.exit(0 if own_exact and all_pos else 1)
