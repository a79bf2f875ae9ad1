"""Host-side logic of the index-sharded n-body path, on CPU: the partition (from the C ABI) and the
per-tick exchange protocol with world_size = 2 over gloo.  The per-rank compute is stood in by the
CPU oracle here (this is a test of the sharding protocol, not of the kernels); the GPU/NCCL version of
the same protocol is exercised by bench.py --gpus N and tests/test_gpu_multi.py on the GPU box."""
import os
import socket

import numpy as np
import pytest

from recs_b200 import scenes
from recs_b200.sharding import shard_plan


@pytest.mark.parametrize("n,world", [(1, 1), (1000, 1), (1000, 2), (1000, 8), (65536, 4), (1048576, 8), (257, 2), (513, 4), (300, 8)])
def test_partition_is_tile_aligned_and_covers_all_rows(n, world):
    plans = [shard_plan(n, r, world) for r in range(world)]
    s = plans[0].slice_rows
    assert s % 256 == 0 and s * world >= n and all(p.slice_rows == s for p in plans)
    assert plans[0].row_begin == 0 and plans[-1].row_end == n
    for p, q in zip(plans, plans[1:]):
        assert p.row_end == q.row_begin
    for p in plans:
        assert p.row_begin == min(p.rank * s, n) and p.row_end == min((p.rank + 1) * s, n)
    if n == 1048576 and world == 8:
        assert s == 131072 and plans[3].gather_bytes_per_rank == 2 * 1024 * 1024  # 2 MiB / rank


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, ticks, out_dir):
    import torch
    import torch.distributed as dist

    import oracle

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    orc = oracle.load_nbody()
    plan = shard_plan(n, rank, world)
    pos, mass, vel, acc = scenes.spawn_bodies(scenes.all_heavy_random_cube_with_bodies(n), 3)
    rb, re = plan.row_begin, plan.row_end
    for _ in range(ticks):
        # gravity on own rows against ALL bodies (positions current on every rank)
        acc_new = orc.gravity(pos, mass, rows=(rb, re))
        acc[rb:re] = acc_new[rb:re]
        orc.axpy(pos, vel, orc.dt, rows=(rb, re))  # movement, own rows
        orc.axpy(vel, acc, orc.dt, rows=(rb, re))  # acceleration, own rows
        # all-gather of equally sized, padded slices (what ncclAllGather moves on the GPU)
        send = torch.zeros(plan.slice_rows, 3)
        send[: re - rb] = torch.from_numpy(pos[rb:re])
        recv = [torch.zeros(plan.slice_rows, 3) for _ in range(world)]
        dist.all_gather(recv, send)
        gathered = torch.cat(recv)[:n].numpy()
        pos[...] = gathered
    np.save(os.path.join(out_dir, f"pos_{rank}.npy"), pos)
    np.save(os.path.join(out_dir, f"vel_{rank}.npy"), vel[rb:re])
    dist.destroy_process_group()


def test_two_rank_protocol_matches_single_process_oracle(tmp_path, nbody_oracle):
    import torch.multiprocessing as mp

    n, ticks, world = 700, 3, 2
    mp.spawn(_worker, args=(world, _free_port(), n, ticks, str(tmp_path)), nprocs=world, join=True)
    pos, mass, vel, acc = scenes.spawn_bodies(scenes.all_heavy_random_cube_with_bodies(n), 3)
    nbody_oracle.run_sequential(pos, mass, vel, acc, ticks)
    p0, p1 = np.load(tmp_path / "pos_0.npy"), np.load(tmp_path / "pos_1.npy")
    # every rank ends with the full, identical position set, bit-equal to the unsharded run
    assert np.array_equal(p0.view(np.uint32), p1.view(np.uint32))
    assert np.array_equal(p0.view(np.uint32), pos.view(np.uint32))
    v = np.concatenate([np.load(tmp_path / "vel_0.npy"), np.load(tmp_path / "vel_1.npy")])
    assert np.array_equal(v.view(np.uint32), vel.view(np.uint32))


# ---- the peer-push exchange as a protocol model (include/recs_gpu.h, "peer push"; recs_b200/csrc/gravity.cu push_begin /
# push_end, the producer's wait_generation) ---------------------------------------------------------------------------
# Three processes share, per rank, what the GPUs share through CUDA IPC: two mirror buffers (generation g lives in
# buffer g & 1), one flag word per peer.  Each rank runs the tick the kernels run: read own slice at once, every other
# slice only after that peer's flag reached the generation; then (tail) wait until every rank has pushed generation g
# before writing generation g + 1 into the buffer that held g - 1, and push.  Ranks are skewed with random sleeps.  Every slice carries the generation it was written for, so a read
# of a slice that was overwritten early or not yet written shows up as a wrong stamp.
def _push_worker(rank, world, n, ticks, mirrors, stamps, flags, out_dir, seed):
    import time

    import oracle

    rng = np.random.default_rng(seed + rank)
    orc = oracle.load_nbody()
    plans = [shard_plan(n, r, world) for r in range(world)]
    rb, re = plans[rank].row_begin, plans[rank].row_end
    pos, mass, vel, acc = scenes.spawn_bodies(scenes.all_heavy_random_cube_with_bodies(n), 3)
    mine_mirror, mine_stamps, mine_flags = mirrors[rank], stamps[rank], flags[rank]

    def wait(flag_row, q, gen):
        t0 = time.time()
        while int(flag_row[q]) < gen:
            time.sleep(0.0002)
            assert time.time() - t0 < 60, "peer never pushed"

    def push(gen):
        for q in range(world):  # "pack" / tail: own rows into every rank's buffer of generation gen, then the flag
            mirrors[q][gen & 1, rb:re] = torch.from_numpy(pos[rb:re])
            stamps[q][gen & 1, rank] = gen
        for q in range(world):
            flags[q][rank] = gen

    import torch

    gen = 1
    push(gen)  # the initial pack
    bad_stamp = False
    for _ in range(ticks):
        time.sleep(float(rng.uniform(0, 0.01)))  # skew
        # interaction kernel: own slice is there; a peer's slice after its flag
        view = np.empty_like(pos)
        view[rb:re] = pos[rb:re]
        for q in [(rank + 1 + k) % world for k in range(world - 1)]:  # rotated order, like the tile walk
            wait(mine_flags, q, gen)
            bad_stamp = bad_stamp or int(mine_stamps[gen & 1, q]) != gen
            b, e = plans[q].row_begin, plans[q].row_end
            view[b:e] = mine_mirror[gen & 1, b:e].numpy()
        acc[rb:re] = orc.gravity(view, mass, rows=(rb, re))[rb:re]
        orc.axpy(pos, vel, orc.dt, rows=(rb, re))
        orc.axpy(vel, acc, orc.dt, rows=(rb, re))
        time.sleep(float(rng.uniform(0, 0.01)))
        # tail: generation gen + 1 overwrites the buffer generation gen - 1 was read from; every rank that has pushed gen
        # has finished reading gen - 1
        for q in range(world):
            wait(mine_flags, q, gen)
        gen += 1
        push(gen)
    np.save(os.path.join(out_dir, f"push_pos_{rank}.npy"), pos[rb:re])
    np.save(os.path.join(out_dir, f"push_vel_{rank}.npy"), vel[rb:re])
    np.save(os.path.join(out_dir, f"push_bad_{rank}.npy"), np.array([bad_stamp]))


def test_peer_push_protocol_model(tmp_path, nbody_oracle):
    import torch
    import torch.multiprocessing as mp

    n, ticks, world = 900, 6, 3
    mirrors = [torch.zeros(2, n, 3).share_memory_() for _ in range(world)]
    stamps = [torch.zeros(2, world, dtype=torch.int64).share_memory_() for _ in range(world)]
    flags = [torch.zeros(world, dtype=torch.int64).share_memory_() for _ in range(world)]
    mp.spawn(_push_worker, args=(world, n, ticks, mirrors, stamps, flags, str(tmp_path), 17), nprocs=world, join=True)
    pos, mass, vel, acc = scenes.spawn_bodies(scenes.all_heavy_random_cube_with_bodies(n), 3)
    nbody_oracle.run_sequential(pos, mass, vel, acc, ticks)
    got_pos = np.concatenate([np.load(tmp_path / f"push_pos_{r}.npy") for r in range(world)])
    got_vel = np.concatenate([np.load(tmp_path / f"push_vel_{r}.npy") for r in range(world)])
    assert not any(np.load(tmp_path / f"push_bad_{r}.npy")[0] for r in range(world)), "a slice of the wrong generation was read"
    assert np.array_equal(got_pos.view(np.uint32), pos.view(np.uint32))
    assert np.array_equal(got_vel.view(np.uint32), vel.view(np.uint32))
