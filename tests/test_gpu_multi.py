"""Index-sharded n-body on >= 2 GPUs (one per rank): peer-push exchange and NCCL exchange; skipped on a single-GPU box."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpu_count():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True).stdout
        return sum(1 for line in out.splitlines() if line.startswith("GPU "))
    except OSError:
        return 0


def _torchrun(world, script, args, env=None, timeout=420):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), script] + [str(a) for a in args]
    full_env = dict(os.environ)
    full_env.update(env or {})
    return subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=timeout, env=full_env)


@pytest.mark.parametrize("bodies,ticks,mode", [(10000, 1, "full"), (5000, 4, "full"), (300, 2, "full"), (7000, 3, "own"),
                                               (30000, 2, "own"),   # shards below the small-shape limit of a single GPU
                                               (70000, 2, "own")])
@pytest.mark.parametrize("exchange", ["p2p", "nccl"])
def test_sharded_ticks_match_oracle(bodies, ticks, mode, exchange):
    """mode "own": every rank uploads / downloads only the rows it owns (recs_gpu_upload_rows) and its host copy of
    every other row is NaN -- nothing outside the shard may be read from the host.  Every rank also compares its rows
    with the single-GPU result of the same ticks bit for bit.

    One GPU per rank, always: kernels of different ranks wait on one another (generation flags, NCCL), and several such
    ranks time-slicing ONE GPU can end in a context-switch timeout (Xid 109, B200_PROFILING.md).  What a single GPU can
    check of the sharded walk -- the two-phase, rotated tile order and its summation tree -- is in
    tests/test_gpu_nbody.py::test_result_does_not_depend_on_grid_or_rank_split; the host-side protocol is in
    tests/test_sharding.py (gloo, CPU)."""
    n_gpu = _gpu_count()
    if n_gpu < 2:
        pytest.skip("needs at least 2 GPUs (one per rank)")
    world = 2 if n_gpu < 4 else 4
    env = {}
    proc = _torchrun(world, "tests/tools/multi_gpu_check.py", [bodies, ticks, mode, exchange], env)
    assert proc.returncode == 0 and "MULTI_GPU_CHECK OK" in proc.stdout, proc.stdout[-2000:] + proc.stderr[-2000:]


def test_sharded_streaming_systems_need_no_exchange():
    """Rain tick, simple_iter and a generic system: each rank streams its own contiguous slice of every archetype
    (independent units, no data-path collective), bit-exact, other rows untouched."""
    n_gpu = _gpu_count()
    if n_gpu < 2:
        pytest.skip("needs at least 2 GPUs (one per rank)")
    world = 2 if n_gpu < 4 else 4
    proc = _torchrun(world, "tests/tools/multi_gpu_stream_check.py", [100003])
    assert proc.returncode == 0 and "MULTI_GPU_STREAM_CHECK OK" in proc.stdout, proc.stdout[-2000:] + proc.stderr[-2000:]


def test_generic_nested_query_systems_on_a_sharded_context():
    """recs_gpu_system_create_pair_kernel with world_size > 1: own outer rows x the whole nested query, the written
    column gathered with recs_gpu_column_allgather; bit-exact against the oracle after several ticks."""
    n_gpu = _gpu_count()
    if n_gpu < 2:
        pytest.skip("needs at least 2 GPUs (NCCL)")
    world = 2 if n_gpu < 4 else 4
    proc = _torchrun(world, "tests/tools/multi_gpu_pair_check.py", [6000, 3])
    assert proc.returncode == 0 and "MULTI_GPU_PAIR_CHECK OK" in proc.stdout, proc.stdout[-2000:] + proc.stderr[-2000:]
