"""Index-sharded n-body: peer-push exchange (also on a single-GPU box, ranks time-slicing it) and NCCL exchange (>= 2 GPUs)."""
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
                                               (70000, 2, "own")])
@pytest.mark.parametrize("exchange", ["p2p", "nccl"])
def test_sharded_ticks_match_oracle(bodies, ticks, mode, exchange):
    """mode "own": every rank uploads / downloads only the rows it owns (recs_gpu_upload_rows) and its host copy of
    every other row is NaN -- nothing outside the shard may be read from the host.  Every rank also compares its rows
    with the single-GPU result of the same ticks bit for bit.

    On a single-GPU box the ranks are separate processes time-slicing GPU 0, which covers the peer-push exchange
    (CUDA IPC mapping, generation flags, one persistent launch per tick); the NCCL exchange needs one GPU per rank."""
    n_gpu = _gpu_count()
    if n_gpu < 1:
        pytest.skip("needs a GPU")
    if n_gpu < 2 and exchange == "nccl":
        pytest.skip("NCCL needs one GPU per rank")
    world = 2 if n_gpu < 4 else 4
    env = {"RECS_TEST_SAME_DEVICE": "1"} if n_gpu < 2 else {}
    proc = _torchrun(world, "tests/tools/multi_gpu_check.py", [bodies, ticks, mode, exchange], env)
    assert proc.returncode == 0 and "MULTI_GPU_CHECK OK" in proc.stdout, proc.stdout[-2000:] + proc.stderr[-2000:]


def test_sharded_streaming_systems_need_no_exchange():
    """Rain tick, simple_iter and a generic system: each rank streams its own contiguous slice of every archetype
    (independent units, no data-path collective), bit-exact, other rows untouched."""
    n_gpu = _gpu_count()
    if n_gpu < 1:
        pytest.skip("needs a GPU")
    world = 2 if n_gpu < 4 else 4
    proc = _torchrun(world, "tests/tools/multi_gpu_stream_check.py", [100003], {"RECS_TEST_SAME_DEVICE": "1"} if n_gpu < 2 else {})
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
