"""Index-sharded n-body over NCCL on >= 2 GPUs (skipped on a single-GPU box)."""
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


@pytest.mark.parametrize("bodies,ticks,mode", [(10000, 1, "full"), (5000, 4, "full"), (300, 2, "full"), (7000, 3, "own")])
def test_sharded_ticks_match_oracle(bodies, ticks, mode):
    """mode "own": every rank uploads / downloads only the rows it owns (recs_gpu_upload_rows) and its host copy of
    every other row is NaN -- nothing outside the shard may be read from the host."""
    n_gpu = _gpu_count()
    if n_gpu < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n_gpu < 4 else 4
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), "tests/tools/multi_gpu_check.py", str(bodies), str(ticks), mode]
    proc = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert proc.returncode == 0 and "MULTI_GPU_CHECK OK" in proc.stdout, proc.stdout[-2000:] + proc.stderr[-2000:]


def test_sharded_streaming_systems_need_no_exchange():
    """Rain tick, simple_iter and a generic system: each rank streams its own contiguous slice of every archetype
    (independent units, no data-path collective), bit-exact, other rows untouched."""
    n_gpu = _gpu_count()
    if n_gpu < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n_gpu < 4 else 4
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), "tests/tools/multi_gpu_stream_check.py", "100003"]
    proc = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert proc.returncode == 0 and "MULTI_GPU_STREAM_CHECK OK" in proc.stdout, proc.stdout[-2000:] + proc.stderr[-2000:]
