"""examples/nbody_bench.c: the reference's n-body benchmark assembled in plain C on the C ABI alone (no Python in the
process).  CPU: both public headers compile as C and the example links against librecs_gpu.so; without a GPU it stops
with RECS_ERR_NO_DEVICE (no fallback).  GPU: it runs, the schedule yields gravity > movement > acceleration and the
column checksums match the CPU oracle on the same xorshift draws."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "examples", "nbody_bench")


def _build():
    subprocess.run(["make", "-C", ROOT, "examples"], check=True, capture_output=True)
    assert os.path.exists(BIN)


def test_headers_are_c_and_example_links():
    for header in ("recs_gpu.h", "recs_ecs.h"):
        src = f'#include "{header}"\nint main(void) {{ return 0; }}\n'
        proc = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), "-x", "c", "-"],
                              input=src, capture_output=True, text=True)
        assert proc.returncode == 0, proc.stderr
    _build()


def test_example_refuses_to_run_without_a_gpu():
    from tests.conftest import HAS_GPU

    if HAS_GPU:
        pytest.skip("a GPU is present")
    _build()
    proc = subprocess.run([BIN, "100", "1"], capture_output=True, text=True)
    assert proc.returncode != 0 and "-> 2" in proc.stderr  # RECS_ERR_NO_DEVICE


def _xorshift_bodies(n):
    """The draws of examples/nbody_bench.c (xorshift64*, 24-bit uniforms), per body: pos xyz, mass, vel xyz, acc xyz."""
    mask = (1 << 64) - 1
    state = 0x9E3779B97F4A7C15
    lo = np.array([-100.0] * 3 + [5.9000e14] + [-10000.0] * 3 + [-1.0] * 3, np.float32)
    hi = np.array([100.0] * 3 + [5.9724e14] + [10000.0] * 3 + [1.0] * 3, np.float32)
    out = np.empty((n, 10), np.float32)
    for i in range(n):
        for k in range(10):
            state ^= state >> 12
            state ^= (state << 25) & mask
            state ^= state >> 27
            r = (state * 0x2545F4914F6CDD1D) & mask
            u = np.float32(r >> 40) / np.float32(16777216.0)
            out[i, k] = lo[k] + u * (hi[k] - lo[k])
    return (np.ascontiguousarray(out[:, 0:3]), np.ascontiguousarray(out[:, 3]), np.ascontiguousarray(out[:, 4:7]),
            np.ascontiguousarray(out[:, 7:10]))


@pytest.mark.gpu
def test_c_example_matches_the_oracle(nbody_oracle):
    _build()
    n, ticks = 3000, 4
    proc = subprocess.run([BIN, str(n), str(ticks)], capture_output=True, text=True, timeout=300)
    assert proc.returncode == 0, proc.stderr
    line = [l for l in proc.stdout.splitlines() if l.startswith("NBODY_BENCH")][0]
    fields = dict(re.findall(r"(\w+)=(\S+)", line))
    assert fields["order"] == "gravity>movement>acceleration"
    assert int(fields["bodies"]) == n and int(fields["ticks"]) == ticks + 1
    pos, mass, vel, acc = _xorshift_bodies(n)
    nbody_oracle.run_sequential(pos, mass, vel, acc, ticks + 1)
    for name, arr in (("sum_abs_pos", pos), ("sum_abs_vel", vel), ("sum_abs_acc", acc)):
        want = float(np.abs(arr.astype(np.float64)).sum())
        assert abs(float(fields[name]) - want) <= 1e-5 * want, (name, fields[name], want)
