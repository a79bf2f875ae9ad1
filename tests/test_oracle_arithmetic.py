"""The float side of the oracle, checked without a GPU.

The reference's own tests assert nothing about the numbers its n-body / rain systems produce
(crates/ecs_bench_suite/tests/debug_benchmarks.rs:33-37) and the Rust reference cannot be built here, so float
parity is UNPINNED against recs itself.  What can be pinned is the restatement:

* a second, independent restatement of the same Rust expressions, written here with numpy float32 scalars one
  operation at a time, must agree with oracle/nbody_oracle.c bit for bit (operation order, no FMA contraction,
  sequential `.sum()` from zero, `<= f32::EPSILON` mask);
* hand-derived known answers (3-4-5 triangle, masked pairs, lone body);
* the committed oracle-generated vectors of tests/golden/nbody_golden.npz (drift of flags / compiler).
"""
import os

import numpy as np
import pytest

from recs_b200 import scenes

f32 = np.float32
G = f32(6.67e-11)  # crates/n-body/src/lib.rs:123
EPS = np.finfo(np.float32).eps  # f32::EPSILON, :118
DT = f32(1.0) / f32(20000.0)  # FIXED_TIME_STEP, :14
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "nbody_golden.npz")


def rust_gravity(pos, mass):
    """crates/n-body/src/lib.rs:104-135 one f32 operation at a time (cgmath 0.18: `Point3 - Point3`,
    `magnitude2 = (x*x + y*y) + z*z`, `normalize_to(m) = v * (m / magnitude())`, `Sum` = left fold from zero)."""
    n = len(pos)
    out = np.zeros((n, 3), np.float32)
    for i in range(n):
        total = [f32(0), f32(0), f32(0)]
        for j in range(n):
            to_body = [f32(pos[j][k]) - f32(pos[i][k]) for k in range(3)]
            d2 = (to_body[0] * to_body[0] + to_body[1] * to_body[1]) + to_body[2] * to_body[2]
            if d2 <= EPS:
                contribution = [f32(0), f32(0), f32(0)]
            else:
                acceleration = (G * f32(mass[j])) / d2
                scale = acceleration / np.sqrt(d2)
                contribution = [to_body[k] * scale for k in range(3)]
            total = [total[k] + contribution[k] for k in range(3)]
        out[i] = total
    return out


@pytest.mark.parametrize("n,seed", [(1, 1), (2, 2), (17, 3), (64, 4)])
def test_gravity_oracle_equals_independent_restatement_bitwise(nbody_oracle, n, seed):
    pos, mass, _, _ = scenes.spawn_bodies(scenes.all_heavy_random_cube_with_bodies(n), seed)
    if n >= 17:
        pos[5] = pos[2]  # coincident
        pos[9] = pos[3] + np.array([1e-4, 0, 0], np.float32)  # below epsilon
    with np.errstate(all="ignore"):
        want = rust_gravity(pos, mass)
    got = nbody_oracle.gravity(pos, mass)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_gravity_known_answer_3_4_5(nbody_oracle):
    pos = np.array([[0, 0, 0], [3, 4, 0]], np.float32)
    mass = np.array([2.0e10, 5.0e10], np.float32)
    acc = nbody_oracle.gravity(pos, mass)
    # |d| = 5 exactly: a_0 = G m_1 / 25 / 5 * (3, 4, 0), a_1 = -G m_0 / 125 * (3, 4, 0)
    s0 = (G * f32(5.0e10)) / f32(25.0) / f32(5.0)
    s1 = (G * f32(2.0e10)) / f32(25.0) / f32(5.0)
    assert np.array_equal(acc[0], np.array([f32(3) * s0, f32(4) * s0, f32(0)], np.float32))
    assert np.array_equal(acc[1], np.array([f32(-3) * s1, f32(-4) * s1, f32(0) * s1], np.float32))
    assert abs(float(acc[0][0]) - 6.67e-11 * 5e10 / 125 * 3) < 1e-7 * abs(float(acc[0][0]))


def test_gravity_epsilon_boundary(nbody_oracle):
    # d2 == EPSILON is masked, the next float above is not (`<=`, crates/n-body/src/lib.rs:118)
    d_masked = np.sqrt(np.float64(EPS))
    for dx, masked in ((f32(2.0**-12), True), (f32(2.0**-11.5), True), (f32(2.0**-11), False)):
        pos = np.array([[0, 0, 0], [dx, 0, 0]], np.float32)
        d2 = f32(dx) * f32(dx)
        assert (d2 <= EPS) == masked, (dx, d_masked)
        acc = nbody_oracle.gravity(pos, np.array([1e15, 1e15], np.float32))
        assert (acc == 0).all() == masked
    # a lone body ends with exactly zero: `.sum()` of nothing but its own masked pair
    assert np.array_equal(nbody_oracle.gravity(np.ones((1, 3), np.float32), np.ones(1, np.float32)), np.zeros((1, 3), np.float32))


def test_integration_is_mul_then_add(nbody_oracle):
    """`position.point += velocity * FIXED_TIME_STEP` (crates/n-body/src/lib.rs:92): rustc never fuses."""
    rng = np.random.default_rng(5)
    pos = (rng.standard_normal((4000, 3)) * 100).astype(np.float32)
    vel = (rng.standard_normal((4000, 3)) * 1e4).astype(np.float32)
    want = pos + vel * DT
    fused = (pos.astype(np.float64) + vel.astype(np.float64) * np.float64(DT)).astype(np.float32)
    assert not np.array_equal(want, fused)
    got = pos.copy()
    nbody_oracle.axpy(got, vel, nbody_oracle.dt)
    assert np.float32(nbody_oracle.dt) == DT
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_rain_systems_equal_independent_restatement_bitwise(nbody_oracle):
    """crates/rain-simulation/src/lib.rs:62-96 in numpy float32."""
    rng = np.random.default_rng(6)
    vel = (rng.standard_normal((3000, 3)) * 6).astype(np.float32)  # some above the terminal speed of 9
    vel[0] = (0, 9.0, 0)  # abs(y) >= 9 skips gravity; |v| >= 9 skips wind
    vel[1] = (0, -9.0, 0)
    vel[2] = (9.0, 0, 0)
    pos = rng.standard_normal((3000, 3)).astype(np.float32)
    direction = np.array([0.6, 0.0, 0.8], np.float32)
    # gravity: if abs(v.y) >= 9 { return } v += (-9.82 * 0.05) * unit_y                      :65-70
    k = f32(-9.82) * f32(0.05)
    want = vel.copy()
    go = np.abs(want[:, 1]) < f32(9.0)
    want[go] = want[go] + np.array([k * f32(0), k * f32(1), k * f32(0)], np.float32)
    got = vel.copy()
    nbody_oracle.rain_gravity(got)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    # wind: if v.magnitude() >= 9 { return } v += (10 * 0.05) * direction                    :86-92
    c = f32(10.0) * f32(0.05)
    w = np.array([c * direction[0], c * direction[1], c * direction[2]], np.float32)
    mag = np.sqrt((vel[:, 0] * vel[:, 0] + vel[:, 1] * vel[:, 1]) + vel[:, 2] * vel[:, 2])
    want = vel.copy()
    go = mag < f32(9.0)
    want[go] = want[go] + w
    got = vel.copy()
    nbody_oracle.rain_wind(got, direction)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    # movement: p += v * 0.05                                                                :94-96
    want = pos + vel * f32(0.05)
    got = pos.copy()
    nbody_oracle.rain_movement(got, vel)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    # wind_direction: rotation about +y by 10 deg/s * 0.05 s keeps the length and turns by 0.5 degrees   :78-84
    d = np.array([1.0, 0.0, 0.0], np.float32)
    for _ in range(720):
        d = nbody_oracle.rain_wind_direction(d)
    assert abs(float(np.linalg.norm(d)) - 1.0) < 1e-4
    assert np.allclose(d, [1.0, 0.0, 0.0], atol=2e-3)  # 720 x 0.5 degrees = one full turn
    one = nbody_oracle.rain_wind_direction(np.array([1.0, 0.0, 0.0], np.float32))
    assert np.allclose(one, [np.cos(np.radians(0.5)), 0.0, -np.sin(np.radians(0.5))], atol=1e-6)


def test_oracle_reproduces_committed_golden_vectors_bitwise(nbody_oracle):
    """tests/golden/nbody_golden.npz (made by tests/golden/make_nbody_golden.py): the oracle built on this machine
    must reproduce it bit for bit -- a different result means its flags / compiler changed the arithmetic."""
    g = np.load(GOLDEN)
    for ticks in (1, 100):
        pos, mass, vel, acc = scenes.spawn_bodies(scenes.EVERYTHING_HEAVY, 1)
        nbody_oracle.run_sequential(pos, mass, vel, acc, ticks)
        for name, arr in (("pos", pos), ("vel", vel), ("acc", acc)):
            assert np.array_equal(arr.view(np.uint32), g[f"heavy1000_seed1_t{ticks}_{name}"].view(np.uint32)), (ticks, name)
    pos, mass, _, _ = scenes.spawn_bodies(scenes.replace(scenes.EVERYTHING_LIGHT, body_count=777), 2)
    assert np.array_equal(nbody_oracle.gravity(pos, mass).view(np.uint32), g["light777_seed2_gravity"].view(np.uint32))


def test_f64_oracle_bounds_the_f32_reference_order_sum(nbody_oracle):
    pos, mass, _, _ = scenes.spawn_bodies(scenes.all_heavy_random_cube_with_bodies(2000), 7)
    a32 = nbody_oracle.gravity(pos, mass).astype(np.float64)
    a64 = nbody_oracle.gravity_f64(pos, mass)
    err = np.linalg.norm(a32 - a64, axis=1) / np.linalg.norm(a64, axis=1)
    assert err.max() < 2e-5  # the reference's own sequential f32 sum at 2 000 bodies


def test_total_energy_equals_numpy_f64_and_is_thread_independent(nbody_oracle):
    """SURVEY 8d's energy-drift diagnostic: E = sum 1/2 m v^2 - sum_{i<j} G m_i m_j / r, pairs the force law masks left out."""
    from recs_b200 import scenes

    n = 700
    pos, mass, vel, _ = scenes.spawn_bodies(scenes.all_heavy_random_cube_with_bodies(n), 4)
    pos[11] = pos[10]  # a coincident pair: no force in the reference (lib.rs:118-120), no potential here
    k, u = nbody_oracle.total_energy(pos, mass, vel, threads=3)
    p, m = pos.astype(np.float64), mass.reshape(-1).astype(np.float64)
    d = p[None, :, :] - p[:, None, :]
    d2 = (d * d).sum(-1)
    i, j = np.triu_indices(n, 1)
    keep = d2[i, j] > np.float64(np.finfo(np.float32).eps)
    assert keep.sum() == i.size - 1
    g = np.float64(np.float32(6.67e-11))
    want_u = -(g * m[i[keep]] * m[j[keep]] / np.sqrt(d2[i, j][keep])).sum()
    want_k = (0.5 * m * (vel.astype(np.float64) ** 2).sum(-1)).sum()
    assert abs(k - want_k) <= 1e-12 * abs(want_k) and abs(u - want_u) <= 1e-12 * abs(want_u)
    k1, u1 = nbody_oracle.total_energy(pos, mass, vel, threads=1)
    assert abs(k1 - k) <= 1e-13 * abs(k) and abs(u1 - u) <= 1e-13 * abs(u)
