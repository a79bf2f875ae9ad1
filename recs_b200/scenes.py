"""Initial conditions of the reference's case studies with a *seeded* generator.

The reference draws from an unseeded `rand::thread_rng()` (crates/n-body/src/scenes.rs:68), so its
runs are not reproducible; here the same distributions are drawn in the same per-body order from
numpy's PCG64 so the GPU path and the CPU oracle are fed identical bytes.
"""
from __future__ import annotations

from dataclasses import dataclass, replace

import numpy as np


@dataclass(frozen=True)
class Scene:
    """`struct Scene` (crates/n-body/src/scenes.rs:252-264)."""

    body_count: int
    initial_position_min: float
    initial_position_max: float
    position_offset: tuple
    minimum_mass: float
    maximum_mass: float
    initial_velocity_min: float
    initial_velocity_max: float
    initial_acceleration_min: float
    initial_acceleration_max: float


# crates/n-body/src/scenes.rs:10-21
EVERYTHING_HEAVY = Scene(
    body_count=1_000,
    initial_position_min=-100.0,
    initial_position_max=100.0,
    position_offset=(0.0, 0.0, 0.0),
    minimum_mass=5.9000e14,
    maximum_mass=5.9724e14,
    initial_velocity_min=-10_000.0,
    initial_velocity_max=10_000.0,
    initial_acceleration_min=-1.0,
    initial_acceleration_max=1.0,
)
# crates/n-body/src/scenes.rs:23-27
EVERYTHING_LIGHT = replace(EVERYTHING_HEAVY, minimum_mass=100.0, maximum_mass=1_000.0)


def all_heavy_random_cube_with_bodies(body_count: int) -> Scene:
    """crates/n-body/src/scenes.rs:49-55."""
    return replace(EVERYTHING_HEAVY, body_count=body_count)


def _uniform(u: np.ndarray, low: float, high: float) -> np.ndarray:
    low32, high32 = np.float32(low), np.float32(high)
    return (low32 + u * (high32 - low32)).astype(np.float32)


def spawn_bodies(scene: Scene, seed: int):
    """`RandomCubeSpawner::spawn_bodies` (crates/n-body/src/scenes.rs:57-94): per body, in this
    order, position xyz (+offset), mass, velocity xyz, acceleration xyz -- 10 uniform draws.

    Returns (position[N,3], mass[N], velocity[N,3], acceleration[N,3]) float32, row k = body k,
    which `create_planet_entity` (scenes.rs:108-117) inserts as entity id k / component index k.
    """
    n = scene.body_count
    rng = np.random.default_rng(seed)
    u = rng.random((n, 10), dtype=np.float32)
    pos = _uniform(u[:, 0:3], scene.initial_position_min, scene.initial_position_max)
    pos = (pos + np.asarray(scene.position_offset, dtype=np.float32)).astype(np.float32)
    mass = _uniform(u[:, 3], scene.minimum_mass, scene.maximum_mass)
    vel = _uniform(u[:, 4:7], scene.initial_velocity_min, scene.initial_velocity_max)
    acc = _uniform(u[:, 7:10], scene.initial_acceleration_min, scene.initial_acceleration_max)
    return (
        np.ascontiguousarray(pos),
        np.ascontiguousarray(mass),
        np.ascontiguousarray(vel),
        np.ascontiguousarray(acc),
    )


def rain_drops(n: int, seed: int):
    """Steady-state raindrop archetype {Velocity, Mass, Position} for the streaming benchmark
    (SURVEY 8d config 4): clouds sit on the grid of `create_evenly_interspersed_clouds`
    (crates/rain-simulation/src/scene.rs:27-54: x*10, 10, z*10), drops start below them with zero
    velocity (`Velocity::default()`, crates/rain-simulation/src/lib.rs:139) and mass 0.1 (:105)."""
    rng = np.random.default_rng(seed)
    side = max(int(np.sqrt(max(n // 64, 1))), 1)
    cx = rng.integers(0, side, size=n).astype(np.float32) * np.float32(10.0)
    cz = rng.integers(0, side, size=n).astype(np.float32) * np.float32(10.0)
    y = _uniform(rng.random(n, dtype=np.float32), 0.5, 10.0)
    pos = np.ascontiguousarray(np.stack([cx, y, cz], axis=1).astype(np.float32))
    vel = np.zeros((n, 3), dtype=np.float32)
    mass = np.full(n, 0.1, dtype=np.float32)
    return vel, mass, pos
