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


# crates/n-body/src/scenes.rs:29-40
ONE_HEAVY_BODY = Scene(
    body_count=1,
    initial_position_min=-10.0,
    initial_position_max=10.0,
    position_offset=(0.0, 0.0, 0.0),
    minimum_mass=1.900e20,
    maximum_mass=1.989e20,
    initial_velocity_min=0.0,
    initial_velocity_max=1e-30,
    initial_acceleration_min=0.0,
    initial_acceleration_max=1e-30,
)


@dataclass(frozen=True)
class ClusterSpawner:
    """`struct ClusterSpawner` (crates/n-body/src/scenes.rs:171-177)."""

    cluster_count: int
    cluster_size: float
    cluster_distance: float
    scene: Scene


# crates/n-body/src/scenes.rs:179-206
SMALL_HEAVY_CLUSTERS = ClusterSpawner(cluster_count=4, cluster_size=10.0, cluster_distance=100.0, scene=EVERYTHING_HEAVY)
SMALL_LIGHT_CLUSTERS = replace(SMALL_HEAVY_CLUSTERS, scene=EVERYTHING_LIGHT)
HUGE_HEAVY_CLUSTERS = ClusterSpawner(cluster_count=4, cluster_size=100.0, cluster_distance=1000.0,
                                     scene=replace(EVERYTHING_HEAVY, minimum_mass=100_000.0, maximum_mass=1_000_000.0))
HUGE_LIGHT_CLUSTERS = replace(HUGE_HEAVY_CLUSTERS, scene=EVERYTHING_LIGHT)


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
    rng = seed if isinstance(seed, np.random.Generator) else np.random.default_rng(seed)
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


def spawn_clusters(spawner: ClusterSpawner, seed: int):
    """`ClusterSpawner::spawn_bodies` (crates/n-body/src/scenes.rs:208-250): per cluster three draws in [-1, 1) for a
    direction, normalised; the cluster sits at direction * cluster_distance from the ORIGIN (`previous_cluster_position`
    is never advanced, :224) and is a random cube of +- cluster_size with body_count / cluster_count bodies.  One seeded
    generator supplies the draws in the reference's order.  Returns the concatenated columns like `spawn_bodies`."""
    rng = np.random.default_rng(seed)
    parts = []
    for _ in range(spawner.cluster_count):
        d = (np.float32(-1.0) + rng.random(3, dtype=np.float32) * np.float32(2.0)).astype(np.float32)
        mag = np.float32(np.sqrt(np.float32(np.float32(d[0] * d[0] + d[1] * d[1]) + d[2] * d[2])))
        d = d * (np.float32(1.0) / mag)  # normalize() = self * (1 / magnitude)
        centre = tuple(float(x) for x in (d * np.float32(spawner.cluster_distance)))
        scene = replace(spawner.scene, body_count=spawner.scene.body_count // spawner.cluster_count, position_offset=centre,
                        initial_position_min=-spawner.cluster_size, initial_position_max=spawner.cluster_size)
        parts.append(spawn_bodies(scene, rng))
    return tuple(np.ascontiguousarray(np.concatenate([p[i] for p in parts])) for i in range(4))


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
