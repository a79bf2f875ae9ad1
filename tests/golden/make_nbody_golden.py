"""Generates tests/golden/nbody_golden.npz from the CPU oracle (oracle/nbody_oracle.c).

These vectors are ORACLE-generated, not reference-generated: the Rust reference cannot be built here,
so they pin the restatement against drift (compiler flags, FMA contraction, summation order), not against recs
itself.  Inputs are the seeded scenes of recs_b200/scenes.py (crates/n-body/src/scenes.rs:10-21, 57-94).

    python tests/golden/make_nbody_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from recs_b200 import scenes  # noqa: E402

orc = oracle.load_nbody()
out = {}
# configs[0]: 1 000 bodies, EVERYTHING_HEAVY, 1 tick and 100 ticks (MINIMUM_NUMBER_OF_TICKS) in schedule order
for ticks in (1, 100):
    pos, mass, vel, acc = scenes.spawn_bodies(scenes.EVERYTHING_HEAVY, 1)
    orc.run_sequential(pos, mass, vel, acc, ticks)
    out[f"heavy1000_seed1_t{ticks}_pos"] = pos
    out[f"heavy1000_seed1_t{ticks}_vel"] = vel
    out[f"heavy1000_seed1_t{ticks}_acc"] = acc
# gravity alone on the light scene, 777 bodies (ragged tile)
pos, mass, vel, acc = scenes.spawn_bodies(scenes.replace(scenes.EVERYTHING_LIGHT, body_count=777), 2)
out["light777_seed2_gravity"] = orc.gravity(pos, mass)
# rain systems: 1 000 drops, 5 ticks of wind -> gravity -> movement with the rotating wind direction
vel, mass, pos = scenes.rain_drops(1000, 3)
vel[:] = np.random.default_rng(3).standard_normal(vel.shape).astype(np.float32) * np.float32(5.0)
direction = np.array([1.0, 0.0, 0.0], np.float32)
for _ in range(5):
    orc.rain_wind(vel, direction)
    direction = orc.rain_wind_direction(direction)
    orc.rain_gravity(vel)
    orc.rain_movement(pos, vel)
out["rain1000_seed3_t5_vel"] = vel
out["rain1000_seed3_t5_pos"] = pos
out["rain_wind_direction_t5"] = direction
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "nbody_golden.npz"), **out)
print({k: v.shape for k, v in out.items()})
