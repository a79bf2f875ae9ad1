"""The CUDA path against the committed golden vectors (tests/golden/nbody_golden.npz, oracle-generated: see
tests/golden/make_nbody_golden.py and DESIGN.md 7 for what they do and do not pin)."""
import os

import numpy as np
import pytest

from recs_b200 import _lib, scenes
from tests.util import rel_err_rows

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "nbody_golden.npz")


def test_nbody_config0_against_golden(gpu_ctx):
    """configs[0]: 1 000 bodies, EVERYTHING_HEAVY.  Tick 1: positions bit-exact, velocity / acceleration within 1e-4
    norm-wise; tick 100 (MINIMUM_NUMBER_OF_TICKS): drift bound 1e-4 of max(|x|, 1)."""
    from recs_b200.nbody import NBodyGpuApp

    g = np.load(GOLDEN)
    pos, mass, vel, acc = scenes.spawn_bodies(scenes.EVERYTHING_HEAVY, 1)
    app = NBodyGpuApp(gpu_ctx, pos, mass, vel, acc)
    app.upload()
    app.tick(1)
    app.download()
    assert np.array_equal(app.pos.view(np.uint32), g["heavy1000_seed1_t1_pos"].view(np.uint32))
    assert rel_err_rows(app.acc, g["heavy1000_seed1_t1_acc"]).max() <= 1e-4
    assert rel_err_rows(app.vel, g["heavy1000_seed1_t1_vel"]).max() <= 1e-4
    app.tick(99)
    app.download()
    assert rel_err_rows(app.pos, g["heavy1000_seed1_t100_pos"], floor=1.0).max() <= 1e-4
    assert rel_err_rows(app.vel, g["heavy1000_seed1_t100_vel"], floor=1.0).max() <= 1e-4


def test_light_scene_gravity_against_golden(gpu_ctx):
    from recs_b200.nbody import NBodyGpuApp

    g = np.load(GOLDEN)
    pos, mass, vel, acc = scenes.spawn_bodies(scenes.replace(scenes.EVERYTHING_LIGHT, body_count=777), 2)
    app = NBodyGpuApp(gpu_ctx, pos, mass, vel, acc)
    app.upload()
    app.run("gravity")
    app.download()
    assert rel_err_rows(app.acc, g["light777_seed2_gravity"]).max() <= 1e-4


def test_rain_ticks_against_golden_bitwise(gpu_ctx):
    """wind -> gravity -> movement with the rotating wind direction, 5 ticks, 1 000 drops: bit-exact."""
    g = np.load(GOLDEN)
    VEL, POS, ARCH = 1, 2, 1
    vel, mass, pos = scenes.rain_drops(1000, 3)
    vel[:] = np.random.default_rng(3).standard_normal(vel.shape).astype(np.float32) * np.float32(5.0)
    ctx = gpu_ctx
    ctx.bind_column(ARCH, VEL, vel)
    ctx.bind_column(ARCH, POS, pos)
    ctx.upload(ARCH, VEL)
    ctx.upload(ARCH, POS)
    w = ctx.create_system(_lib.SYS_RAIN_WIND, [VEL])
    d = ctx.create_system(_lib.SYS_RAIN_WIND_DIRECTION, [])
    gr = ctx.create_system(_lib.SYS_RAIN_GRAVITY, [VEL])
    m = ctx.create_system(_lib.SYS_RAIN_MOVEMENT, [POS, VEL])
    for s in (w, gr, m):
        ctx.set_archetypes(s, [ARCH])
    ctx.set_tick_order([w, d, gr, m])
    ctx.tick(5)
    ctx.download(ARCH, VEL)
    ctx.download(ARCH, POS)
    assert np.array_equal(vel.view(np.uint32), g["rain1000_seed3_t5_vel"].view(np.uint32))
    assert np.array_equal(pos.view(np.uint32), g["rain1000_seed3_t5_pos"].view(np.uint32))
    assert np.array_equal(np.array(ctx.get_wind_direction(), np.float32).view(np.uint32), g["rain_wind_direction_t5"].view(np.uint32))
