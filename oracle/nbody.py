"""ctypes wrapper of oracle/liboracle.so (oracle/nbody_oracle.c) -- test infrastructure only."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")

SYS_GRAVITY, SYS_MOVEMENT, SYS_ACCELERATION = 0, 1, 2
# tick order the reference's PrecedenceGraph produces for the benchmark's system set
ORDER_BENCH = (SYS_GRAVITY, SYS_MOVEMENT, SYS_ACCELERATION)

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "nbody_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _SO


class NBodyOracle:
    def __init__(self, lib: C.CDLL):
        self.lib = lib
        z = C.c_size_t
        lib.oracle_gravity_rows.argtypes = [_f32p, _f32p, z, z, _f32p, _f32p, z]
        lib.oracle_gravity_rows.restype = None
        lib.oracle_gravity_rows_f64.argtypes = [_f32p, _f64p, z, z, _f32p, _f32p, z]
        lib.oracle_gravity_rows_f64.restype = None
        lib.oracle_gravity_sample.argtypes = [_f32p, _u64p, z, _f32p, _f32p, z, C.c_void_p, C.c_void_p]
        lib.oracle_gravity_sample.restype = None
        lib.oracle_gravity_sample_mt.argtypes = [_f32p, _u64p, z, _f32p, _f32p, z, C.c_void_p, C.c_void_p, C.c_int]
        lib.oracle_gravity_sample_mt.restype = None
        lib.oracle_total_energy_mt.argtypes = [_f64p, _f32p, _f64p, z, C.c_int, _f64p]
        lib.oracle_total_energy_mt.restype = None
        lib.oracle_axpy_rows.argtypes = [_f32p, _f32p, C.c_float, z, z]
        lib.oracle_axpy_rows.restype = None
        lib.oracle_add_rows.argtypes = [_f32p, _f32p, z, z]
        lib.oracle_add_rows.restype = None
        lib.oracle_nbody_dt.restype = C.c_float
        lib.oracle_rain_gravity.argtypes = [_f32p, z, z]
        lib.oracle_rain_gravity.restype = None
        lib.oracle_rain_wind.argtypes = [_f32p, z, z, _f32p]
        lib.oracle_rain_wind.restype = None
        lib.oracle_rain_movement.argtypes = [_f32p, _f32p, z, z]
        lib.oracle_rain_movement.restype = None
        lib.oracle_rain_wind_direction.argtypes = [_f32p]
        lib.oracle_rain_wind_direction.restype = None
        i3 = C.c_int * 3
        lib.oracle_nbody_run_pool.argtypes = [_f32p, _f32p, _f32p, _f32p, z, C.c_int, C.c_int, i3]
        lib.oracle_nbody_run_pool.restype = C.c_double
        lib.oracle_nbody_run_sequential.argtypes = [_f32p, _f32p, _f32p, _f32p, z, C.c_int, i3]
        lib.oracle_nbody_run_sequential.restype = None
        lib.oracle_nbody_run_f64.argtypes = [_f64p, _f32p, _f64p, _f64p, z, C.c_int, i3]
        lib.oracle_nbody_run_f64.restype = None
        self.dt = float(lib.oracle_nbody_dt())

    # ---- n-body ------------------------------------------------------------------------------
    def gravity(self, pos, mass, qpos=None, qmass=None, rows=None):
        """acc[N,3] (f32, reference summation order) for `rows` (default all)."""
        qpos = pos if qpos is None else qpos
        qmass = mass if qmass is None else qmass
        acc = np.zeros_like(pos)
        b, e = (0, pos.shape[0]) if rows is None else rows
        self.lib.oracle_gravity_rows(pos, acc, b, e, qpos, qmass, qpos.shape[0])
        return acc

    def gravity_f64(self, pos, mass, rows=None):
        acc = np.zeros(pos.shape, dtype=np.float64)
        b, e = (0, pos.shape[0]) if rows is None else rows
        self.lib.oracle_gravity_rows_f64(pos, acc, b, e, pos, mass, pos.shape[0])
        return acc

    def gravity_sample(self, pos, mass, rows, f64=True, threads=None):
        """Reference-order f32 (and f64) accelerations of the sampled `rows` against all bodies; the rows are
        spread over `threads` host threads (default: all), which does not change a row's result."""
        rows = np.ascontiguousarray(rows, dtype=np.uint64)
        out32 = np.zeros((rows.size, 3), dtype=np.float32)
        out64 = np.zeros((rows.size, 3), dtype=np.float64) if f64 else None
        if threads is None:
            threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        self.lib.oracle_gravity_sample_mt(
            pos, rows, rows.size, pos, mass, pos.shape[0], out32.ctypes.data, out64.ctypes.data if f64 else None, int(threads)
        )
        return out32, out64

    def total_energy(self, pos, mass, vel, threads=None):
        """(kinetic, potential) of the whole set in f64 (SURVEY 8d's energy-drift diagnostic); f32 columns are widened."""
        pos64 = np.ascontiguousarray(pos, dtype=np.float64)
        vel64 = np.ascontiguousarray(vel, dtype=np.float64)
        mass = np.ascontiguousarray(mass, dtype=np.float32).reshape(-1)
        if threads is None:
            threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        out = np.zeros(2, dtype=np.float64)
        self.lib.oracle_total_energy_mt(pos64, mass, vel64, pos64.shape[0], int(min(threads, 256)), out)
        return float(out[0]), float(out[1])

    def axpy(self, y, x, a, rows=None):
        b, e = (0, y.shape[0]) if rows is None else rows
        self.lib.oracle_axpy_rows(y, x, a, b, e)

    def add(self, y, x):
        self.lib.oracle_add_rows(y, x, 0, y.shape[0])

    def run_sequential(self, pos, mass, vel, acc, ticks, order=ORDER_BENCH):
        self.lib.oracle_nbody_run_sequential(pos, mass, vel, acc, pos.shape[0], ticks, (C.c_int * 3)(*order))

    def run_pool(self, pos, mass, vel, acc, ticks, threads, order=ORDER_BENCH) -> float:
        return float(
            self.lib.oracle_nbody_run_pool(pos, mass, vel, acc, pos.shape[0], ticks, threads, (C.c_int * 3)(*order))
        )

    def run_f64(self, pos, mass, vel, acc, ticks, order=ORDER_BENCH):
        self.lib.oracle_nbody_run_f64(pos, mass, vel, acc, pos.shape[0], ticks, (C.c_int * 3)(*order))

    # ---- rain --------------------------------------------------------------------------------
    def rain_gravity(self, vel):
        self.lib.oracle_rain_gravity(vel, 0, vel.shape[0])

    def rain_wind(self, vel, direction):
        self.lib.oracle_rain_wind(vel, 0, vel.shape[0], np.ascontiguousarray(direction, dtype=np.float32))

    def rain_movement(self, pos, vel):
        self.lib.oracle_rain_movement(pos, vel, 0, pos.shape[0])

    def rain_wind_direction(self, direction):
        d = np.ascontiguousarray(direction, dtype=np.float32).copy()
        self.lib.oracle_rain_wind_direction(d)
        return d


def load_native() -> NBodyOracle:
    """The oracle compiled on THIS machine with -march=native (the reference's .cargo/config.toml:2 asks for
    target-cpu=native), for the timed CPU baseline.  Same source, same -ffp-contract=off: the sequential f32 fold cannot be
    vectorised, so results are bit-identical to the portable build (checked here on a small sample before use)."""
    native = os.path.join(_HERE, "liboracle_native.so")
    src = os.path.join(_HERE, "nbody_oracle.c")
    subprocess.run(["make", "-B", "-C", _HERE, "native"], check=True, capture_output=True)  # always for THIS cpu
    if not os.path.exists(native) or os.path.getmtime(native) < os.path.getmtime(src):
        raise RuntimeError("native oracle build produced no library")
    orc = NBodyOracle(C.CDLL(native))
    rng = np.random.default_rng(0)
    pos = rng.uniform(-100, 100, size=(257, 3)).astype(np.float32)
    mass = rng.uniform(1e14, 2e14, size=257).astype(np.float32)
    if not np.array_equal(orc.gravity(pos, mass).view(np.uint32), load().gravity(pos, mass).view(np.uint32)):
        raise RuntimeError("native oracle build differs from the portable build")
    return orc


_cached = None


def load() -> NBodyOracle:
    global _cached
    if _cached is None:
        _cached = NBodyOracle(C.CDLL(build()))
    return _cached
