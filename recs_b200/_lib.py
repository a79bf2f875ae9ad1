"""ctypes binding of librecs_gpu.so (the C ABI declared in include/recs_gpu.h).

There is deliberately no fallback: if the shared library is missing or cannot be loaded the
import fails loudly, and every compute entry point returns RECS_ERR_NO_DEVICE without a GPU.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librecs_gpu.so")

RECS_OK = 0
ERR_NAMES = {
    0: "RECS_OK",
    1: "RECS_ERR_INVALID_ARGUMENT",
    2: "RECS_ERR_NO_DEVICE",
    3: "RECS_ERR_CUDA",
    4: "RECS_ERR_NCCL",
    5: "RECS_ERR_MISSING_PARAMETER",
    6: "RECS_ERR_STALE_BINDING",
    7: "RECS_ERR_SIZE_MISMATCH",
    8: "RECS_ERR_BORROW_CONFLICT",
    9: "RECS_ERR_SCHEDULE",
    10: "RECS_ERR_WORLD",
    11: "RECS_ERR_UNSUPPORTED",
    12: "RECS_ERR_COMPILE",
}
RECS_ERR_NO_DEVICE = 2

# recs_gpu_system_kind
SYS_NBODY_GRAVITY = 1
SYS_NBODY_MOVEMENT = 2
SYS_NBODY_ACCELERATION = 3
SYS_RAIN_GRAVITY = 4
SYS_RAIN_WIND = 5
SYS_RAIN_MOVEMENT = 6
SYS_RAIN_WIND_DIRECTION = 7
SYS_QUERY_ADD = 8
SYS_NBODY_COLLISION = 9
SYS_RAIN_GROUND_COLLISION = 10
SYS_KERNEL = 11

# recs_gpu_kernel_param.kind
KPARAM_READ = 0
KPARAM_WRITE = 1
KPARAM_ENTITY = 2
KPARAM_COMMANDS = 3
KPARAM_QUERY_READ = 4
KPARAM_QUERY_ENTITY = 5
ENTITY_COMPONENT = 0xFFFFFFFFFFFFFFFE

GPU_ASYNC = 0
GPU_SYNC = 1
NCCL_ID_BYTES = 128
IPC_HANDLE_BYTES = 64
MAX_ACCESSES = 8


class Config(C.Structure):
    _fields_ = [
        ("device", C.c_int32),
        ("rank", C.c_int32),
        ("world_size", C.c_int32),
        ("use_cuda_graph", C.c_int32),
        ("nbody_dt", C.c_float),
        ("nbody_g", C.c_float),
        ("nbody_epsilon", C.c_float),
        ("rain_dt", C.c_float),
        ("rain_gravity", C.c_float),
        ("rain_terminal", C.c_float),
        ("rain_wind_accel", C.c_float),
        ("rain_wind_deg_per_s", C.c_float),
        ("rain_ground_height", C.c_float),
        ("nbody_collision_radius", C.c_float),
    ]


class Access(C.Structure):
    _fields_ = [("component_id", C.c_uint64), ("is_write", C.c_uint8)]


class SystemDesc(C.Structure):
    _fields_ = [("name", C.c_char_p), ("n_access", C.c_uint32), ("access", Access * MAX_ACCESSES)]


class KernelParam(C.Structure):
    _fields_ = [("component_id", C.c_uint64), ("elem_size", C.c_uint32), ("kind", C.c_uint32), ("type_name", C.c_char_p)]


class Timing(C.Structure):
    _fields_ = [
        ("last_tick_chain_ms", C.c_float),
        ("gravity_ms", C.c_float),
        ("gravity_reduce_ms", C.c_float),
        ("movement_ms", C.c_float),
        ("acceleration_ms", C.c_float),
        ("pack_ms", C.c_float),
        ("other_ms", C.c_float),
        ("gather_ms", C.c_float),
        ("kernel_launches", C.c_uint64),
        ("gravity_launches", C.c_uint64),
    ]


# name -> (restype, argtypes); every symbol include/recs_gpu.h declares.
_ctx = C.c_void_p
GPU_SYMBOLS = {
    "recs_gpu_abi_version": (C.c_int32, []),
    "recs_gpu_default_config": (None, [C.POINTER(Config)]),
    "recs_gpu_create": (C.c_int32, [C.POINTER(Config), C.POINTER(_ctx)]),
    "recs_gpu_destroy": (C.c_int32, [_ctx]),
    "recs_gpu_last_error": (C.c_char_p, [_ctx]),
    "recs_gpu_clear_error": (C.c_int32, [_ctx]),
    "recs_gpu_synchronize": (C.c_int32, [_ctx]),
    "recs_gpu_bind_column": (C.c_int32, [_ctx, C.c_uint32, C.c_uint64, C.c_void_p, C.c_uint32, C.c_uint64]),
    "recs_gpu_upload": (C.c_int32, [_ctx, C.c_uint32, C.c_uint64]),
    "recs_gpu_download": (C.c_int32, [_ctx, C.c_uint32, C.c_uint64]),
    "recs_gpu_upload_rows": (C.c_int32, [_ctx, C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint64]),
    "recs_gpu_download_rows": (C.c_int32, [_ctx, C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint64]),
    "recs_gpu_column_append": (C.c_int32, [_ctx, C.c_uint32, C.c_uint64, C.c_void_p, C.c_uint64]),
    "recs_gpu_archetype_move_rows": (
        C.c_int32,
        [_ctx, C.c_uint32, C.POINTER(C.c_uint64), C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_uint32, C.c_uint64],
    ),
    "recs_gpu_column_rebind_host": (C.c_int32, [_ctx, C.c_uint32, C.c_uint64, C.c_void_p]),
    "recs_gpu_get_transfer_bytes": (C.c_int32, [_ctx, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "recs_gpu_invalidate": (C.c_int32, [_ctx, C.c_uint32]),
    "recs_gpu_column_device_ptr": (C.c_int32, [_ctx, C.c_uint32, C.c_uint64, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]),
    "recs_gpu_system_create": (C.c_int32, [_ctx, C.c_int32, C.POINTER(C.c_uint64), C.c_uint32, C.POINTER(C.c_uint32)]),
    "recs_gpu_system_create_kernel": (
        C.c_int32,
        [_ctx, C.c_char_p, C.c_char_p, C.c_char_p, C.POINTER(KernelParam), C.c_uint32, C.c_char_p, C.c_uint32, C.POINTER(C.c_uint32)],
    ),
    "recs_gpu_system_create_pair_kernel": (
        C.c_int32,
        [_ctx, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.POINTER(KernelParam), C.c_uint32, C.c_char_p,
         C.c_uint32, C.POINTER(C.c_uint32)],
    ),
    "recs_gpu_pair_kernel_system_compile": (
        C.c_int32,
        [C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.POINTER(KernelParam), C.c_uint32, C.c_char_p, C.c_char_p, C.c_uint32,
         C.POINTER(C.c_uint64)],
    ),
    "recs_gpu_system_set_uniform": (C.c_int32, [_ctx, C.c_uint32, C.c_void_p, C.c_uint32]),
    "recs_gpu_kernel_system_compile": (
        C.c_int32,
        [C.c_char_p, C.c_char_p, C.POINTER(KernelParam), C.c_uint32, C.c_char_p, C.c_int32, C.c_char_p, C.c_uint32,
         C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)],
    ),
    "recs_gpu_system_describe": (C.c_int32, [_ctx, C.c_uint32, C.POINTER(SystemDesc)]),
    "recs_gpu_system_set_archetypes": (C.c_int32, [_ctx, C.c_uint32, C.POINTER(C.c_uint32), C.c_uint32]),
    "recs_gpu_system_set_query_archetypes": (C.c_int32, [_ctx, C.c_uint32, C.POINTER(C.c_uint32), C.c_uint32]),
    "recs_gpu_system_removals": (C.c_int32, [_ctx, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32), C.c_uint32, C.POINTER(C.c_uint32)]),
    "recs_gpu_system_creates": (C.c_int32, [_ctx, C.c_uint32, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "recs_gpu_system_set_create_capacity": (C.c_int32, [_ctx, C.c_uint32, C.c_uint32]),
    "recs_gpu_run_system": (C.c_int32, [_ctx, C.c_uint32, C.c_uint32]),
    "recs_gpu_set_tick_order": (C.c_int32, [_ctx, C.POINTER(C.c_uint32), C.c_uint32]),
    "recs_gpu_tick": (C.c_int32, [_ctx, C.c_uint32]),
    "recs_gpu_set_wind_direction": (C.c_int32, [_ctx, C.POINTER(C.c_float)]),
    "recs_gpu_get_wind_direction": (C.c_int32, [_ctx, C.POINTER(C.c_float)]),
    "recs_gpu_comm_get_unique_id": (C.c_int32, [_ctx, C.POINTER(C.c_uint8)]),
    "recs_gpu_comm_init": (C.c_int32, [_ctx, C.POINTER(C.c_uint8)]),
    "recs_gpu_exchange_export": (C.c_int32, [_ctx, C.c_uint64, C.POINTER(C.c_uint8)]),
    "recs_gpu_exchange_import": (C.c_int32, [_ctx, C.POINTER(C.c_uint8), C.c_uint32]),
    "recs_gpu_exchange_mode": (C.c_int32, [_ctx]),
    "recs_gpu_column_allgather": (C.c_int32, [_ctx, C.c_uint32, C.c_uint64]),
    "recs_gpu_shard_rows": (C.c_int32, [_ctx, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "recs_gpu_shard_plan": (C.c_int32, [C.c_uint64, C.c_int32, C.c_int32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "recs_gpu_enable_timing": (C.c_int32, [_ctx, C.c_int32]),
    "recs_gpu_get_timing": (C.c_int32, [_ctx, C.POINTER(Timing)]),
    "recs_gpu_reset_timing": (C.c_int32, [_ctx]),
    "recs_gpu_timer_start": (C.c_int32, [_ctx]),
    "recs_gpu_timer_stop": (C.c_int32, [_ctx, C.POINTER(C.c_float)]),
    "recs_gpu_measure_fp32_peak": (C.c_int32, [_ctx, C.c_int32, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "recs_gpu_flush_l2": (C.c_int32, [_ctx]),
    "recs_gpu_host_alloc": (C.c_int32, [_ctx, C.c_size_t, C.POINTER(C.c_void_p)]),
    "recs_gpu_host_free": (C.c_int32, [_ctx, C.c_void_p]),
}


class RecsLibraryMissing(ImportError):
    pass


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise RecsLibraryMissing(
            f"{LIB_PATH} is missing: build it with `make` (or `python -c 'import __graft_entry__ as g; g.build()'`). "
            "recs_b200 has no CPU or PyTorch fallback for its CUDA path."
        )
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    return lib


lib = _load()


def _bind(symbols: dict) -> None:
    for name, (restype, argtypes) in symbols.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export it
        fn.restype = restype
        fn.argtypes = argtypes


_bind(GPU_SYMBOLS)
