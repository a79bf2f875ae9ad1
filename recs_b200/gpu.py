"""Thin object wrapper over the C ABI (include/recs_gpu.h): one `GpuContext` per device.

Only plumbing lives here (ctypes marshalling, keeping bound numpy buffers alive); all compute is
in librecs_gpu.so.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, Sequence

import numpy as np

from . import _lib
from ._lib import lib


class RecsGpuError(RuntimeError):
    def __init__(self, code: int, where: str, detail: str = ""):
        self.code = code
        name = _lib.ERR_NAMES.get(code, str(code))
        super().__init__(f"{where}: {name}{': ' + detail if detail else ''}")


def _u32s(values: Iterable[int]):
    values = list(values)
    arr = (C.c_uint32 * max(len(values), 1))(*values)
    return arr, len(values)


class GpuContext:
    """Owns a `recs_gpu_ctx`.  Mirrors the lifecycle of the reference's executor
    (`E::default()` ... `Drop`, crates/ecs/src/lib.rs:386-400)."""

    def __init__(self, device: int = 0, rank: int = 0, world_size: int = 1, use_cuda_graph: bool = True, **overrides):
        cfg = _lib.Config()
        lib.recs_gpu_default_config(C.byref(cfg))
        cfg.device = device
        cfg.rank = rank
        cfg.world_size = world_size
        cfg.use_cuda_graph = 1 if use_cuda_graph else 0
        for key, value in overrides.items():
            if not hasattr(cfg, key):
                raise TypeError(f"unknown recs_gpu_config field {key!r}")
            setattr(cfg, key, value)
        self.config = cfg
        self._ctx = C.c_void_p()
        self._bound = {}  # (arch, comp) -> ndarray kept alive while bound
        code = lib.recs_gpu_create(C.byref(cfg), C.byref(self._ctx))
        if code != _lib.RECS_OK:
            self._ctx = C.c_void_p()
            raise RecsGpuError(code, "recs_gpu_create", "no usable sm_100a device; recs_b200 has no CPU fallback")

    # -- lifecycle ---------------------------------------------------------------------------
    def close(self) -> None:
        if self._ctx:
            lib.recs_gpu_destroy(self._ctx)
            self._ctx = C.c_void_p()
            self._bound.clear()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, code: int, where: str) -> None:
        if code != _lib.RECS_OK:
            detail = lib.recs_gpu_last_error(self._ctx)
            raise RecsGpuError(code, where, detail.decode() if detail else "")

    def clear_error(self) -> None:
        lib.recs_gpu_clear_error(self._ctx)

    def last_error(self) -> str:
        return lib.recs_gpu_last_error(self._ctx).decode()

    def synchronize(self) -> None:
        self._check(lib.recs_gpu_synchronize(self._ctx), "recs_gpu_synchronize")

    # -- columns -----------------------------------------------------------------------------
    def bind_column(self, archetype: int, component: int, array: np.ndarray, elem_size: int | None = None) -> None:
        if not array.flags["C_CONTIGUOUS"]:
            raise ValueError("component columns must be C-contiguous")
        if elem_size is None:
            elem_size = array.dtype.itemsize * (int(np.prod(array.shape[1:])) if array.ndim > 1 else 1)
        count = array.shape[0] if array.ndim else 0
        self._check(
            lib.recs_gpu_bind_column(self._ctx, archetype, component, array.ctypes.data_as(C.c_void_p), elem_size, count),
            "recs_gpu_bind_column",
        )
        self._bound[(archetype, component)] = array

    def upload(self, archetype: int, component: int) -> None:
        self._check(lib.recs_gpu_upload(self._ctx, archetype, component), "recs_gpu_upload")

    def upload_rows(self, archetype: int, component: int, row_begin: int, row_end: int) -> None:
        self._check(lib.recs_gpu_upload_rows(self._ctx, archetype, component, row_begin, row_end), "recs_gpu_upload_rows")

    def download_rows(self, archetype: int, component: int, row_begin: int, row_end: int) -> None:
        self._check(lib.recs_gpu_download_rows(self._ctx, archetype, component, row_begin, row_end), "recs_gpu_download_rows")

    def download(self, archetype: int, component: int) -> np.ndarray:
        self._check(lib.recs_gpu_download(self._ctx, archetype, component), "recs_gpu_download")
        return self._bound[(archetype, component)]

    def invalidate(self, archetype: int) -> None:
        self._check(lib.recs_gpu_invalidate(self._ctx, archetype), "recs_gpu_invalidate")

    def column_device_ptr(self, archetype: int, component: int) -> tuple[int, int]:
        ptr = C.c_void_p()
        count = C.c_uint64()
        self._check(
            lib.recs_gpu_column_device_ptr(self._ctx, archetype, component, C.byref(ptr), C.byref(count)),
            "recs_gpu_column_device_ptr",
        )
        return int(ptr.value or 0), int(count.value)

    # -- systems -----------------------------------------------------------------------------
    def create_system(self, kind: int, components: Sequence[int]) -> int:
        ids = (C.c_uint64 * max(len(components), 1))(*components)
        out = C.c_uint32()
        self._check(lib.recs_gpu_system_create(self._ctx, kind, ids, len(components), C.byref(out)), "recs_gpu_system_create")
        return int(out.value)

    def create_kernel_system(self, name: str, source: str, body: str, params, uniform_type: str | None = None,
                             uniform_bytes: int = 0) -> int:
        """Registers a generic system (recs_gpu_system_create_kernel).  `params` is a sequence of
        (component_id, elem_size, kind, type_name) with kind in {"read", "write", "entity"}."""
        arr = kernel_params(params)
        out = C.c_uint32()
        self._check(
            lib.recs_gpu_system_create_kernel(
                self._ctx, name.encode(), source.encode(), body.encode(), arr, len(arr),
                uniform_type.encode() if uniform_type else None, uniform_bytes, C.byref(out)),
            "recs_gpu_system_create_kernel",
        )
        return int(out.value)

    def create_pair_kernel_system(self, name: str, source: str, acc_type: str, pair_fn: str, finish_fn: str, params,
                                  uniform_type: str | None = None, uniform_bytes: int = 0) -> int:
        """A generic system with a nested Query (recs_gpu_system_create_pair_kernel); nested-query parameters use
        the kinds "query_read" / "query_entity"."""
        arr = kernel_params(params)
        out = C.c_uint32()
        self._check(
            lib.recs_gpu_system_create_pair_kernel(
                self._ctx, name.encode(), source.encode(), acc_type.encode(), pair_fn.encode(), finish_fn.encode(), arr,
                len(arr), uniform_type.encode() if uniform_type else None, uniform_bytes, C.byref(out)),
            "recs_gpu_system_create_pair_kernel",
        )
        return int(out.value)

    def set_system_uniform(self, system: int, data: bytes) -> None:
        buf = C.create_string_buffer(bytes(data), len(data))
        self._check(lib.recs_gpu_system_set_uniform(self._ctx, system, buf, len(data)), "recs_gpu_system_set_uniform")

    def describe_system(self, system: int) -> tuple[str, list[tuple[int, bool]]]:
        desc = _lib.SystemDesc()
        self._check(lib.recs_gpu_system_describe(self._ctx, system, C.byref(desc)), "recs_gpu_system_describe")
        accesses = [(int(desc.access[i].component_id), bool(desc.access[i].is_write)) for i in range(desc.n_access)]
        return desc.name.decode(), accesses

    def set_archetypes(self, system: int, archetypes: Iterable[int]) -> None:
        arr, n = _u32s(archetypes)
        self._check(lib.recs_gpu_system_set_archetypes(self._ctx, system, arr, n), "recs_gpu_system_set_archetypes")

    def set_query_archetypes(self, system: int, archetypes: Iterable[int]) -> None:
        arr, n = _u32s(archetypes)
        self._check(
            lib.recs_gpu_system_set_query_archetypes(self._ctx, system, arr, n), "recs_gpu_system_set_query_archetypes"
        )

    def system_removals(self, system: int, archetype: int) -> np.ndarray:
        """Rows (ascending component indices) the system's last run asked to remove."""
        n = C.c_uint32()
        self._check(lib.recs_gpu_system_removals(self._ctx, system, archetype, None, 0, C.byref(n)), "recs_gpu_system_removals")
        out = (C.c_uint32 * max(n.value, 1))()
        self._check(lib.recs_gpu_system_removals(self._ctx, system, archetype, out, n.value, C.byref(n)), "recs_gpu_system_removals")
        return np.array(out[: n.value], dtype=np.uint32)

    def system_creates(self, system: int, dtype) -> np.ndarray:
        """Records the system's last run handed to `Commands::create`, in sequential-run order, as a structured array."""
        n, size = C.c_uint32(), C.c_uint32()
        self._check(lib.recs_gpu_system_creates(self._ctx, system, None, 0, C.byref(n), C.byref(size)), "recs_gpu_system_creates")
        dtype = np.dtype(dtype)
        assert n.value == 0 or dtype.itemsize == size.value, (dtype.itemsize, size.value)
        out = np.zeros(n.value, dtype=dtype)
        if n.value:
            self._check(lib.recs_gpu_system_creates(self._ctx, system, out.ctypes.data_as(C.c_void_p), n.value, C.byref(n), C.byref(size)),
                        "recs_gpu_system_creates")
        return out

    def set_create_capacity(self, system: int, max_records: int) -> None:
        self._check(lib.recs_gpu_system_set_create_capacity(self._ctx, system, max_records), "recs_gpu_system_set_create_capacity")

    def run_system(self, system: int, sync: bool = True) -> None:
        self._check(
            lib.recs_gpu_run_system(self._ctx, system, _lib.GPU_SYNC if sync else _lib.GPU_ASYNC), "recs_gpu_run_system"
        )

    def set_tick_order(self, systems: Iterable[int]) -> None:
        arr, n = _u32s(systems)
        self._check(lib.recs_gpu_set_tick_order(self._ctx, arr, n), "recs_gpu_set_tick_order")

    def tick(self, n_ticks: int = 1) -> None:
        self._check(lib.recs_gpu_tick(self._ctx, n_ticks), "recs_gpu_tick")

    # -- rain --------------------------------------------------------------------------------
    def set_wind_direction(self, xyz: Sequence[float]) -> None:
        arr = (C.c_float * 3)(*xyz)
        self._check(lib.recs_gpu_set_wind_direction(self._ctx, arr), "recs_gpu_set_wind_direction")

    def get_wind_direction(self) -> np.ndarray:
        arr = (C.c_float * 3)()
        self._check(lib.recs_gpu_get_wind_direction(self._ctx, arr), "recs_gpu_get_wind_direction")
        return np.array(list(arr), dtype=np.float32)

    # -- sharding ----------------------------------------------------------------------------
    def comm_unique_id(self) -> bytes:
        buf = (C.c_uint8 * _lib.NCCL_ID_BYTES)()
        self._check(lib.recs_gpu_comm_get_unique_id(self._ctx, buf), "recs_gpu_comm_get_unique_id")
        return bytes(buf)

    def comm_init(self, unique_id: bytes) -> None:
        buf = (C.c_uint8 * _lib.NCCL_ID_BYTES)(*unique_id)
        self._check(lib.recs_gpu_comm_init(self._ctx, buf), "recs_gpu_comm_init")

    def exchange_export(self, max_bodies: int) -> bytes:
        """Allocates the peer-push exchange buffers; returns this rank's CUDA IPC handle (recs_gpu_exchange_export)."""
        buf = (C.c_uint8 * _lib.IPC_HANDLE_BYTES)()
        self._check(lib.recs_gpu_exchange_export(self._ctx, max_bodies, buf), "recs_gpu_exchange_export")
        return bytes(buf)

    def exchange_import(self, handles: Sequence[bytes]) -> None:
        """Maps every rank's exchange buffers; `handles` in rank order (recs_gpu_exchange_import)."""
        blob = b"".join(handles)
        buf = (C.c_uint8 * len(blob))(*blob)
        self._check(lib.recs_gpu_exchange_import(self._ctx, buf, len(handles)), "recs_gpu_exchange_import")

    def exchange_mode(self) -> str:
        return {0: "single", 1: "nccl-allgather", 2: "peer-push"}[lib.recs_gpu_exchange_mode(self._ctx)]

    def column_allgather(self, archetype: int, component: int) -> None:
        """Every rank's own rows of the column to all ranks (recs_gpu_column_allgather)."""
        self._check(lib.recs_gpu_column_allgather(self._ctx, archetype, component), "recs_gpu_column_allgather")

    def shard_rows(self, n_rows: int) -> tuple[int, int]:
        b, e = C.c_uint64(), C.c_uint64()
        self._check(lib.recs_gpu_shard_rows(self._ctx, n_rows, C.byref(b), C.byref(e)), "recs_gpu_shard_rows")
        return int(b.value), int(e.value)

    # -- measurement -------------------------------------------------------------------------
    def enable_timing(self, enabled: bool = True) -> None:
        self._check(lib.recs_gpu_enable_timing(self._ctx, 1 if enabled else 0), "recs_gpu_enable_timing")

    def timing(self) -> dict:
        t = _lib.Timing()
        self._check(lib.recs_gpu_get_timing(self._ctx, C.byref(t)), "recs_gpu_get_timing")
        return {name: getattr(t, name) for name, _ in _lib.Timing._fields_}

    def transfer_bytes(self) -> tuple[int, int]:
        """Component-column bytes copied (host->device, device->host) since the context was created."""
        h2d, d2h = C.c_uint64(), C.c_uint64()
        self._check(lib.recs_gpu_get_transfer_bytes(self._ctx, C.byref(h2d), C.byref(d2h)), "recs_gpu_get_transfer_bytes")
        return int(h2d.value), int(d2h.value)

    def reset_timing(self) -> None:
        self._check(lib.recs_gpu_reset_timing(self._ctx), "recs_gpu_reset_timing")

    def timer_start(self) -> None:
        self._check(lib.recs_gpu_timer_start(self._ctx), "recs_gpu_timer_start")

    def timer_stop(self) -> float:
        ms = C.c_float()
        self._check(lib.recs_gpu_timer_stop(self._ctx, C.byref(ms)), "recs_gpu_timer_stop")
        return float(ms.value)

    def measure_fp32_peak(self, packed: bool = True) -> tuple[float, float]:
        tf, mhz = C.c_float(), C.c_float()
        self._check(
            lib.recs_gpu_measure_fp32_peak(self._ctx, 1 if packed else 0, C.byref(tf), C.byref(mhz)),
            "recs_gpu_measure_fp32_peak",
        )
        return float(tf.value), float(mhz.value)

    def flush_l2(self) -> None:
        self._check(lib.recs_gpu_flush_l2(self._ctx), "recs_gpu_flush_l2")

    def pinned_empty(self, shape, dtype=np.float32) -> np.ndarray:
        """A numpy array over page-locked host memory (end-to-end path: H2D/D2H at full PCIe rate)."""
        dtype = np.dtype(dtype)
        nbytes = int(np.prod(shape)) * dtype.itemsize
        ptr = C.c_void_p()
        self._check(lib.recs_gpu_host_alloc(self._ctx, nbytes, C.byref(ptr)), "recs_gpu_host_alloc")
        buf = (C.c_uint8 * max(nbytes, 1)).from_address(ptr.value)
        arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
        return arr


_KPARAM_KINDS = {"read": _lib.KPARAM_READ, "write": _lib.KPARAM_WRITE, "entity": _lib.KPARAM_ENTITY,
                 "commands": _lib.KPARAM_COMMANDS, "query_read": _lib.KPARAM_QUERY_READ,
                 "query_entity": _lib.KPARAM_QUERY_ENTITY}


def kernel_params(params):
    """ctypes array of recs_gpu_kernel_param from (component_id, elem_size, kind, type_name) tuples."""
    arr = (_lib.KernelParam * max(len(params), 1))()
    for i, (comp, elem, kind, type_name) in enumerate(params):
        arr[i].component_id = comp
        arr[i].elem_size = elem
        arr[i].kind = _KPARAM_KINDS[kind]
        arr[i].type_name = type_name.encode() if type_name else None  # ("commands", size, record type): Commands::create record
    arr._n = len(params)
    return arr


def compile_kernel_system(source: str, body: str, params, uniform_type: str | None = None, force_direct: bool = False):
    """recs_gpu_kernel_system_compile: generation + NVRTC only (no GPU needed).
    Returns (status, log, cubin_bytes, staged_chunk)."""
    arr = kernel_params(params)
    log = C.create_string_buffer(16384)
    size = C.c_uint64()
    chunk = C.c_uint32()
    st = lib.recs_gpu_kernel_system_compile(
        source.encode(), body.encode(), arr, len(params), uniform_type.encode() if uniform_type else None,
        1 if force_direct else 0, log, len(log), C.byref(size), C.byref(chunk))
    return int(st), log.value.decode(errors="replace"), int(size.value), int(chunk.value)


def compile_pair_kernel_system(source: str, acc_type: str, pair_fn: str, finish_fn: str, params, uniform_type: str | None = None):
    """recs_gpu_pair_kernel_system_compile: (status, log, cubin_bytes); needs no GPU."""
    arr = kernel_params(params)
    log = C.create_string_buffer(16384)
    size = C.c_uint64()
    st = lib.recs_gpu_pair_kernel_system_compile(
        source.encode(), acc_type.encode(), pair_fn.encode(), finish_fn.encode(), arr, len(params),
        uniform_type.encode() if uniform_type else None, log, len(log), C.byref(size))
    return int(st), log.value.decode(errors="replace"), int(size.value)
