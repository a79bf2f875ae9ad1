"""ctypes view of the host-side ECS layer of librecs_gpu.so (include/recs_ecs.h): World, system
parameters, segments, PrecedenceGraph and the Application run loop.  Names follow the reference
(crates/ecs, crates/scheduler); all logic is in C++ -- this file only marshals."""
from __future__ import annotations

import ctypes as C
from typing import Iterable, Sequence

import numpy as np

from . import _lib
from ._lib import Access, lib
from .gpu import GpuContext, RecsGpuError

# recs_param_op
READ, WRITE, ENTITY, COMMANDS, WITH, NOT, AND, OR, XOR, QUERY_BEGIN, QUERY_END = range(1, 12)
SCHEDULE_NEW_TICK, SCHEDULE_WOULD_BLOCK = 100, 101
REACTION_RETURN_ERROR, REACTION_RETURN_NEW_TICK = 0, 1
SELECT_CROSSBEAM, SELECT_FIRST = 0, 1


class Entity(C.Structure):
    _fields_ = [("id", C.c_uint32), ("generation", C.c_uint32)]

    def key(self):
        return (int(self.id), int(self.generation))

    def __repr__(self):
        return f"Entity(id={self.id}, generation={self.generation})"


class ParamToken(C.Structure):
    _fields_ = [("op", C.c_uint32), ("component_id", C.c_uint64)]


class SegmentSlice(C.Structure):
    _fields_ = [("segment", C.c_uint32), ("column", C.c_uint32), ("begin", C.c_uint64), ("len", C.c_uint64)]


class SystemInfo(C.Structure):
    _fields_ = [("name", C.c_char_p), ("n_access", C.c_uint32), ("access", C.POINTER(Access))]


CPU_SYSTEM_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_uint32, C.c_uint64, C.POINTER(C.c_void_p), C.c_uint32)
CPU_SEGMENT_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_uint32, C.c_uint64, C.c_uint64, C.POINTER(C.c_void_p), C.c_uint32)

_w, _s, _a = C.c_void_p, C.c_void_p, C.c_void_p
_u32p, _u64p = C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)
_vpp = C.POINTER(C.c_void_p)
ECS_SYMBOLS = {
    "recs_world_create": (C.c_int32, [C.POINTER(_w)]),
    "recs_world_destroy": (C.c_int32, [_w]),
    "recs_world_last_error": (C.c_char_p, [_w]),
    "recs_world_register_component": (C.c_int32, [_w, C.c_uint64, C.c_uint32, C.c_char_p]),
    "recs_world_create_empty_entity": (C.c_int32, [_w, C.POINTER(Entity)]),
    "recs_world_create_entity": (C.c_int32, [_w, _u64p, _vpp, C.c_uint32, C.POINTER(Entity)]),
    "recs_world_create_entities": (C.c_int32, [_w, C.c_uint64, _u64p, _vpp, C.c_uint32, C.POINTER(Entity)]),
    "recs_world_delete_entity": (C.c_int32, [_w, Entity]),
    "recs_world_add_components": (C.c_int32, [_w, Entity, _u64p, _vpp, C.c_uint32]),
    "recs_world_remove_components": (C.c_int32, [_w, Entity, _u64p, C.c_uint32]),
    "recs_world_archetype_count": (C.c_int32, [_w, _u32p]),
    "recs_world_entity_count": (C.c_int32, [_w, _u64p, _u64p]),
    "recs_world_entity_at": (C.c_int32, [_w, C.c_uint32, C.POINTER(Entity), C.POINTER(C.c_int32)]),
    "recs_world_entity_archetype": (C.c_int32, [_w, Entity, _u32p]),
    "recs_world_entity_component_index": (C.c_int32, [_w, Entity, _u64p]),
    "recs_world_archetype_entities": (C.c_int32, [_w, C.c_uint32, C.POINTER(Entity), C.c_uint64, _u64p]),
    "recs_world_archetype_components": (C.c_int32, [_w, C.c_uint32, _u64p, C.c_uint32, _u32p]),
    "recs_world_column": (C.c_int32, [_w, C.c_uint32, C.c_uint64, _vpp, _u32p, _u64p]),
    "recs_world_borrow_column": (C.c_int32, [_w, C.c_uint32, C.c_uint64, C.c_int32]),
    "recs_world_release_column": (C.c_int32, [_w, C.c_uint32, C.c_uint64, C.c_int32]),
    "recs_world_get_archetype_indices": (C.c_int32, [_w, _u64p, C.c_uint32, _u32p, C.c_uint32, _u32p]),
    "recs_world_archetype_version": (C.c_int32, [_w, C.c_uint32, _u64p]),
    "recs_params_component_accesses": (C.c_int32, [C.POINTER(ParamToken), C.c_uint32, C.POINTER(Access), C.c_uint32, _u32p]),
    "recs_params_supports_parallelization": (C.c_int32, [C.POINTER(ParamToken), C.c_uint32, C.POINTER(C.c_int32)]),
    "recs_params_controls_iteration": (C.c_int32, [C.POINTER(ParamToken), C.c_uint32, C.POINTER(C.c_int32)]),
    "recs_params_get_archetype_indices": (C.c_int32, [_w, C.POINTER(ParamToken), C.c_uint32, C.c_int32, _u32p, C.c_uint32, _u32p]),
    "recs_auto_segment_size": (C.c_uint64, [C.c_uint64, C.c_uint32]),
    "recs_segment_count": (C.c_uint64, [C.c_uint64, C.c_uint64]),
    "recs_split_segments": (C.c_int32, [_u64p, C.c_uint32, C.c_uint64, C.POINTER(SegmentSlice), C.c_uint64, _u64p, _u64p]),
    "recs_precedence": (C.c_int32, [C.POINTER(SystemInfo), C.POINTER(SystemInfo)]),
    "recs_schedule_generate": (C.c_int32, [C.POINTER(SystemInfo), C.c_uint32, C.c_int32, C.POINTER(_s)]),
    "recs_schedule_destroy": (C.c_int32, [_s]),
    "recs_schedule_edges": (C.c_int32, [_s, _u32p, _u32p, C.c_uint32, _u32p]),
    "recs_schedule_next_batch": (C.c_int32, [_s, C.c_int32, _u32p, C.c_uint32, _u32p]),
    "recs_schedule_complete": (C.c_int32, [_s, C.c_uint32]),
    "recs_schedule_set_select_policy": (C.c_int32, [_s, C.c_int32]),
    "recs_schedule_tick_order": (C.c_int32, [_s, _u32p, _u32p, C.c_uint32, _u32p]),
    "recs_app_create": (C.c_int32, [C.c_void_p, C.POINTER(_a)]),
    "recs_app_destroy": (C.c_int32, [_a]),
    "recs_app_last_error": (C.c_char_p, [_a]),
    "recs_app_world": (_w, [_a]),
    "recs_app_add_gpu_system": (C.c_int32, [_a, C.c_int32, _u64p, C.c_uint32, _u32p]),
    "recs_app_add_gpu_kernel_system": (
        C.c_int32,
        [_a, C.c_char_p, C.c_char_p, C.c_char_p, C.POINTER(ParamToken), C.c_uint32, C.POINTER(C.c_char_p), C.c_char_p, C.c_uint32, _u32p],
    ),
    "recs_app_add_gpu_kernel_system_creating": (
        C.c_int32,
        [_a, C.c_char_p, C.c_char_p, C.c_char_p, C.POINTER(ParamToken), C.c_uint32, C.POINTER(C.c_char_p), C.c_char_p, C.c_uint32,
         _u64p, C.c_uint32, C.c_char_p, _u32p],
    ),
    "recs_app_set_system_uniform": (C.c_int32, [_a, C.c_uint32, C.c_void_p, C.c_uint32]),
    "recs_app_add_gpu_pair_kernel_system": (
        C.c_int32,
        [_a, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.POINTER(ParamToken), C.c_uint32, C.POINTER(C.c_char_p),
         C.c_char_p, C.c_uint32, _u32p],
    ),
    "recs_app_add_cpu_segment_system": (C.c_int32, [_a, C.c_char_p, C.POINTER(ParamToken), C.c_uint32, CPU_SEGMENT_FN, C.c_void_p, _u32p]),
    "recs_app_set_executor": (C.c_int32, [_a, C.c_uint32]),
    "recs_fold_swap_removes": (C.c_int32, [C.c_uint64, _u32p, C.c_uint32, _u32p, _u32p, C.c_uint32, _u32p, C.POINTER(C.c_uint64)]),
    "recs_app_add_cpu_system": (C.c_int32, [_a, C.c_char_p, C.POINTER(ParamToken), C.c_uint32, CPU_SYSTEM_FN, C.c_void_p, _u32p]),
    "recs_app_run": (C.c_int32, [_a, C.c_uint32, C.c_int32]),
    "recs_app_command_create": (C.c_int32, [_a, _u64p, _vpp, C.c_uint32]),
    "recs_app_command_remove": (C.c_int32, [_a, Entity]),
    "recs_app_last_playback": (C.c_int32, [_a, _u64p, _u64p]),
    "recs_app_last_tick_order": (C.c_int32, [_a, _u32p, _u32p, C.c_uint32, _u32p]),
    "recs_app_sync_to_host": (C.c_int32, [_a]),
    "recs_app_add_components": (C.c_int32, [_a, Entity, _u64p, _vpp, C.c_uint32]),
    "recs_app_remove_components": (C.c_int32, [_a, Entity, _u64p, C.c_uint32]),
    "recs_app_mark_host_dirty": (C.c_int32, [_a, C.c_uint32]),
}
_lib._bind(ECS_SYMBOLS)


class EcsError(RuntimeError):
    def __init__(self, code, where, detail=""):
        self.code = code
        super().__init__(f"{where}: {_lib.ERR_NAMES.get(code, code)}{': ' + detail if detail else ''}")


def tokens(params) -> tuple:
    """Nested parameter description -> prefix token array.  A parameter is ("read", c),
    ("write", c), ("entity",), ("commands",), ("with", c), ("not", f), ("and"|"or"|"xor", f, g) or
    ("query", [params...])."""
    out = []

    def emit(p):
        kind = p[0]
        if kind == "read":
            out.append((READ, p[1]))
        elif kind == "write":
            out.append((WRITE, p[1]))
        elif kind == "entity":
            out.append((ENTITY, 0))
        elif kind == "commands":
            out.append((COMMANDS, 0))
        elif kind == "with":
            out.append((WITH, p[1]))
        elif kind == "not":
            out.append((NOT, 0))
            emit(p[1])
        elif kind in ("and", "or", "xor"):
            out.append(({"and": AND, "or": OR, "xor": XOR}[kind], 0))
            emit(p[1])
            emit(p[2])
        elif kind == "query":
            out.append((QUERY_BEGIN, 0))
            for q in p[1]:
                emit(q)
            out.append((QUERY_END, 0))
        else:
            raise ValueError(f"unknown parameter {p!r}")

    for p in params:
        emit(p)
    arr = (ParamToken * max(len(out), 1))()
    for i, (op, comp) in enumerate(out):
        arr[i].op, arr[i].component_id = op, comp
    return arr, len(out)


def component_accesses(params) -> list:
    arr, n = tokens(params)
    out = (Access * 64)()
    cnt = C.c_uint32()
    code = lib.recs_params_component_accesses(arr, n, out, 64, C.byref(cnt))
    if code:
        raise EcsError(code, "recs_params_component_accesses")
    return [(int(out[i].component_id), bool(out[i].is_write)) for i in range(cnt.value)]


def supports_parallelization(params) -> bool:
    arr, n = tokens(params)
    b = C.c_int32()
    code = lib.recs_params_supports_parallelization(arr, n, C.byref(b))
    if code:
        raise EcsError(code, "recs_params_supports_parallelization")
    return bool(b.value)


def controls_iteration(params) -> bool:
    arr, n = tokens(params)
    b = C.c_int32()
    code = lib.recs_params_controls_iteration(arr, n, C.byref(b))
    if code:
        raise EcsError(code, "recs_params_controls_iteration")
    return bool(b.value)


def auto_segment_size(entity_count: int, thread_count: int) -> int:
    return int(lib.recs_auto_segment_size(entity_count, thread_count))


def segment_count(entity_count: int, segment_size: int) -> int:
    return int(lib.recs_segment_count(entity_count, segment_size))


def split_segments(column_lengths: Sequence[int], segment_size: int) -> list:
    n = len(column_lengths)
    lens = (C.c_uint64 * max(n, 1))(*column_lengths)
    n_slices, n_segments = C.c_uint64(), C.c_uint64()
    lib.recs_split_segments(lens, n, segment_size, None, 0, C.byref(n_slices), C.byref(n_segments))
    out = (SegmentSlice * max(n_slices.value, 1))()
    code = lib.recs_split_segments(lens, n, segment_size, out, n_slices.value, C.byref(n_slices), C.byref(n_segments))
    if code:
        raise EcsError(code, "recs_split_segments")
    segments = [[] for _ in range(n_segments.value)]
    for i in range(n_slices.value):
        s = out[i]
        segments[s.segment].append((int(s.column), int(s.begin), int(s.len)))
    return segments


class World:
    """`ecs::World` (crates/ecs/src/lib.rs:665-929)."""

    def __init__(self, handle=None):
        self._owned = handle is None
        self._h = _w()
        if handle is None:
            lib.recs_world_create(C.byref(self._h))
        else:
            self._h = _w(handle)
        self._dtypes = {}

    def close(self):
        if self._owned and self._h:
            lib.recs_world_destroy(self._h)
        self._h = _w()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, code, where):
        if code:
            raise EcsError(code, where, lib.recs_world_last_error(self._h).decode())

    def register_component(self, component: int, dtype, name: str | None = None) -> None:
        dtype = np.dtype(dtype) if dtype is not None else None
        size = dtype.itemsize if dtype is not None else 0
        self._dtypes[component] = dtype
        self._check(
            lib.recs_world_register_component(self._h, component, size, (name or f"component#{component}").encode()),
            "recs_world_register_component",
        )

    def _marshal(self, components: dict):
        ids = (C.c_uint64 * max(len(components), 1))(*components.keys())
        keep, ptrs = [], (C.c_void_p * max(len(components), 1))()
        for i, (comp, value) in enumerate(components.items()):
            dt = self._dtypes[comp]
            if dt is None:
                ptrs[i] = None
                continue
            arr = np.ascontiguousarray(np.asarray(value, dtype=dt.base if dt.subdtype else dt).reshape(-1))
            keep.append(arr)
            ptrs[i] = arr.ctypes.data
        return ids, ptrs, keep

    def create_empty_entity(self) -> Entity:
        e = Entity()
        self._check(lib.recs_world_create_empty_entity(self._h, C.byref(e)), "recs_world_create_empty_entity")
        return e

    def create_entity(self, components: dict) -> Entity:
        ids, ptrs, keep = self._marshal(components)
        e = Entity()
        self._check(lib.recs_world_create_entity(self._h, ids, ptrs, len(components), C.byref(e)), "recs_world_create_entity")
        return e

    def create_entities(self, count: int, columns: dict) -> None:
        ids = (C.c_uint64 * max(len(columns), 1))(*columns.keys())
        ptrs = (C.c_void_p * max(len(columns), 1))()
        keep = []
        for i, (comp, col) in enumerate(columns.items()):
            arr = np.ascontiguousarray(col)
            keep.append(arr)
            ptrs[i] = arr.ctypes.data
        self._check(lib.recs_world_create_entities(self._h, count, ids, ptrs, len(columns), None), "recs_world_create_entities")

    def delete_entity(self, e: Entity) -> None:
        self._check(lib.recs_world_delete_entity(self._h, e), "recs_world_delete_entity")

    def add_components(self, e: Entity, components: dict) -> None:
        ids, ptrs, keep = self._marshal(components)
        self._check(lib.recs_world_add_components(self._h, e, ids, ptrs, len(components)), "recs_world_add_components")

    def remove_components(self, e: Entity, components: Sequence[int]) -> None:
        ids = (C.c_uint64 * max(len(components), 1))(*components)
        self._check(lib.recs_world_remove_components(self._h, e, ids, len(components)), "recs_world_remove_components")

    def archetype_count(self) -> int:
        n = C.c_uint32()
        self._check(lib.recs_world_archetype_count(self._h, C.byref(n)), "recs_world_archetype_count")
        return int(n.value)

    def entity_slots(self) -> list:
        alive, slots = C.c_uint64(), C.c_uint64()
        lib.recs_world_entity_count(self._h, C.byref(alive), C.byref(slots))
        out = []
        for i in range(slots.value):
            e, a = Entity(), C.c_int32()
            lib.recs_world_entity_at(self._h, i, C.byref(e), C.byref(a))
            out.append(e.key() if a.value else None)
        return out

    def entity_archetype(self, e: Entity) -> int:
        a = C.c_uint32()
        self._check(lib.recs_world_entity_archetype(self._h, e, C.byref(a)), "recs_world_entity_archetype")
        return int(a.value)

    def entity_component_index(self, e: Entity) -> int:
        i = C.c_uint64()
        self._check(lib.recs_world_entity_component_index(self._h, e, C.byref(i)), "recs_world_entity_component_index")
        return int(i.value)

    def archetype_entities(self, archetype: int) -> list:
        n = C.c_uint64()
        self._check(lib.recs_world_archetype_entities(self._h, archetype, None, 0, C.byref(n)), "recs_world_archetype_entities")
        buf = (Entity * max(n.value, 1))()
        lib.recs_world_archetype_entities(self._h, archetype, buf, n.value, C.byref(n))
        return [buf[i].key() for i in range(n.value)]

    def archetype_components(self, archetype: int) -> list:
        n = C.c_uint32()
        buf = (C.c_uint64 * 64)()
        self._check(lib.recs_world_archetype_components(self._h, archetype, buf, 64, C.byref(n)), "recs_world_archetype_components")
        return [int(buf[i]) for i in range(n.value)]

    def column(self, archetype: int, component: int) -> np.ndarray:
        """A numpy VIEW of the archetype's column (invalidated by the next structural change)."""
        ptr, elem, count = C.c_void_p(), C.c_uint32(), C.c_uint64()
        self._check(lib.recs_world_column(self._h, archetype, component, C.byref(ptr), C.byref(elem), C.byref(count)), "recs_world_column")
        dt = self._dtypes.get(component)
        if dt is None or count.value == 0 or elem.value == 0:
            return np.zeros((count.value,) if dt is None or not dt.subdtype else (count.value,) + dt.shape, dtype=(dt.base if dt is not None and dt.subdtype else dt) or np.uint8)
        buf = (C.c_uint8 * (elem.value * count.value)).from_address(ptr.value)
        base = dt.base if dt.subdtype else dt
        arr = np.frombuffer(buf, dtype=base)
        return arr.reshape((count.value,) + dt.shape) if dt.subdtype else arr

    def borrow_column(self, archetype: int, component: int, write: bool) -> None:
        self._check(lib.recs_world_borrow_column(self._h, archetype, component, 1 if write else 0), "recs_world_borrow_column")

    def release_column(self, archetype: int, component: int, write: bool) -> None:
        self._check(lib.recs_world_release_column(self._h, archetype, component, 1 if write else 0), "recs_world_release_column")

    def get_archetype_indices(self, signature: Sequence[int]) -> list:
        sig = (C.c_uint64 * max(len(signature), 1))(*signature)
        out = (C.c_uint32 * 4096)()
        n = C.c_uint32()
        self._check(lib.recs_world_get_archetype_indices(self._h, sig, len(signature), out, 4096, C.byref(n)), "recs_world_get_archetype_indices")
        return [int(out[i]) for i in range(n.value)]

    def query_archetypes(self, params, query_index: int = -1) -> list:
        arr, n = tokens(params)
        out = (C.c_uint32 * 4096)()
        cnt = C.c_uint32()
        self._check(lib.recs_params_get_archetype_indices(self._h, arr, n, query_index, out, 4096, C.byref(cnt)), "recs_params_get_archetype_indices")
        return [int(out[i]) for i in range(cnt.value)]

    def archetype_version(self, archetype: int) -> int:
        v = C.c_uint64()
        self._check(lib.recs_world_archetype_version(self._h, archetype, C.byref(v)), "recs_world_archetype_version")
        return int(v.value)


def _infos(systems):
    """systems: [(name, [(component, is_write), ...]), ...] -> (SystemInfo array, keep-alive)."""
    arr = (SystemInfo * max(len(systems), 1))()
    keep = []
    for i, (name, accesses) in enumerate(systems):
        acc = (Access * max(len(accesses), 1))()
        for j, (comp, w) in enumerate(accesses):
            acc[j].component_id, acc[j].is_write = comp, 1 if w else 0
        nm = name.encode()
        keep.append((acc, nm))
        arr[i].name, arr[i].n_access, arr[i].access = nm, len(accesses), acc
    return arr, keep


def precedence(a, b) -> int:
    """Orderable::precedence_to: -1 Before, 0 Equal, +1 After."""
    arr, keep = _infos([a, b])
    return int(lib.recs_precedence(C.byref(arr[0]), C.byref(arr[1])))


class PrecedenceGraph:
    """`scheduler::schedule::PrecedenceGraph` (ordered=True: OrderedPrecedenceGraph)."""

    def __init__(self, systems, ordered: bool = False, select: int = SELECT_CROSSBEAM):
        self.systems = list(systems)
        arr, keep = _infos(self.systems)
        self._h = _s()
        code = lib.recs_schedule_generate(arr, len(self.systems), 1 if ordered else 0, C.byref(self._h))
        if code:
            raise EcsError(code, "recs_schedule_generate")
        lib.recs_schedule_set_select_policy(self._h, select)

    def __del__(self):
        try:
            if self._h:
                lib.recs_schedule_destroy(self._h)
        except Exception:
            pass

    def edges(self) -> list:
        n = C.c_uint32()
        lib.recs_schedule_edges(self._h, None, None, 0, C.byref(n))
        f, t = (C.c_uint32 * max(n.value, 1))(), (C.c_uint32 * max(n.value, 1))()
        lib.recs_schedule_edges(self._h, f, t, n.value, C.byref(n))
        return [(int(f[i]), int(t[i])) for i in range(n.value)]

    def next_batch(self, reaction: int = REACTION_RETURN_ERROR):
        """Returns a list of system indices, or the strings "new_tick" / "would_block"."""
        out = (C.c_uint32 * (len(self.systems) + 1))()
        n = C.c_uint32()
        code = lib.recs_schedule_next_batch(self._h, reaction, out, len(self.systems) + 1, C.byref(n))
        if code == SCHEDULE_NEW_TICK:
            return "new_tick"
        if code == SCHEDULE_WOULD_BLOCK:
            return "would_block"
        if code:
            raise EcsError(code, "recs_schedule_next_batch")
        return [int(out[i]) for i in range(n.value)]

    def currently_executable_systems(self):
        return self.next_batch(REACTION_RETURN_NEW_TICK)

    def complete(self, system: int) -> None:
        code = lib.recs_schedule_complete(self._h, system)
        if code:
            raise EcsError(code, "recs_schedule_complete")

    def tick_order(self):
        cap = len(self.systems) + 1
        order, batch = (C.c_uint32 * cap)(), (C.c_uint32 * cap)()
        n = C.c_uint32()
        code = lib.recs_schedule_tick_order(self._h, order, batch, cap, C.byref(n))
        if code:
            raise EcsError(code, "recs_schedule_tick_order")
        return [int(order[i]) for i in range(n.value)], [int(batch[i]) for i in range(n.value)]


class Application:
    """`BasicApplication` (crates/ecs/src/lib.rs:240-341) whose executor dispatches GPU systems on
    the context stream.  `gpu=None` gives a CPU-systems-only application."""

    def __init__(self, gpu: GpuContext | None = None):
        self.gpu = gpu
        self._h = _a()
        code = lib.recs_app_create(gpu._ctx if gpu is not None else None, C.byref(self._h))
        if code:
            raise EcsError(code, "recs_app_create")
        self.world = World(handle=lib.recs_app_world(self._h))
        self.system_names = []
        self._callbacks = []

    def close(self):
        if self._h:
            lib.recs_app_destroy(self._h)
            self._h = _a()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, code, where):
        if code:
            raise EcsError(code, where, lib.recs_app_last_error(self._h).decode())

    def add_gpu_system(self, kind: int, components: Sequence[int], name: str | None = None) -> int:
        ids = (C.c_uint64 * max(len(components), 1))(*components)
        idx = C.c_uint32()
        self._check(lib.recs_app_add_gpu_system(self._h, kind, ids, len(components), C.byref(idx)), "recs_app_add_gpu_system")
        self.system_names.append(name or f"gpu:{kind}")
        return int(idx.value)

    def add_gpu_kernel_system(self, name: str, source: str, body: str, params, type_names: Sequence[str],
                              uniform_type: str | None = None, uniform_bytes: int = 0, creates: Sequence[int] = (),
                              create_record_type: str | None = None) -> int:
        """A generic GPU system: `params` is the function's parameter list (same grammar as CPU systems, filters
        included); `type_names` are the C++ types of its Read / Write / Entity arguments in `source`, in order.
        `creates` / `create_record_type`: the components of the tuple its Commands parameter hands to create(), and the POD
        struct in `source` holding them in that order (recs_app_add_gpu_kernel_system_creating)."""
        arr, n = tokens(params)
        names = (C.c_char_p * max(len(type_names), 1))(*[t.encode() for t in type_names])
        idx = C.c_uint32()
        comps = (C.c_uint64 * max(len(creates), 1))(*creates)
        self._check(
            lib.recs_app_add_gpu_kernel_system_creating(
                self._h, name.encode(), source.encode(), body.encode(), arr, n, names,
                uniform_type.encode() if uniform_type else None, uniform_bytes, comps, len(creates),
                create_record_type.encode() if create_record_type else None, C.byref(idx)),
            "recs_app_add_gpu_kernel_system_creating",
        )
        self.system_names.append(name)
        return int(idx.value)

    def add_gpu_pair_kernel_system(self, name: str, source: str, acc_type: str, pair_fn: str, finish_fn: str, params,
                                   type_names: Sequence[str], uniform_type: str | None = None, uniform_bytes: int = 0) -> int:
        """A generic GPU system with a nested Query; type_names: outer parameters first, then the query's, in token order."""
        arr, n = tokens(params)
        names = (C.c_char_p * max(len(type_names), 1))(*[t.encode() for t in type_names])
        idx = C.c_uint32()
        self._check(
            lib.recs_app_add_gpu_pair_kernel_system(
                self._h, name.encode(), source.encode(), acc_type.encode(), pair_fn.encode(), finish_fn.encode(), arr, n,
                names, uniform_type.encode() if uniform_type else None, uniform_bytes, C.byref(idx)),
            "recs_app_add_gpu_pair_kernel_system",
        )
        self.system_names.append(name)
        return int(idx.value)

    def set_system_uniform(self, system: int, data: bytes) -> None:
        buf = C.create_string_buffer(bytes(data), len(data))
        self._check(lib.recs_app_set_system_uniform(self._h, system, buf, len(data)), "recs_app_set_system_uniform")

    def add_cpu_system(self, name: str, params, fn) -> int:
        """fn(archetype, count, columns) with columns = list of raw host pointers in parameter order."""

        def trampoline(_user, archetype, count, cols, n_cols):
            fn(int(archetype), int(count), [cols[i] for i in range(n_cols)])

        cb = CPU_SYSTEM_FN(trampoline)
        self._callbacks.append(cb)
        arr, n = tokens(params)
        idx = C.c_uint32()
        self._check(lib.recs_app_add_cpu_system(self._h, name.encode(), arr, n, cb, None, C.byref(idx)), "recs_app_add_cpu_system")
        self.system_names.append(name)
        return int(idx.value)

    def add_cpu_segment_system(self, name: str, params, fn) -> int:
        """Par-iter CPU system: fn(archetype, row_begin, row_count, columns) with columns = raw host BASE pointers
        of the archetype's columns in parameter order; may be called from worker threads."""

        def trampoline(_user, archetype, row_begin, row_count, cols, n_cols):
            fn(int(archetype), int(row_begin), int(row_count), [cols[i] for i in range(n_cols)])

        cb = CPU_SEGMENT_FN(trampoline)
        self._callbacks.append(cb)
        arr, n = tokens(params)
        idx = C.c_uint32()
        self._check(lib.recs_app_add_cpu_segment_system(self._h, name.encode(), arr, n, cb, None, C.byref(idx)),
                    "recs_app_add_cpu_segment_system")
        self.system_names.append(name)
        return int(idx.value)

    def set_executor(self, worker_threads: int) -> None:
        """0 = Sequential executor, n > 0 = WorkerPool with n worker threads."""
        self._check(lib.recs_app_set_executor(self._h, worker_threads), "recs_app_set_executor")

    def command_create(self, components: dict) -> None:
        """`commands.create(...)` from inside a CPU system."""
        ids, ptrs, keep = self.world._marshal(components)
        self._check(lib.recs_app_command_create(self._h, ids, ptrs, len(components)), "recs_app_command_create")

    def command_remove(self, entity: Entity) -> None:
        self._check(lib.recs_app_command_remove(self._h, entity), "recs_app_command_remove")

    def last_playback(self):
        c, r = C.c_uint64(), C.c_uint64()
        self._check(lib.recs_app_last_playback(self._h, C.byref(c), C.byref(r)), "recs_app_last_playback")
        return int(c.value), int(r.value)

    def run(self, n_ticks: int = 1, ordered: bool = False) -> None:
        self._check(lib.recs_app_run(self._h, n_ticks, 1 if ordered else 0), "recs_app_run")

    def last_tick_order(self):
        cap = len(self.system_names) + 1
        order, batch = (C.c_uint32 * cap)(), (C.c_uint32 * cap)()
        n = C.c_uint32()
        self._check(lib.recs_app_last_tick_order(self._h, order, batch, cap, C.byref(n)), "recs_app_last_tick_order")
        return [int(order[i]) for i in range(n.value)], [int(batch[i]) for i in range(n.value)]

    def sync_to_host(self) -> None:
        self._check(lib.recs_app_sync_to_host(self._h), "recs_app_sync_to_host")

    def add_components(self, e: Entity, components: dict) -> None:
        """Application::add_component between ticks; resident device columns of both archetypes follow the move."""
        ids, ptrs, keep = self.world._marshal(components)
        self._check(lib.recs_app_add_components(self._h, e, ids, ptrs, len(components)), "recs_app_add_components")

    def remove_components(self, e: Entity, components: Sequence[int]) -> None:
        ids = (C.c_uint64 * max(len(components), 1))(*components)
        self._check(lib.recs_app_remove_components(self._h, e, ids, len(components)), "recs_app_remove_components")

    def mark_host_dirty(self, archetype: int) -> None:
        self._check(lib.recs_app_mark_host_dirty(self._h, archetype), "recs_app_mark_host_dirty")


def fold_swap_removes(len_before: int, rows: Sequence[int]):
    """recs_fold_swap_removes: (dst, src, len_after) equivalent to swap_remove(rows[0]), swap_remove(rows[1]), ..."""
    n = len(rows)
    arr = (C.c_uint32 * max(n, 1))(*rows)
    dst, src = (C.c_uint32 * max(n, 1))(), (C.c_uint32 * max(n, 1))()
    n_moves, len_after = C.c_uint32(), C.c_uint64()
    code = lib.recs_fold_swap_removes(len_before, arr, n, dst, src, max(n, 1), C.byref(n_moves), C.byref(len_after))
    if code:
        raise EcsError(code, "recs_fold_swap_removes")
    return [int(dst[i]) for i in range(n_moves.value)], [int(src[i]) for i in range(n_moves.value)], int(len_after.value)
