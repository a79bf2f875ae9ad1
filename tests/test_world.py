"""Storage / query-membership / segment parity: the reference's unit tests
(crates/ecs/src/lib.rs:931-1372, crates/ecs/src/archetypes.rs:789-1119,
crates/ecs/src/systems/mod.rs:1183-1281, crates/ecs/src/systems/iteration.rs:249-408,
crates/ecs/src/filter.rs:249-529) restated against BOTH the Python oracle and the C++ host layer,
plus random operation sequences comparing the two bit for bit."""
import numpy as np
import pytest
from hypothesis import given, settings
from hypothesis import strategies as st

from oracle import ecs_oracle as O
from recs_b200 import ecs as H

# component ids standing in for TypeId::of::<u32>(), <i32>(), <usize>(), <f32>(), <i64>(), A, B, C
U32, I32, USIZE, F32, I64, CA, CB, CC = 1, 2, 3, 4, 5, 6, 7, 8
DTYPES = {U32: np.uint32, I32: np.int32, USIZE: np.uint64, F32: np.float32, I64: np.int64, CA: None, CB: None, CC: None}


class OracleWorld:
    """Adapter giving the oracle the same surface as the host World wrapper."""

    def __init__(self):
        self.w = O.World()

    def create_empty_entity(self):
        return self.w.create_empty_entity()

    def create_entity(self, comps):
        return self.w.create_entity(dict(comps))

    def add_components(self, e, comps):
        self.w.add_components(e, dict(comps))

    def remove_components(self, e, types):
        self.w.remove_components(e, list(types))

    def delete_entity(self, e):
        self.w.delete_entity(e)

    def entity_archetype(self, e):
        return self.w.entity_to_archetype_index[e]

    def entity_component_index(self, e):
        return self.w.archetypes[self.w.entity_to_archetype_index[e]].entity_to_component_index[e]

    def archetype_count(self):
        return len(self.w.archetypes)

    def archetype_entities(self, a):
        return [(e.id, e.generation) for e in self.w.archetypes[a].entities]

    def archetype_components(self, a):
        return sorted(self.w.archetypes[a].columns)

    def column_values(self, a, c):
        return list(self.w.archetypes[a].columns[c])

    def entity_slots(self):
        return [None if e is None else (e.id, e.generation) for e in self.w.entities]

    def get_archetype_indices(self, sig):
        return sorted(self.w.get_archetype_indices(list(sig)))

    def query_archetypes(self, params):
        return O.get_archetype_indices(self.w, params)

    def key(self, e):
        return (e.id, e.generation)

    error = O.WorldError


class HostWorld:
    def __init__(self):
        self.w = H.World()
        for c, dt in DTYPES.items():
            self.w.register_component(c, dt)

    def create_empty_entity(self):
        return self.w.create_empty_entity()

    def create_entity(self, comps):
        return self.w.create_entity(dict(comps))

    def add_components(self, e, comps):
        self.w.add_components(e, dict(comps))

    def remove_components(self, e, types):
        self.w.remove_components(e, list(types))

    def delete_entity(self, e):
        self.w.delete_entity(e)

    def entity_archetype(self, e):
        return self.w.entity_archetype(e)

    def entity_component_index(self, e):
        return self.w.entity_component_index(e)

    def archetype_count(self):
        return self.w.archetype_count()

    def archetype_entities(self, a):
        return self.w.archetype_entities(a)

    def archetype_components(self, a):
        return self.w.archetype_components(a)

    def column_values(self, a, c):
        return [v.item() if hasattr(v, "item") else v for v in self.w.column(a, c)]

    def entity_slots(self):
        return self.w.entity_slots()

    def get_archetype_indices(self, sig):
        return self.w.get_archetype_indices(list(sig))

    def query_archetypes(self, params):
        return self.w.query_archetypes(params)

    def key(self, e):
        return e.key()

    error = H.EcsError


WORLDS = [pytest.param(OracleWorld, id="oracle"), pytest.param(HostWorld, id="host-c++")]


def setup3(W):
    """setup_world_with_3_entities_with_u32_and_i32_components (lib.rs:1041-1052)."""
    w = W()
    e1 = w.create_entity({U32: 1, I32: 1})
    e2 = w.create_entity({U32: 2, I32: 2})
    e3 = w.create_entity({U32: 3, I32: 3})
    return w, w.entity_archetype(e1), e1, e2, e3


@pytest.mark.parametrize("W", WORLDS)
class TestWorldGoldens:
    def test_entities_change_archetype_after_component_addition(self, W):  # lib.rs:976-990
        w = W()
        e = w.create_entity({CA: None})
        assert w.entity_archetype(e) == 1
        w.add_components(e, {CB: None})
        assert w.entity_archetype(e) == 2

    def test_entities_change_archetype_after_component_removal(self, W):  # lib.rs:992-1009
        w = W()
        e = w.create_entity({CA: None})
        assert w.entity_archetype(e) == 1
        w.remove_components(e, [CA])
        assert w.entity_archetype(e) == 0

    def test_last_entity_moves_between_archetypes_as_expected(self, W):  # lib.rs:1011-1039
        w = W()
        e = w.create_empty_entity()
        w.add_components(e, {CA: None})
        w.add_components(e, {CB: None})
        assert w.entity_archetype(e) == 2
        w.remove_components(e, [CA])
        assert w.entity_archetype(e) == 3
        w.remove_components(e, [CB])
        assert w.entity_archetype(e) == 0

    def test_body_archetype_layout_of_the_benchmark(self, W):
        w = W()
        es = [w.create_entity({U32: k, I32: -k, F32: 0.5 * k, I64: k}) for k in range(5)]
        assert w.archetype_count() == 2
        assert [w.key(e) for e in es] == [(k, 0) for k in range(5)]
        assert all(w.entity_archetype(e) == 1 and w.entity_component_index(e) == k for k, e in enumerate(es))
        assert w.archetype_entities(0) == [] and w.archetype_entities(1) == [(k, 0) for k in range(5)]
        assert w.column_values(1, U32) == list(range(5))

    def test_deleted_entities_leave_archetype_empty(self, W):  # lib.rs:1054-1068, 1109-1117
        w, arch, e1, e2, e3 = setup3(W)
        for e in (e1, e2, e3):
            w.delete_entity(e)
        assert w.column_values(arch, U32) == [] and w.column_values(arch, I32) == []
        assert w.entity_slots() == [None, None, None]

    def test_removing_entity_does_not_change_other_values(self, W):  # lib.rs:1070-1107
        w, arch, e1, e2, e3 = setup3(W)
        w.delete_entity(e2)
        u, i = w.column_values(arch, U32), w.column_values(arch, I32)
        assert (u[w.entity_component_index(e1)], u[w.entity_component_index(e3)]) == (1, 3)
        assert (i[w.entity_component_index(e1)], i[w.entity_component_index(e3)]) == (1, 3)

    def test_values_after_adding_components(self, W):  # lib.rs:1119-1132
        w, arch, *_ = setup3(W)
        assert w.column_values(arch, U32) == [1, 2, 3] and w.column_values(arch, I32) == [1, 2, 3]

    def test_values_swap_index_after_move_by_addition(self, W):  # lib.rs:1134-1174, archetypes.rs:914-926
        w, arch, e1, e2, e3 = setup3(W)
        w.add_components(e1, {USIZE: 1})
        assert w.column_values(arch, U32) == [3, 2] and w.column_values(arch, I32) == [3, 2]
        assert w.archetype_entities(arch) == [w.key(e3), w.key(e2)]
        assert w.entity_slots() == [w.key(e1), w.key(e2), w.key(e3)]  # lib.rs:1176-1188
        assert w.entity_archetype(e1) == arch + 1 and w.entity_archetype(e2) == arch and w.entity_archetype(e3) == arch
        assert w.entity_component_index(e2) == 1 and w.entity_component_index(e3) == 0  # archetypes.rs:978-996

    def test_values_after_removing_components(self, W):  # lib.rs:1207-1227, 1229-1269, archetypes.rs:928-1019
        w, arch, e1, e2, e3 = setup3(W)
        w.remove_components(e1, [U32])
        assert w.column_values(arch, U32) == [3, 2] and w.column_values(arch, I32) == [3, 2]
        assert w.column_values(arch + 1, I32) == [1]
        assert w.archetype_entities(arch) == [w.key(e3), w.key(e2)]
        assert w.entity_archetype(e1) == arch + 1
        assert w.entity_component_index(e1) == 0 and w.entity_component_index(e2) == 1 and w.entity_component_index(e3) == 0

    def test_error_when_adding_existing_component_type(self, W):  # lib.rs:1287-1299
        w = W()
        e = w.create_empty_entity()
        w.add_components(e, {USIZE: 10})
        with pytest.raises(W.error):
            w.add_components(e, {USIZE: 20})

    def test_error_when_removing_nonexistent_component_type(self, W):  # lib.rs:1301-1311
        w = W()
        e = w.create_empty_entity()
        with pytest.raises(W.error):
            w.remove_components(e, [I32])

    def test_entities_maintain_component_values_after_moving(self, W):  # lib.rs:1313-1342
        w = W()
        e = w.create_empty_entity()
        w.add_components(e, {USIZE: 10})
        w.add_components(e, {F32: 123.0})
        w.add_components(e, {I64: 321})
        w.remove_components(e, [F32])
        assert w.entity_archetype(e) == 4
        assert w.column_values(4, USIZE) == [10] and w.column_values(4, I64) == [321]

    def test_entity_ids_are_reused_with_bumped_generation(self, W):  # archetypes.rs:344-358, lib.rs:751-760
        w, arch, e1, e2, e3 = setup3(W)
        w.delete_entity(e1)
        w.delete_entity(e3)
        n1 = w.create_entity({U32: 7, I32: 7})
        n2 = w.create_entity({U32: 8, I32: 8})
        # push_front + pop_front: the most recently deleted id is reused first
        assert w.key(n1) == (2, 1) and w.key(n2) == (0, 1)
        assert w.column_values(arch, U32) == [2, 7, 8]

    def test_borrowing_nonexistent_component_matches_nothing(self, W):  # archetypes.rs:1040-1059
        w, *_ = setup3(W)
        assert w.get_archetype_indices([F32]) == []

    def test_query_membership_with_signature(self, W):  # lib.rs:947-974, archetypes.rs:1021-1038
        w, arch, *_ = setup3(W)
        assert w.get_archetype_indices([U32]) == [arch]
        assert w.get_archetype_indices([U32, I32]) == [arch]
        assert w.get_archetype_indices([]) == [0, 1]

    def test_entity_only_query_sees_every_archetype(self, W):  # iteration.rs:375-407
        w = W()
        w.create_empty_entity()
        w.create_entity({CA: None})
        w.create_entity({CA: None, CB: None})
        w.create_entity({CA: None, CB: None, CC: None})
        archs = w.query_archetypes([("entity",)])
        assert archs == [0, 1, 2, 3]
        assert sum(len(w.archetype_entities(a)) for a in archs) == 4

    def test_filters(self, W):  # filter.rs:42-247 semantics on a fixed world
        w = W()
        w.create_entity({CA: None})  # arch 1 {A}
        w.create_entity({CA: None, CB: None})  # arch 2 {A,B}
        w.create_entity({CB: None, CC: None})  # arch 3 {B,C}
        w.create_entity({CA: None, CB: None, CC: None})  # arch 4 {A,B,C}
        q = w.query_archetypes
        assert q([("read", CA)]) == [1, 2, 4]
        assert q([("read", CA), ("with", CB)]) == [2, 4]
        assert q([("read", CA), ("not", ("with", CB))]) == [1]
        assert q([("read", CA), ("or", ("with", CB), ("with", CC))]) == [2, 4]
        assert q([("read", CA), ("and", ("with", CB), ("with", CC))]) == [4]
        assert q([("read", CA), ("xor", ("with", CB), ("with", CC))]) == [2]
        assert q([("read", CA), ("not", ("or", ("with", CB), ("with", CC)))]) == [1]
        assert q([]) == []  # impl SystemParameters for ()
        # nested Query has its own match (mod.rs:1042) and no base signature (mod.rs:1097-1099)
        assert q([("read", CA), ("query", [("read", CB)])]) == [1, 2, 4]


def test_host_lock_conflict_message():
    """world_panics_when_trying_to_mutably_borrow_same_components_twice (archetypes.rs:838-851)."""
    w = H.World()
    w.register_component(CA, None, "ecs::archetypes::tests::A")
    w.create_entity({CA: None})
    w.borrow_column(1, CA, True)
    with pytest.raises(H.EcsError) as e:
        w.borrow_column(1, CA, True)
    assert e.value.code == 8
    assert "Lock of ComponentVec<ecs::archetypes::tests::A> is already taken!" in str(e.value)
    w.release_column(1, CA, True)
    w.borrow_column(1, CA, False)  # :866-874 two immutable borrows are fine
    w.borrow_column(1, CA, False)
    with pytest.raises(H.EcsError):
        w.borrow_column(1, CA, True)


# ---- parameters ------------------------------------------------------------------------------------
PARAM_CASES = [
    [],
    [("read", 1)],
    [("write", 1), ("read", 2)],
    [("read", 1), ("write", 4), ("query", [("read", 1), ("read", 2)])],  # gravity (crates/n-body/src/lib.rs:106-110)
    [("read", 1), ("query", [("write", 2)])],
    [("entity",), ("read", 3), ("commands",)],  # ground_collision
    [("read", 1), ("or", ("with", 2), ("not", ("with", 3)))],
]


@pytest.mark.parametrize("params", PARAM_CASES, ids=repr)
def test_parameter_traits_match_oracle(params):
    assert H.component_accesses(params) == [(a.component, a.write) for a in O.component_accesses(params)]
    assert H.supports_parallelization(params) == O.supports_parallelization(params)
    assert H.controls_iteration(params) == O.controls_iteration(params)


def test_gravity_access_order_and_parallelisation():
    gravity = PARAM_CASES[3]
    # flattened in parameter order (mod.rs:973-978): R Position, W Acceleration, R Position, R Mass
    assert H.component_accesses(gravity) == [(1, False), (4, True), (1, False), (2, False)]
    assert H.supports_parallelization(gravity) is True  # read-only nested query (mod.rs:1101-1105)
    assert H.supports_parallelization(PARAM_CASES[4]) is False
    assert H.supports_parallelization([]) is False  # mod.rs:954-957


# ---- segments (mod.rs:394-411, 1216-1280) ------------------------------------------------------------
def test_segment_size_constants():
    # MINIMUM_ENTITIES_PER_SEGMENT = 16, SEGMENTS_PER_THREAD = 4
    for f in (H.auto_segment_size, O.auto_segment_size):
        assert f(1000, 8) == 31 and f(100, 8) == 16 and f(65536, 16) == 1024 and f(0, 4) == 16
    for f in (H.segment_count, O.segment_count):
        assert f(0, 16) == 1 and f(16, 16) == 1 and f(17, 16) == 2 and f(1000, 31) == 33


@settings(max_examples=300, deadline=None)
@given(st.lists(st.integers(0, 200), min_size=0, max_size=8), st.integers(1, 64))
def test_segments_property_and_host_matches_oracle(lengths, size):
    """segments_are_correct_size proptest (mod.rs:1216-1280): every segment but the last holds
    exactly `size` items, the last 1 + (N - 1) % size; plus host == oracle slice for slice."""
    seg_o = O.split_segments(lengths, size)
    seg_h = H.split_segments(lengths, size)
    assert seg_h == seg_o
    n = sum(lengths)
    sizes = [sum(s[2] for s in seg) for seg in seg_o]
    assert len(seg_o) == O.segment_count(n, size)
    if n:
        # every item appears exactly once
        covered = sorted((c, b + k) for seg in seg_o for c, b, ln in seg for k in range(ln))
        assert covered == [(c, k) for c, ln in enumerate(lengths) for k in range(ln)]
        by_size = sorted(sizes, reverse=True)
        assert all(x == size for x in by_size[:-1]) and by_size[-1] == 1 + (n - 1) % size
    single_o, single_h = O.split_segments(lengths, 0), H.split_segments(lengths, 0)
    assert single_o == single_h == [[(c, 0, ln) for c, ln in enumerate(lengths)]]


# ---- random structural operation sequences: host vs oracle ----------------------------------------------
COMPS = [U32, I32, USIZE, F32, I64]
op = st.one_of(
    st.tuples(st.just("create"), st.sets(st.sampled_from(COMPS), min_size=1, max_size=4), st.integers(0, 1000)),
    st.tuples(st.just("empty")),
    st.tuples(st.just("add"), st.integers(0, 40), st.sets(st.sampled_from(COMPS), min_size=1, max_size=2), st.integers(0, 1000)),
    st.tuples(st.just("remove"), st.integers(0, 40), st.sets(st.sampled_from(COMPS), min_size=1, max_size=2)),
    st.tuples(st.just("delete"), st.integers(0, 40)),
)


@settings(max_examples=150, deadline=None)
@given(st.lists(op, min_size=1, max_size=40))
def test_random_operations_host_matches_oracle(ops):
    o, h = OracleWorld(), HostWorld()
    eo, eh = [], []  # live entity handles, same order on both sides
    for step in ops:
        kind = step[0]
        try:
            if kind == "create":
                comps = {c: step[2] + c for c in sorted(step[1])}
                eo.append(o.create_entity(comps))
                eh.append(h.create_entity(comps))
            elif kind == "empty":
                eo.append(o.create_empty_entity())
                eh.append(h.create_empty_entity())
            elif not eo:
                continue
            elif kind == "add":
                k = step[1] % len(eo)
                comps = {c: step[3] + c for c in sorted(step[2])}
                err_o = err_h = None
                try:
                    o.add_components(eo[k], comps)
                except O.WorldError as e:
                    err_o = e
                try:
                    h.add_components(eh[k], comps)
                except H.EcsError as e:
                    err_h = e
                assert (err_o is None) == (err_h is None)
            elif kind == "remove":
                k = step[1] % len(eo)
                err_o = err_h = None
                try:
                    o.remove_components(eo[k], sorted(step[2]))
                except O.WorldError as e:
                    err_o = e
                try:
                    h.remove_components(eh[k], sorted(step[2]))
                except H.EcsError as e:
                    err_h = e
                assert (err_o is None) == (err_h is None)
            elif kind == "delete":
                k = step[1] % len(eo)
                o.delete_entity(eo[k])
                h.delete_entity(eh[k])
                eo.pop(k)
                eh.pop(k)
        finally:
            pass
        assert [o.key(e) for e in eo] == [h.key(e) for e in eh]
    # full structural comparison: bit-exact ids, archetype indices, component indices, columns
    assert o.archetype_count() == h.archetype_count()
    assert o.entity_slots() == h.entity_slots()
    for a in range(o.archetype_count()):
        assert o.archetype_entities(a) == h.archetype_entities(a)
        assert o.archetype_components(a) == h.archetype_components(a)
        for c in o.archetype_components(a):
            assert o.column_values(a, c) == [int(v) if c != F32 else float(v) for v in h.column_values(a, c)]
    for x, y in zip(eo, eh):
        assert o.entity_archetype(x) == h.entity_archetype(y)
        assert o.entity_component_index(x) == h.entity_component_index(y)
    for sig in ([], [U32], [U32, I32], [F32, I64, USIZE]):
        assert o.get_archetype_indices(sig) == h.get_archetype_indices(sig)


# ---- resident command playback: swap-removes folded into row moves ---------------------------------------------------
@settings(max_examples=200, deadline=None)
@given(st.data())
def test_fold_swap_removes_equals_sequential_swap_removes(data):
    """recs_fold_swap_removes (what recs_app replays on device columns through recs_gpu_archetype_move_rows) against
    `Vec::swap_remove` applied one call at a time (crates/ecs/src/archetypes.rs:233-243)."""
    from recs_b200.ecs import fold_swap_removes

    n = data.draw(st.integers(min_value=1, max_value=60))
    column = list(range(100, 100 + n))
    rows = []
    cur = list(column)
    for _ in range(data.draw(st.integers(min_value=0, max_value=n))):
        r = data.draw(st.integers(min_value=0, max_value=len(cur) - 1))
        rows.append(r)
        cur[r] = cur[-1]  # swap_remove
        cur.pop()
    dst, src, len_after = fold_swap_removes(n, rows)
    assert len_after == len(cur)
    assert len(set(dst)) == len(dst) and all(d < len_after for d in dst)
    assert len(dst) <= len(rows)
    got = list(column)
    moved = {d: column[s] for d, s in zip(dst, src)}  # every source is read before any destination is written
    for d, v in moved.items():
        got[d] = v
    assert got[:len_after] == cur


def test_fold_swap_removes_rejects_rows_out_of_range():
    from recs_b200.ecs import EcsError, fold_swap_removes

    with pytest.raises(EcsError):
        fold_swap_removes(3, [3])
    with pytest.raises(EcsError):
        fold_swap_removes(1, [0, 0])
    assert fold_swap_removes(4, [3, 2]) == ([], [], 2)  # removing from the end moves nothing


def test_batch_create_equals_single_creates():
    """recs_world_create_entities takes a batch path once the archetype of the component set exists (no stop in the
    empty-entity archetype per entity).  The end state must be the one the reference's entity-by-entity creation produces:
    entity ids and generations (deleted ids are reused most-recent first), archetype indices, row order, component
    indices, column bytes -- through random interleavings of batch creates over several component sets, single creates and
    deletions."""
    from recs_b200 import ecs

    A, B, Cc, Z = 1, 2, 3, 4
    dtypes = {A: np.dtype((np.float32, 3)), B: np.float32, Cc: np.dtype((np.uint8, 6)), Z: np.dtype((np.uint8, 0))}
    sets = [(A,), (B, A), (A, B, Cc), (Cc,), (Z, A)]
    rng = np.random.default_rng(77)
    worlds = [ecs.World(), ecs.World()]
    for w in worlds:
        for c, dt in dtypes.items():
            w.register_component(c, dt, f"c{c}")

    def columns(ids, m):
        out = {}
        for c in ids:
            dt = np.dtype(dtypes[c])
            shape = (m,) + dt.shape
            base = dt.base
            if base == np.float32:
                out[c] = rng.standard_normal(shape).astype(np.float32)
            else:
                out[c] = rng.integers(0, 255, shape).astype(np.uint8) if dt.shape != (0,) else np.zeros((m, 0), np.uint8)
        return out

    def state(w):
        snap = {"slots": w.entity_slots(), "archetypes": []}
        for a in range(w.archetype_count()):
            ents = w.archetype_entities(a)
            comps = w.archetype_components(a)
            snap["archetypes"].append((ents, comps, [w.column(a, c).tobytes() for c in comps],
                                       [w.entity_component_index(ecs.Entity(*e)) for e in ents]))
        return snap

    for step in range(60):
        op = rng.choice(["batch", "batch", "single", "delete"])
        if op == "batch":
            ids = sets[rng.integers(len(sets))]
            m = int(rng.integers(1, 40))
            cols = columns(ids, m)
            worlds[0].create_entities(m, cols)
            for k in range(m):
                worlds[1].create_entity({c: cols[c][k] for c in ids})
        elif op == "single":
            ids = sets[rng.integers(len(sets))]
            cols = columns(ids, 1)
            for w in worlds:
                w.create_entity({c: cols[c][0] for c in ids})
        else:
            alive = [e for e in worlds[0].entity_slots() if e is not None]
            for e in rng.permutation(len(alive))[: int(rng.integers(0, 12))]:
                for w in worlds:
                    w.delete_entity(ecs.Entity(*alive[e]))
        assert state(worlds[0]) == state(worlds[1]), step
    assert worlds[0].archetype_count() == len(sets) + 1
    assert any(gen > 0 for slot in worlds[0].entity_slots() if slot for gen in [slot[1]])  # ids were reused
