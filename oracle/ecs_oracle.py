"""ecs_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Pure-Python restatement (small cases only) of the integer / structural side of the reference's hot
path: archetype storage and indexing, query membership and filters, the par-iter segment split,
and the precedence-graph schedule.  Used by tests/ as the checker of the C++ host layer in
librecs_gpu.so (include/recs_ecs.h).  Nothing under recs_b200/ imports this module.

Pinned against the reference's own unit tests (tests/golden/ecs_golden.py lists each golden case
with its file:line).  Restated third-party behaviour the reference depends on:

* petgraph 0.6.3 `Graph` adjacency order (Cargo.lock): `add_edge` pushes the new edge at the HEAD of
  the source's outgoing list and of the target's incoming list, so `neighbors()` yields the most
  recently added edge first; `node_identifiers()` / `edge_references()` go in index order.
* daggy 0.8.0 `Dag::add_edge` = petgraph add + "would this create a cycle" rejection;
  `update_edge` = find-or-add.
* crossbeam-channel 0.5.x `Select::select()`: the registered operations are shuffled with a
  thread-local 32-bit xorshift generator seeded with 1_406_868_647 and the first ready one in
  shuffled order is taken.  The reference's schedule tests each run in a fresh thread and drop all
  guards of a batch at once, so their expected batch sequences depend on this shuffle; the
  restatement below reproduces them (tests/test_ecs_oracle.py).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Iterable, Sequence

# ==================================================================================================
# precedence (crates/scheduler/src/precedence.rs:39-82)
# ==================================================================================================
BEFORE, EQUAL, AFTER = -1, 0, 1


@dataclass(frozen=True)
class Access:
    component: int
    write: bool


@dataclass
class SystemInfo:
    name: str
    accesses: tuple  # of Access, in parameter order


def read(c):
    return Access(c, False)


def write(c):
    return Access(c, True)


def precedence_to(self: SystemInfo, other: SystemInfo) -> int:
    overlapping = [(a, b) for a in self.accesses for b in other.accesses if a.component == b.component]
    if not overlapping:
        return EQUAL
    if all((not a.write) and (not b.write) for a, b in overlapping):
        return EQUAL
    if any(b.write for _, b in overlapping):
        return AFTER
    if any(a.write and not b.write for a, b in overlapping):
        return BEFORE
    return EQUAL


# ==================================================================================================
# petgraph / daggy
# ==================================================================================================
END = -1


class Dag:
    """petgraph::Graph<_, _, Directed> adjacency lists + daggy's acyclicity check."""

    def __init__(self, n_nodes: int = 0):
        self.node_next = []  # [out_head, in_head] per node
        self.edges = []  # [source, target, next_out, next_in]
        for _ in range(n_nodes):
            self.add_node()

    def add_node(self) -> int:
        self.node_next.append([END, END])
        return len(self.node_next) - 1

    def node_count(self) -> int:
        return len(self.node_next)

    def _push_edge(self, a: int, b: int) -> int:
        idx = len(self.edges)
        self.edges.append([a, b, self.node_next[a][0], self.node_next[b][1]])
        self.node_next[a][0] = idx
        self.node_next[b][1] = idx
        return idx

    def neighbors(self, a: int):  # outgoing, newest edge first
        e = self.node_next[a][0]
        while e != END:
            yield self.edges[e][1]
            e = self.edges[e][2]

    def neighbors_incoming(self, a: int):
        e = self.node_next[a][1]
        while e != END:
            yield self.edges[e][0]
            e = self.edges[e][3]

    def degree(self, a: int) -> int:  # neighbors_undirected(a).count()
        return sum(1 for _ in self.neighbors(a)) + sum(1 for _ in self.neighbors_incoming(a))

    def find_edge(self, a: int, b: int) -> bool:
        return any(t == b for t in self.neighbors(a))

    def _reaches(self, src: int, dst: int) -> bool:
        seen, stack = {src}, [src]
        while stack:
            x = stack.pop()
            if x == dst:
                return True
            for y in self.neighbors(x):
                if y not in seen:
                    seen.add(y)
                    stack.append(y)
        return False

    def add_edge(self, a: int, b: int) -> bool:
        """daggy::Dag::add_edge: False == Err(WouldCycle)."""
        if a == b or (not self.find_edge(a, b) and self._reaches(b, a)):
            return False
        self._push_edge(a, b)
        return True

    def update_edge(self, a: int, b: int) -> bool:
        if self.find_edge(a, b):
            return True
        return self.add_edge(a, b)

    def edge_list(self):
        return [(e[0], e[1]) for e in self.edges]


class ScheduleError(Exception):
    pass


class NewTick(Exception):
    pass


class WouldBlock(Exception):
    pass


def _find_nodes(dag: Dag, systems, of_node: int, prec: int):
    # schedule/mod.rs:359-383
    return [o for o in range(dag.node_count()) if o != of_node and precedence_to(systems[of_node], systems[o]) == prec]


def generate_dag(systems: Sequence[SystemInfo], ordered: bool = False) -> Dag:
    """PrecedenceGraph::generate (schedule/mod.rs:116-149) / OrderedPrecedenceGraph::generate (:492-520)."""
    dag = Dag()
    for _ in systems:
        node = dag.add_node()
        for dependent_on in _find_nodes(dag, systems, node, AFTER):
            if ordered:
                if not dag.add_edge(dependent_on, node):
                    raise ScheduleError("Should not happen since every edge is incoming")
            elif not dag.add_edge(node, dependent_on):
                raise ScheduleError(f"dependency from `{systems[node].name}` to `{systems[dependent_on].name}` would cause a cycle")
        for dependency_of in _find_nodes(dag, systems, node, BEFORE):
            if not dag.add_edge(dependency_of, node):
                if ordered or not dag.add_edge(node, dependency_of):
                    raise ScheduleError("cycle is an impossibility when having changed edge direction")
    if not ordered:
        dag = reduce_makespan(dag)
    return dag


def reduce_makespan(dag: Dag) -> Dag:
    """schedule/mod.rs:167-198: every edge points from the node with fewer neighbours to the one with more."""
    out = Dag(dag.node_count())
    for start, end in dag.edge_list():
        if dag.degree(start) > dag.degree(end):
            source, target = end, start
        else:
            source, target = start, end
        if not out.update_edge(source, target):
            raise ScheduleError("Cycle should never be created when adjusting edge direction.")
    return out


class XorShiftShuffle:
    """crossbeam_channel::utils::shuffle with its thread-local generator (fresh thread state)."""

    def __init__(self, seed: int = 1_406_868_647):
        self.x = seed

    def shuffle(self, v: list) -> None:
        for i in range(1, len(v)):
            x = self.x
            x ^= (x << 13) & 0xFFFFFFFF
            x ^= x >> 17
            x ^= (x << 5) & 0xFFFFFFFF
            self.x = x
            n = i + 1
            j = (x * n) >> 32
            v[i], v[j] = v[j], v[i]


class PrecedenceGraph:
    """Execution protocol of schedule/mod.rs:205-333.  A dispatched system completes when
    `complete(system)` is called (== dropping its SystemExecutionGuard)."""

    RETURN_ERROR, RETURN_NEW_TICK = 0, 1

    def __init__(self, systems: Sequence[SystemInfo], ordered: bool = False, select: str = "crossbeam"):
        self.systems = list(systems)
        self.dag = generate_dag(systems, ordered)
        self.pending = []  # [node, completed]
        self.already_executed = []
        self.wait_until_next_call = False
        self.select = select
        self.rng = XorShiftShuffle()

    def complete(self, node: int) -> None:
        for p in self.pending:
            if p[0] == node and not p[1]:
                p[1] = True
                return
        raise ScheduleError(f"system {node} is not pending")

    def _initial(self):
        return [n for n in range(self.dag.node_count()) if next(self.dag.neighbors_incoming(n), None) is None]

    def _dispatch(self, nodes):
        for n in nodes:
            self.pending.append([n, False])
        return list(nodes)

    def _select_completed(self) -> int:
        """Index into self.pending of the completion `Select::select()` would report."""
        if self.select == "crossbeam":
            handles = list(range(len(self.pending)))
            if len(handles) > 1:  # run_select's single-operation fast path skips the shuffle
                self.rng.shuffle(handles)
        else:  # "first": lowest pending index that has completed
            handles = list(range(len(self.pending)))
        for i in handles:
            if self.pending[i][1]:
                return i
        raise WouldBlock()

    def _without_pending_dependencies(self):
        seen, out = set(), []
        for node, _ in self.pending:
            for nb in self.dag.neighbors(node):
                if nb in seen:
                    continue
                seen.add(nb)
                if all(p in self.already_executed for p in self.dag.neighbors_incoming(nb)):
                    out.append(nb)
        return out

    def next_batch(self, reaction: int = 0):
        while True:
            all_exec = len(self.already_executed) == self.dag.node_count() and not self.pending
            none_exec = not self.already_executed and not self.pending
            if all_exec or none_exec:
                should_return = (reaction == self.RETURN_ERROR and not self.wait_until_next_call) or reaction == self.RETURN_NEW_TICK
                if should_return:
                    self.wait_until_next_call = True
                    self.already_executed.clear()
                    self.pending.clear()
                    return self._dispatch(self._initial())
                self.wait_until_next_call = False
                raise NewTick()
            elif self.pending:
                idx = self._select_completed()
                self.already_executed.append(self.pending[idx][0])
                free = self._without_pending_dependencies()
                del self.pending[idx]
                if free:
                    return self._dispatch(free)
            else:
                raise ScheduleError("the schedule has deadlocked")

    def currently_executable_systems(self):
        """schedule tests' helper: `currently_executable_systems()` == reaction ReturnNewTick
        (crates/ecs/src/lib.rs:495-510)."""
        return self.next_batch(self.RETURN_NEW_TICK)

    def tick_batches(self):
        """All batches of one tick when every dispatched system completes before the next request."""
        batches = []
        try:
            while True:
                b = self.next_batch(self.RETURN_ERROR)
                batches.append(b)
                for n in b:
                    self.complete(n)
        except NewTick:
            return batches


# ==================================================================================================
# World / Archetype (crates/ecs/src/lib.rs:665-929, crates/ecs/src/archetypes.rs:176-787)
# ==================================================================================================
@dataclass(frozen=True, order=True)
class Entity:
    id: int
    generation: int = 0


class WorldError(Exception):
    pass


@dataclass
class Archetype:
    columns: dict = field(default_factory=dict)  # component id -> list of values
    entity_to_component_index: dict = field(default_factory=dict)
    entities: list = field(default_factory=list)

    def store_entity(self, e: Entity) -> None:  # archetypes.rs:221-232
        index = len(self.entity_to_component_index)
        if e in self.entity_to_component_index:
            raise WorldError(f"archetype already contains entity: {e}")
        self.entity_to_component_index[e] = index
        self.entities.append(e)

    def update_after_entity_removal(self, e: Entity) -> None:  # archetypes.rs:315-328
        removed = self.entity_to_component_index[e]
        if self.entities:
            self.entity_to_component_index[self.entities[-1]] = removed
        _swap_remove(self.entities, removed)
        del self.entity_to_component_index[e]


def _swap_remove(lst: list, index: int):
    value = lst[index]
    lst[index] = lst[-1]
    lst.pop()
    return value


class World:
    def __init__(self):
        self.entities = []  # Vec<Option<Entity>>
        self.deleted_entities = []  # VecDeque, index 0 == front
        self.entity_to_archetype_index = {}
        self.archetypes = []
        self.component_to_archetypes = {}
        self.typeset_to_archetype = {}

    # -- creation (archetypes.rs:338-373, lib.rs:720-726) --------------------------------------
    def create_empty_entity(self) -> Entity:
        if not self.archetypes:
            self.archetypes.append(Archetype())
        if self.deleted_entities:
            reuse = self.deleted_entities.pop(0)
            e = Entity(reuse.id, reuse.generation + 1)
            if self.entities[reuse.id] is not None:
                raise WorldError("an entity in use should never be reused")
            self.entities[reuse.id] = e
        else:
            e = Entity(len(self.entities), 0)
            self.entities.append(e)
        self.archetypes[0].store_entity(e)
        self.entity_to_archetype_index[e] = 0
        return e

    def create_entity(self, components: dict) -> Entity:
        e = self.create_empty_entity()
        self.add_components(e, components)
        return e

    def _exactly_matching(self, ids: Iterable[int]):
        ids = tuple(sorted(ids))
        if not ids:
            return 0
        return self.typeset_to_archetype.get(ids)

    def _move_between(self, e: Entity, target_index: int, to_move: Iterable[int]) -> None:  # archetypes.rs:474-518
        source_index = self.entity_to_archetype_index[e]
        if source_index == target_index:
            raise WorldError("assertion failed: source and target archetype are the same")
        source, target = self.archetypes[source_index], self.archetypes[target_index]
        ci = source.entity_to_component_index[e]
        for t in to_move:
            value = _swap_remove(source.columns[t], ci)
            target.columns.setdefault(t, []).append(value)
        target.store_entity(e)
        self.entity_to_archetype_index[e] = target_index

    def _move_to_new(self, e: Entity, to_move: Iterable[int]) -> int:  # archetypes.rs:520-543
        to_move = list(to_move)
        index = len(self.archetypes)
        self.archetypes.append(Archetype())
        for t in to_move:
            self.component_to_archetypes[t].add(index)
        self._move_between(e, index, to_move)
        return index

    def add_components(self, e: Entity, components: dict) -> None:  # archetypes.rs:556-651
        if e not in self.entity_to_archetype_index:
            raise WorldError(f"could not find the given entity: {e}")
        source_index = self.entity_to_archetype_index[e]
        source_ids = sorted(self.archetypes[source_index].columns)
        for t in components:
            if t in source_ids:
                raise WorldError(f"component of same type {t} already exists for entity {e}")
        target_ids = source_ids + list(components)
        target_index = self._exactly_matching(target_ids)
        if target_index is not None:
            self._move_between(e, target_index, source_ids)
        else:
            target_index = self._move_to_new(e, source_ids)
            self.typeset_to_archetype[tuple(sorted(target_ids))] = target_index
            for t in components:
                self.archetypes[target_index].columns.setdefault(t, [])
                self.component_to_archetypes.setdefault(t, set()).add(target_index)
        for t, value in components.items():
            self.archetypes[target_index].columns[t].append(value)
        self.archetypes[source_index].update_after_entity_removal(e)

    def remove_components(self, e: Entity, types: Sequence[int]) -> None:  # archetypes.rs:664-727
        if e not in self.entity_to_archetype_index:
            raise WorldError(f"could not find the given entity: {e}")
        source_index = self.entity_to_archetype_index[e]
        target_ids = sorted(self.archetypes[source_index].columns)
        for t in types:
            if t not in target_ids:
                raise WorldError(f"component of type {t} does not exist for entity {e}")
            target_ids.remove(t)
        target_index = self._exactly_matching(target_ids)
        if target_index is not None:
            self._move_between(e, target_index, target_ids)
        else:
            target_index = self._move_to_new(e, target_ids)
            self.typeset_to_archetype[tuple(sorted(target_ids))] = target_index
        source = self.archetypes[source_index]
        ci = source.entity_to_component_index[e]
        for t in types:
            _swap_remove(source.columns[t], ci)
        source.update_after_entity_removal(e)

    def delete_entity(self, e: Entity) -> None:  # lib.rs:728-768
        if e.id >= len(self.entities):
            raise WorldError(f"could not find the given entity: {e}")
        stored = self.entities[e.id]
        if stored is None:
            raise WorldError(f"{e} has already been deleted")
        self.entities[e.id] = None
        self.deleted_entities.insert(0, stored)
        if e not in self.entity_to_archetype_index:
            raise WorldError(f"could not find the given entity: {e}")
        arch = self.archetypes[self.entity_to_archetype_index[e]]
        if e not in arch.entity_to_component_index:
            raise WorldError(f"could not find the given entity: {e}")
        ci = arch.entity_to_component_index[e]
        for col in arch.columns.values():
            _swap_remove(col, ci)
        arch.update_after_entity_removal(e)

    # -- queries (lib.rs:842-865) ----------------------------------------------------------------
    def get_archetype_indices(self, signature: Sequence[int]) -> set:
        if not signature:
            return set(range(len(self.archetypes)))
        sets = []
        for t in signature:
            if t not in self.component_to_archetypes:
                return set()
            sets.append(self.component_to_archetypes[t])
        return set.intersection(*sets)


# ==================================================================================================
# parameters / filters (crates/ecs/src/systems/mod.rs:940-1106, crates/ecs/src/filter.rs)
# ==================================================================================================
# A parameter is a tuple: ("read", c) ("write", c) ("entity",) ("commands",) ("with", c) ("not", f)
# ("and", f, g) ("or", f, g) ("xor", f, g) ("query", [params...])
def component_accesses(params) -> list:
    out = []
    for p in params:
        if p[0] == "read":
            out.append(read(p[1]))
        elif p[0] == "write":
            out.append(write(p[1]))
        elif p[0] == "query":
            out.extend(component_accesses(p[1]))
    return out


def supports_parallelization(params) -> bool:
    if not params:
        return False
    for p in params:
        if p[0] == "query" and any(a.write for a in component_accesses(p[1])):
            return False
    return True


def controls_iteration(params) -> bool:
    return any(p[0] in ("read", "write", "entity") for p in params)


def _filter(world: World, p, universe: set) -> set:
    kind = p[0]
    if kind == "with":
        return world.get_archetype_indices([p[1]])
    if kind == "not":
        return universe - _filter(world, p[1], universe)
    if kind == "and":
        return _filter(world, p[1], universe) & _filter(world, p[2], universe)
    if kind == "or":
        return _filter(world, p[1], universe) | _filter(world, p[2], universe)
    if kind == "xor":
        return _filter(world, p[1], universe) ^ _filter(world, p[2], universe)
    return set(universe)


def get_archetype_indices(world: World, params) -> list:
    if not params:
        return []
    signature = [p[1] for p in params if p[0] in ("read", "write", "with")]
    universe = world.get_archetype_indices(signature)
    result = set(universe)
    for p in params:
        result &= _filter(world, p, universe)
    return sorted(result)


# ==================================================================================================
# segments (crates/ecs/src/systems/mod.rs:394-411, 473-540)
# ==================================================================================================
def auto_segment_size(entity_count: int, thread_count: int) -> int:
    return max(16, entity_count // (thread_count * 4))


def segment_count(entity_count: int, segment_size: int) -> int:
    return 1 + (entity_count - 1) // segment_size if entity_count != 0 else 1


def split_segments(column_lengths: Sequence[int], segment_size: int) -> list:
    """Returns segments; each segment is a list of (column, begin, len) slices, in the order the
    reference pushes segments.  segment_size == 0 -> SegmentConfig::Single."""
    if segment_size == 0:
        return [[(c, 0, n) for c, n in enumerate(column_lengths)]]
    segments, current, current_size = [], [], 0
    for c, n in enumerate(column_lengths):
        if segment_size - current_size > n:
            current_size += n
            current.append((c, 0, n))
            continue
        left = segment_size - current_size
        if left:
            current.append((c, 0, left))
            segments.append(current)
            current, current_size = [], 0
        right = n - left
        chunks, remainder = divmod(right, segment_size)
        if remainder:
            current_size += remainder
            current.append((c, left + chunks * segment_size, remainder))
        for k in range(chunks):
            segments.append([(c, left + k * segment_size, segment_size)])
    if current_size > 0 or not segments:
        segments.append(current)
    return segments
