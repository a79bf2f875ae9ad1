"""Schedule parity: the reference's own golden cases (tests/golden/ecs_golden.py) against BOTH the
Python oracle (oracle/ecs_oracle.py) and the C++ host layer (librecs_gpu.so via recs_b200.ecs), and
the two against each other on random system sets (the reference's proptests,
crates/scheduler/src/schedule/mod.rs:952-1073)."""
import pytest
from hypothesis import given, settings
from hypothesis import strategies as st

from oracle import ecs_oracle as O
from recs_b200 import ecs as H
from tests.golden import ecs_golden as G


class OracleSchedule:
    def __init__(self, systems, ordered=False, select="crossbeam"):
        infos = [O.SystemInfo(n, tuple(O.Access(c, w) for c, w in acc)) for n, acc in systems]
        self.g = O.PrecedenceGraph(infos, ordered=ordered, select=select)

    def edges(self):
        return self.g.dag.edge_list()

    def batch(self):
        return self.g.currently_executable_systems()

    def tick_batch(self):
        try:
            return self.g.next_batch(0)
        except O.NewTick:
            return "new_tick"

    def complete(self, i):
        self.g.complete(i)


class HostSchedule:
    def __init__(self, systems, ordered=False, select="crossbeam"):
        self.g = H.PrecedenceGraph(systems, ordered=ordered, select=H.SELECT_CROSSBEAM if select == "crossbeam" else H.SELECT_FIRST)

    def edges(self):
        return self.g.edges()

    def batch(self):
        return self.g.currently_executable_systems()

    def tick_batch(self):
        return self.g.next_batch(H.REACTION_RETURN_ERROR)

    def complete(self, i):
        self.g.complete(i)


IMPLS = [pytest.param(OracleSchedule, id="oracle"), pytest.param(HostSchedule, id="host-c++")]


def precedence_oracle(a, b):
    mk = lambda s: O.SystemInfo(s[0], tuple(O.Access(c, w) for c, w in s[1]))  # noqa: E731
    return O.precedence_to(mk(a), mk(b))


@pytest.mark.parametrize("fn", [pytest.param(precedence_oracle, id="oracle"), pytest.param(H.precedence, id="host-c++")])
@pytest.mark.parametrize("case", G.PRECEDENCE, ids=lambda c: f"{c[0][0]}->{c[1][0]}")
def test_precedence_goldens(fn, case):
    a, b, expected = case
    assert fn(a, b) == expected


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("case", G.DAGS, ids=lambda c: c[0].split()[0])
def test_dag_goldens(impl, case):
    _, systems, expected = case
    assert sorted(impl(systems).edges()) == sorted(expected)


@pytest.mark.parametrize("impl", IMPLS)
def test_cyclic_dependencies_generate(impl):
    for _, systems in G.DAG_GENERATES:
        edges = impl(systems).edges()
        assert len(edges) == 3  # every conflicting pair is ordered, no cycle


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("case", G.BATCHES, ids=lambda c: c[0].split()[0])
def test_batch_sequence_goldens(impl, case):
    _, systems, expected, repeats = case
    s = impl(systems)
    got = []
    for _ in range(repeats):
        for _ in expected:
            b = s.batch()
            got.append(b)
            for i in b:  # the tests drop every guard of the batch before asking again
                s.complete(i)
    assert got == expected * repeats


def _layers(impl, systems, select="first"):
    s = impl(systems, select=select)
    names = [n for n, _ in systems]
    layers, order = [], []
    while True:
        b = s.tick_batch()
        if b == "new_tick":
            break
        layers.append({names[i] for i in b})
        order.extend(names[i] for i in b)
        for i in b:
            s.complete(i)
    return layers, order


@pytest.mark.parametrize("impl", IMPLS)
def test_nbody_benchmark_tick_order(impl):
    """crates/ecs_bench_suite/src/recs/n_body.rs:54-59 -> gravity, then movement, then acceleration."""
    layers, order = _layers(impl, G.NBODY_BENCH)
    assert layers == G.NBODY_BENCH_LAYERS
    assert [n for n in order if n != "benchmark_system"] == ["gravity", "movement", "acceleration"]


@pytest.mark.parametrize("impl", IMPLS)
def test_nbody_demo_tick_order_differs(impl):
    """With the demo's extra `collision` system reduce_makespan flips movement/acceleration."""
    layers, _ = _layers(impl, G.NBODY_DEMO)
    assert layers == G.NBODY_DEMO_LAYERS


@pytest.mark.parametrize("impl", IMPLS)
def test_nbody_test_triple_has_no_concurrent_conflicts(impl):
    """crates/scheduler/src/schedule/mod.rs:1307-1339."""
    layers, order = _layers(impl, G.NBODY_TEST_TRIPLE)
    assert sorted(order) == ["acceleration", "gravity", "movement"]
    assert all(len(layer) == 1 for layer in layers)
    assert order == ["movement", "acceleration", "gravity"]


@pytest.mark.parametrize("impl", IMPLS)
def test_rain_benchmark_partial_order(impl):
    _, order = _layers(impl, G.RAIN_BENCH)
    assert sorted(order) == sorted(n for n, _ in G.RAIN_BENCH)
    for before, after in G.RAIN_MUST_PRECEDE:
        assert order.index(before) < order.index(after), (before, after)


@pytest.mark.parametrize("impl", IMPLS)
def test_tick_barrier_and_would_block(impl):
    """tick_barrier_prevents_fast_system... (schedule/mod.rs:1277-1305): the next tick cannot start
    while a system of the current one is still pending."""
    s = impl([G.read_a, G.read_b])
    first = s.batch()
    assert first == [0, 1]
    s.complete(0)
    if impl is HostSchedule:
        assert s.batch() == "would_block"
    else:
        with pytest.raises(O.WouldBlock):
            s.batch()
    s.complete(1)
    assert s.batch() == [0, 1]


# ---- random system sets: the reference's proptests + oracle/host cross-check ----------------------
access = st.tuples(st.integers(0, 7), st.booleans())
system = st.lists(access, min_size=0, max_size=4, unique_by=lambda a: a[0])
systems_strategy = st.lists(system, min_size=1, max_size=10).map(lambda ss: [(f"s{i}", acc) for i, acc in enumerate(ss)])


def _conflict(a, b):
    return any(ca == cb and (wa or wb) for ca, wa in a[1] for cb, wb in b[1])


@settings(max_examples=200, deadline=None)
@given(systems_strategy, st.booleans())
def test_random_systems_oracle_and_host_agree_and_never_overlap_conflicts(systems, ordered):
    o, h = OracleSchedule(systems, ordered=ordered), HostSchedule(systems, ordered=ordered)
    assert o.edges() == h.edges()  # identical DAG, edge for edge, in petgraph edge-index order
    executed = []
    for _ in range(2):  # two ticks
        while True:
            bo, bh = o.tick_batch(), h.tick_batch()
            assert bo == bh
            if bo == "new_tick":
                break
            # currently_executable_systems_does_not_contain_concurrent_{writes, reads_and_writes}
            for i in bo:
                for j in bo:
                    if i != j:
                        assert not _conflict(systems[i], systems[j]), (systems[i], systems[j])
            executed.extend(bo)
            for i in bo:
                o.complete(i)
                h.complete(i)
    # currently_executable_systems_walk_through_all_systems_once_completed
    assert sorted(executed) == sorted(list(range(len(systems))) * 2)
