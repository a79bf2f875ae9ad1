"""Golden cases transcribed from the reference's own unit tests (file:line under the reference
checkout).  Each case is data only; tests/test_ecs_oracle.py runs them against the Python oracle
and tests/test_ecs_host.py against the C++ host layer of librecs_gpu.so.

Component ids stand in for the `TypeId`s of test_utils' A, B, C (crates/test_utils/src/lib.rs:9-24).
A system is (name, [(component, is_write), ...]) in parameter order.
"""
A, B, C = 1, 2, 3
R = lambda c: (c, False)  # noqa: E731
W = lambda c: (c, True)  # noqa: E731

# crates/test_utils/src/lib.rs:26-42
read_a = ("read_a", [R(A)])
read_b = ("read_b", [R(B)])
read_c = ("read_c", [R(C)])
other_read_a = ("other_read_a", [R(A)])
write_a = ("write_a", [W(A)])
write_b = ("write_b", [W(B)])
write_c = ("write_c", [W(C)])
other_write_a = ("other_write_a", [W(A)])
read_b_write_a = ("read_b_write_a", [R(B), W(A)])
read_a_write_c = ("read_a_write_c", [R(A), W(C)])
read_c_write_b = ("read_c_write_b", [R(C), W(B)])
read_a_write_b = ("read_a_write_b", [R(A), W(B)])
read_ab = ("read_ab", [R(A), R(B)])
write_ab = ("write_ab", [W(A), W(B)])

BEFORE, EQUAL, AFTER = -1, 0, 1

# crates/scheduler/src/precedence.rs:92-169: (self, other, expected self.precedence_to(other))
PRECEDENCE = [
    (read_a, write_a, AFTER),  # :98-107
    (write_a, read_a, BEFORE),
    (write_a, other_write_a, AFTER),  # :110-120
    (other_write_a, write_a, AFTER),
    (read_a, other_read_a, EQUAL),  # :123-132
    (other_read_a, read_a, EQUAL),
    (read_b_write_a, read_a, BEFORE),  # :135-144
    (read_a, read_b_write_a, AFTER),
    (read_b_write_a, read_a_write_c, BEFORE),  # :147-158
    (read_a_write_c, read_b_write_a, AFTER),
    (read_a_write_b, read_ab, BEFORE),  # :161-168
    (read_ab, read_a_write_b, AFTER),
]

# crates/scheduler/src/schedule/mod.rs:703-937: (name, systems, expected edges as index pairs,
# "a runs before b").  The reference compares by graph isomorphism; node order is the same on both
# sides so index pairs are compared as sets.
DAGS = [
    ("schedule_does_not_connect_independent_systems :703-715", [read_a, read_b], []),
    ("schedule_represents_component_dependencies_as_edges :717-738", [read_a, write_a, read_b, write_b], [(0, 1), (2, 3)]),
    ("schedule_allows_multiple_reads_to_run_simultaneously :740-760", [read_a, read_a, write_a], [(0, 2), (1, 2)]),
    ("schedule_ignores_non_overlapping_components :762-783", [read_a, read_a_write_c, read_b_write_a], [(0, 2), (1, 2)]),
    ("schedule_places_multiple_writes_in_sequence :785-807", [read_a, write_a, write_a], [(2, 1), (0, 2), (0, 1)]),
    ("schedule_allows_concurrent_component_writes_to_separate_components :826-848", [write_a, read_a, write_b, read_b], [(1, 0), (3, 2)]),
    (
        "preserves_precedence_in_multi_layer_schedule :850-889",
        [write_ab, write_b, write_a, read_b, read_a, read_a_write_c, read_c],
        [(1, 0), (2, 0), (3, 0), (4, 0), (5, 0), (3, 1), (4, 2), (5, 2), (6, 5)],
    ),
    ("multiple_writes_are_placed_in_sequence_for_multicomponent_systems :891-912", [read_ab, write_ab, read_b_write_a], [(0, 1), (2, 1), (0, 2)]),
    ("prevents_deadlock_by_scheduling_deadlocking_accesses_sequentially :914-929", [read_a_write_b, read_b_write_a], [(1, 0)]),
    ("schedule_reorders_systems_to_reduce_makespan :931-950", [write_a, write_ab, write_b], [(0, 1), (2, 1)]),
]
# :809-824 must merely generate without error
DAG_GENERATES = [("schedule_avoids_cycle_when_systems_have_cyclic_dependencies :809-824", [read_a_write_c, read_b_write_a, read_c_write_b])]

_MULTI = [write_ab, write_b, write_a, read_b, read_a, read_a_write_c, read_c]
# Batch sequences when every guard of a batch is dropped before the next request
# (crates/scheduler/src/schedule/mod.rs:1093-1275): (name, systems, expected batches as indices, repeats)
BATCHES = [
    ("dag_execution_traversal_begins_with... :1093-1118 + ...gives_entire_layers... :1120-1157", _MULTI, [[3, 4, 6], [5], [2], [1], [0]], 1),
    ("dag_execution_repeats_once_fully_executed :1208-1224", [read_a, write_a], [[0], [1], [0]], 1),
    ("dag_execution_remains_same_during_several_loops :1226-1252", [read_ab, read_a, read_a_write_b, write_ab], [[0, 1], [2], [3]], 3),
    ("multiple_writes_are_executed_in_sequence_for_multicomponent_systems :1254-1275", [read_ab, write_ab, read_a_write_b], [[0], [2], [1]], 1),
]

# n-body triple of crates/scheduler/src/schedule/mod.rs:1307-1339 (ACC, POS, VEL ids)
ACC, POS, VEL, MASS = 10, 11, 12, 13
NBODY_TEST_TRIPLE = [("acceleration", [R(ACC), W(VEL)]), ("gravity", [W(ACC), R(POS)]), ("movement", [R(VEL), W(POS)])]

# Benchmark registration (crates/ecs_bench_suite/src/recs/n_body.rs:54-59) with the real accesses
# (crates/n-body/src/lib.rs:89, 97, 106-110) -- tick order derived in SURVEY 3.1.
NBODY_BENCH = [
    ("movement", [W(POS), R(VEL)]),
    ("acceleration", [W(VEL), R(ACC)]),
    ("gravity", [R(POS), W(ACC), R(POS), R(MASS)]),
    ("benchmark_system", []),
]
NBODY_BENCH_LAYERS = [{"gravity", "benchmark_system"}, {"movement"}, {"acceleration"}]
# demo set: + collision(Read<Position>, Query<Read<Position>>...) (crates/n-body/examples/n_body_demo.rs:39-42)
NBODY_DEMO = NBODY_BENCH[:3] + [("collision", [R(POS), R(POS)])]
NBODY_DEMO_LAYERS = [{"gravity", "collision"}, {"acceleration"}, {"movement"}]

# rain benchmark registration (crates/ecs_bench_suite/src/recs/rain_simulation.rs:56-65);
# accesses from crates/rain-simulation/src/lib.rs:65, 86, 78, 94, 128-133, 152, 158
RVEL, RPOS, RMASS, CLOUD = 20, 21, 22, 23
RAIN_BENCH = [
    ("gravity", [W(RVEL)]),
    ("wind", [W(RVEL)]),
    ("wind_direction", []),
    ("movement", [W(RPOS), R(RVEL)]),
    ("rain", [R(CLOUD), R(RPOS), W(RMASS)]),
    ("ground_collision", [R(RPOS)]),
    ("vapor_accumulation", [R(CLOUD), W(RMASS)]),
    ("benchmark_system", []),
]
# partial order that constrains results: wind before gravity before movement
RAIN_MUST_PRECEDE = [("wind", "gravity"), ("gravity", "movement"), ("wind", "movement"), ("rain", "movement"),
                     ("ground_collision", "movement"), ("vapor_accumulation", "rain")]
