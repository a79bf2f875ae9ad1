// PrecedenceGraph / OrderedPrecedenceGraph: crates/scheduler/src/precedence.rs:39-82 and
// crates/scheduler/src/schedule/mod.rs:116-333, 492-620 restated, including the container
// behaviour the per-tick system ORDER depends on:
//   * petgraph 0.6.3 adjacency lists: a new edge becomes the head of its source's outgoing list
//     and of its target's incoming list, so `neighbors()` yields the newest edge first;
//   * daggy 0.8.0 `add_edge` rejects an edge that would close a cycle, `update_edge` is
//     find-or-add;
//   * crossbeam-channel `Select::select()` over the pending completion channels: operations are
//     shuffled with a thread-local xorshift32 generator (seed 1406868647) and the first ready one
//     is taken.  Which completed system is acknowledged first decides which systems become
//     runnable next, so the reference's own expected batch sequences
//     (schedule/mod.rs:1093-1275) can only be reproduced with this policy
//     (RECS_SELECT_CROSSBEAM).  RECS_SELECT_FIRST takes the oldest pending completion instead.
#include <algorithm>

#include "host.h"

namespace recs {
namespace host {

// Orderable::precedence_to (precedence.rs:39-66)
int precedence(const SysInfo& self, const SysInfo& other) {
    bool any_overlap = false, both_read_all = true, other_writes_any = false, other_reads_any = false;
    for (const recs_gpu_access& a : self.access)
        for (const recs_gpu_access& b : other.access) {
            if (a.component_id != b.component_id) continue;
            any_overlap = true;
            if (a.is_write || b.is_write) both_read_all = false;
            if (b.is_write) other_writes_any = true;
            if (a.is_write && !b.is_write) other_reads_any = true;
        }
    if (!any_overlap) return 0;
    if (both_read_all) return 0;
    if (other_writes_any) return 1;   // Precedence::After
    if (other_reads_any) return -1;   // Precedence::Before
    return 0;
}

namespace {

constexpr uint32_t kEnd = 0xFFFFFFFFu;

struct Dag {
    struct Node {
        uint32_t next[2] = {kEnd, kEnd};
    };
    struct Edge {
        uint32_t node[2];
        uint32_t next[2];
    };
    std::vector<Node> nodes;
    std::vector<Edge> edges;

    uint32_t add_node() {
        nodes.push_back(Node());
        return static_cast<uint32_t>(nodes.size() - 1);
    }
    void push_edge(uint32_t a, uint32_t b) {
        Edge e;
        e.node[0] = a;
        e.node[1] = b;
        e.next[0] = nodes[a].next[0];
        e.next[1] = nodes[b].next[1];
        const uint32_t idx = static_cast<uint32_t>(edges.size());
        edges.push_back(e);
        nodes[a].next[0] = idx;
        nodes[b].next[1] = idx;
    }
    template <typename F>
    void for_out(uint32_t a, F f) const {
        for (uint32_t e = nodes[a].next[0]; e != kEnd; e = edges[e].next[0]) f(edges[e].node[1]);
    }
    template <typename F>
    void for_in(uint32_t a, F f) const {
        for (uint32_t e = nodes[a].next[1]; e != kEnd; e = edges[e].next[1]) f(edges[e].node[0]);
    }
    uint32_t degree(uint32_t a) const {  // neighbors_undirected(a).count()
        uint32_t d = 0;
        for_out(a, [&](uint32_t) { ++d; });
        for_in(a, [&](uint32_t) { ++d; });
        return d;
    }
    bool find_edge(uint32_t a, uint32_t b) const {
        bool found = false;
        for_out(a, [&](uint32_t t) { found = found || t == b; });
        return found;
    }
    bool reaches(uint32_t src, uint32_t dst) const {
        std::vector<char> seen(nodes.size(), 0);
        std::vector<uint32_t> stack(1, src);
        seen[src] = 1;
        while (!stack.empty()) {
            const uint32_t x = stack.back();
            stack.pop_back();
            if (x == dst) return true;
            for_out(x, [&](uint32_t y) {
                if (!seen[y]) {
                    seen[y] = 1;
                    stack.push_back(y);
                }
            });
        }
        return false;
    }
    bool add_edge(uint32_t a, uint32_t b) {  // false == Err(WouldCycle)
        if (a == b || (!find_edge(a, b) && reaches(b, a))) return false;
        push_edge(a, b);
        return true;
    }
    bool update_edge(uint32_t a, uint32_t b) { return find_edge(a, b) ? true : add_edge(a, b); }
};

}  // namespace
}  // namespace host
}  // namespace recs

using namespace recs::host;

struct recs_schedule {
    std::vector<SysInfo> systems;
    Dag dag;
    struct Pending {
        uint32_t node;
        bool completed;
    };
    std::vector<Pending> pending;
    std::vector<uint32_t> already_executed;
    bool wait_until_next_call = false;
    uint32_t rng = 1406868647u;  // crossbeam_channel::utils::shuffle's thread-local seed
    int32_t select_policy = 0;   // 0 = crossbeam shuffle, 1 = first completed
    std::string error;
};

namespace {

std::vector<uint32_t> find_nodes(const recs_schedule& s, uint32_t of_node, int prec) {
    std::vector<uint32_t> found;  // schedule/mod.rs:359-383
    for (uint32_t o = 0; o < s.dag.nodes.size(); ++o) {
        if (o == of_node) continue;
        if (precedence(s.systems[of_node], s.systems[o]) == prec) found.push_back(o);
    }
    return found;
}

bool reduce_makespan(recs_schedule& s) {  // schedule/mod.rs:167-198
    Dag out;
    for (size_t i = 0; i < s.dag.nodes.size(); ++i) out.add_node();
    for (const Dag::Edge& e : s.dag.edges) {
        const uint32_t start = e.node[0], end = e.node[1];
        uint32_t source = start, target = end;
        if (s.dag.degree(start) > s.dag.degree(end)) {
            source = end;
            target = start;
        }
        if (!out.update_edge(source, target)) return false;
    }
    s.dag = out;
    return true;
}

void shuffle(recs_schedule& s, std::vector<uint32_t>& v) {
    for (size_t i = 1; i < v.size(); ++i) {
        uint32_t x = s.rng;
        x ^= x << 13;
        x ^= x >> 17;
        x ^= x << 5;
        s.rng = x;
        const uint64_t n = i + 1;
        const size_t j = static_cast<size_t>((static_cast<uint64_t>(x) * n) >> 32);
        std::swap(v[i], v[j]);
    }
}

std::vector<uint32_t> initial_systems(const recs_schedule& s) {  // schedule/mod.rs:615-620
    std::vector<uint32_t> out;
    for (uint32_t n = 0; n < s.dag.nodes.size(); ++n)
        if (s.dag.nodes[n].next[1] == kEnd) out.push_back(n);
    return out;
}

std::vector<uint32_t> without_pending_dependencies(const recs_schedule& s) {  // schedule/mod.rs:299-312
    std::vector<uint32_t> out, seen;
    for (const recs_schedule::Pending& p : s.pending) {
        s.dag.for_out(p.node, [&](uint32_t nb) {
            if (std::find(seen.begin(), seen.end(), nb) != seen.end()) return;
            seen.push_back(nb);
            bool all = true;
            s.dag.for_in(nb, [&](uint32_t prereq) {
                if (std::find(s.already_executed.begin(), s.already_executed.end(), prereq) == s.already_executed.end())
                    all = false;
            });
            if (all) out.push_back(nb);
        });
    }
    return out;
}

int32_t dispatch(recs_schedule& s, const std::vector<uint32_t>& nodes, uint32_t* out, uint32_t cap, uint32_t* out_n) {
    for (uint32_t n : nodes) s.pending.push_back(recs_schedule::Pending{n, false});
    if (out_n) *out_n = static_cast<uint32_t>(nodes.size());
    if (out)
        for (uint32_t i = 0; i < nodes.size() && i < cap; ++i) out[i] = nodes[i];
    return RECS_OK;
}

// get_next_systems_to_run (schedule/mod.rs:205-268)
int32_t next_batch(recs_schedule& s, int32_t reaction, uint32_t* out, uint32_t cap, uint32_t* out_n) {
    for (;;) {
        const bool all_exec = s.already_executed.size() == s.dag.nodes.size() && s.pending.empty();
        const bool none_exec = s.already_executed.empty() && s.pending.empty();
        if (all_exec || none_exec) {
            const bool should_return = (reaction == RECS_REACTION_RETURN_ERROR && !s.wait_until_next_call) ||
                                       reaction == RECS_REACTION_RETURN_NEW_TICK;
            if (should_return) {
                s.wait_until_next_call = true;
                s.already_executed.clear();
                s.pending.clear();
                return dispatch(s, initial_systems(s), out, cap, out_n);
            }
            s.wait_until_next_call = false;
            if (out_n) *out_n = 0;
            return RECS_SCHEDULE_NEW_TICK;
        } else if (!s.pending.empty()) {
            // wait_for_pending_completion: Select over the pending receivers
            std::vector<uint32_t> handles(s.pending.size());
            for (uint32_t i = 0; i < handles.size(); ++i) handles[i] = i;
            if (s.select_policy == 0 && handles.size() > 1) shuffle(s, handles);
            uint32_t idx = kEnd;
            for (uint32_t h : handles)
                if (s.pending[h].completed) {
                    idx = h;
                    break;
                }
            if (idx == kEnd) {
                if (out_n) *out_n = 0;
                return RECS_SCHEDULE_WOULD_BLOCK;
            }
            s.already_executed.push_back(s.pending[idx].node);
            const std::vector<uint32_t> free_now = without_pending_dependencies(s);
            s.pending.erase(s.pending.begin() + idx);
            if (!free_now.empty()) return dispatch(s, free_now, out, cap, out_n);
        } else {
            s.error = "the schedule has deadlocked due to there being no pending systems but all systems have still not run";
            return RECS_ERR_SCHEDULE;
        }
    }
}

SysInfo to_info(const recs_system_info& in) {
    SysInfo s;
    s.name = in.name ? in.name : "";
    s.access.assign(in.access, in.access + in.n_access);
    return s;
}

}  // namespace

extern "C" {

int32_t recs_precedence(const recs_system_info* self, const recs_system_info* other) {
    if (!self || !other) return 0;
    return precedence(to_info(*self), to_info(*other));
}

int32_t recs_schedule_generate(const recs_system_info* systems, uint32_t n, int32_t ordered, recs_schedule** out) {
    if (!out || (n && !systems)) return RECS_ERR_INVALID_ARGUMENT;
    recs_schedule* s = new recs_schedule();
    for (uint32_t i = 0; i < n; ++i) s->systems.push_back(to_info(systems[i]));
    // Schedule::generate (schedule/mod.rs:116-149; ordered variant :492-520)
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t node = s->dag.add_node();
        for (uint32_t dependent_on : find_nodes(*s, node, 1)) {
            const bool ok = ordered ? s->dag.add_edge(dependent_on, node) : s->dag.add_edge(node, dependent_on);
            if (!ok) {
                delete s;
                return RECS_ERR_SCHEDULE;
            }
        }
        for (uint32_t dependency_of : find_nodes(*s, node, -1)) {
            if (!s->dag.add_edge(dependency_of, node)) {
                if (ordered || !s->dag.add_edge(node, dependency_of)) {
                    delete s;
                    return RECS_ERR_SCHEDULE;
                }
            }
        }
    }
    if (!ordered && !reduce_makespan(*s)) {
        delete s;
        return RECS_ERR_SCHEDULE;
    }
    *out = s;
    return RECS_OK;
}

int32_t recs_schedule_destroy(recs_schedule* s) {
    if (!s) return RECS_ERR_INVALID_ARGUMENT;
    delete s;
    return RECS_OK;
}

int32_t recs_schedule_edges(recs_schedule* s, uint32_t* from, uint32_t* to, uint32_t cap, uint32_t* out_n) {
    if (!s) return RECS_ERR_INVALID_ARGUMENT;
    if (out_n) *out_n = static_cast<uint32_t>(s->dag.edges.size());
    for (uint32_t i = 0; i < s->dag.edges.size() && i < cap; ++i) {
        if (from) from[i] = s->dag.edges[i].node[0];
        if (to) to[i] = s->dag.edges[i].node[1];
    }
    return RECS_OK;
}

int32_t recs_schedule_set_select_policy(recs_schedule* s, int32_t policy) {
    if (!s || policy < 0 || policy > 1) return RECS_ERR_INVALID_ARGUMENT;
    s->select_policy = policy;
    return RECS_OK;
}

int32_t recs_schedule_next_batch(recs_schedule* s, int32_t reaction, uint32_t* out, uint32_t cap, uint32_t* out_n) {
    if (!s) return RECS_ERR_INVALID_ARGUMENT;
    return next_batch(*s, reaction, out, cap, out_n);
}

int32_t recs_schedule_complete(recs_schedule* s, uint32_t system) {
    if (!s) return RECS_ERR_INVALID_ARGUMENT;
    for (recs_schedule::Pending& p : s->pending)
        if (p.node == system && !p.completed) {
            p.completed = true;
            return RECS_OK;
        }
    s->error = "system is not pending";
    return RECS_ERR_SCHEDULE;
}

int32_t recs_schedule_tick_order(recs_schedule* s, uint32_t* out_order, uint32_t* out_batch_of, uint32_t cap,
                                 uint32_t* out_n) {
    if (!s) return RECS_ERR_INVALID_ARGUMENT;
    std::vector<uint32_t> batch(s->systems.size() + 1);
    uint32_t n = 0, batch_no = 0;
    for (;;) {
        uint32_t got = 0;
        const int32_t st = next_batch(*s, RECS_REACTION_RETURN_ERROR, batch.data(), (uint32_t)batch.size(), &got);
        if (st == RECS_SCHEDULE_NEW_TICK) break;
        if (st != RECS_OK) return st;
        for (uint32_t i = 0; i < got; ++i) {
            if (n < cap) {
                if (out_order) out_order[n] = batch[i];
                if (out_batch_of) out_batch_of[n] = batch_no;
            }
            ++n;
            recs_schedule_complete(s, batch[i]);
        }
        ++batch_no;
    }
    if (out_n) *out_n = n;
    return RECS_OK;
}

}  // extern "C"
