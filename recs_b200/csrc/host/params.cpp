// System parameters, query membership and the par-iter segment split:
// crates/ecs/src/systems/mod.rs:264-352, 394-411, 457-566, 940-1106 and crates/ecs/src/filter.rs.
#include <algorithm>

#include "host.h"

namespace recs {
namespace host {

namespace {

bool is_filter(uint32_t op) {
    return op == RECS_PARAM_WITH || op == RECS_PARAM_NOT || op == RECS_PARAM_AND || op == RECS_PARAM_OR ||
           op == RECS_PARAM_XOR;
}

// Recursive-descent over the prefix token string.
bool parse_one(const recs_param_token* t, uint32_t n, uint32_t* pos, Param* out, bool filter_only, std::string* err) {
    if (*pos >= n) {
        *err = "parameter list ends in the middle of an expression";
        return false;
    }
    const recs_param_token tok = t[(*pos)++];
    out->op = tok.op;
    out->comp = tok.component_id;
    if (filter_only && !is_filter(tok.op)) {
        *err = "filter combinators (Not/And/Or/Xor) only take filters";
        return false;
    }
    switch (tok.op) {
        case RECS_PARAM_READ:
        case RECS_PARAM_WRITE:
        case RECS_PARAM_ENTITY:
        case RECS_PARAM_COMMANDS:
        case RECS_PARAM_WITH:
            return true;
        case RECS_PARAM_NOT:
            out->children.resize(1);
            return parse_one(t, n, pos, &out->children[0], true, err);
        case RECS_PARAM_AND:
        case RECS_PARAM_OR:
        case RECS_PARAM_XOR:
            out->children.resize(2);
            return parse_one(t, n, pos, &out->children[0], true, err) &&
                   parse_one(t, n, pos, &out->children[1], true, err);
        case RECS_PARAM_QUERY_BEGIN:
            while (*pos < n && t[*pos].op != RECS_PARAM_QUERY_END) {
                Param child;
                if (!parse_one(t, n, pos, &child, false, err)) return false;
                out->children.push_back(child);
            }
            if (*pos >= n) {
                *err = "QUERY_BEGIN without QUERY_END";
                return false;
            }
            ++*pos;
            return true;
        default:
            *err = "unknown parameter op " + std::to_string(tok.op);
            return false;
    }
}

// SystemParameter::base_signature (mod.rs:329-339; Read/Write :560-566, With filter.rs:69-71)
bool base_signature(const Param& p, uint64_t* out) {
    if (p.op == RECS_PARAM_READ || p.op == RECS_PARAM_WRITE || p.op == RECS_PARAM_WITH) {
        *out = p.comp;
        return true;
    }
    return false;
}

std::set<uint32_t> set_and(const std::set<uint32_t>& a, const std::set<uint32_t>& b) {
    std::set<uint32_t> r;
    for (uint32_t x : a)
        if (b.count(x)) r.insert(x);
    return r;
}

// SystemParameter::filter (mod.rs:341-347 default = universe; filter.rs:73-78, 133-137, 238-246)
std::set<uint32_t> filter(const recs_world& w, const Param& p, const std::set<uint32_t>& universe) {
    switch (p.op) {
        case RECS_PARAM_WITH:
            return world_archetype_indices(w, &p.comp, 1);
        case RECS_PARAM_NOT: {
            const std::set<uint32_t> inner = filter(w, p.children[0], universe);
            std::set<uint32_t> r;
            for (uint32_t x : universe)
                if (!inner.count(x)) r.insert(x);
            return r;
        }
        case RECS_PARAM_AND:
            return set_and(filter(w, p.children[0], universe), filter(w, p.children[1], universe));
        case RECS_PARAM_OR: {
            std::set<uint32_t> r = filter(w, p.children[0], universe);
            const std::set<uint32_t> b = filter(w, p.children[1], universe);
            r.insert(b.begin(), b.end());
            return r;
        }
        case RECS_PARAM_XOR: {
            const std::set<uint32_t> a = filter(w, p.children[0], universe);
            const std::set<uint32_t> b = filter(w, p.children[1], universe);
            std::set<uint32_t> r;
            for (uint32_t x : a)
                if (!b.count(x)) r.insert(x);
            for (uint32_t x : b)
                if (!a.count(x)) r.insert(x);
            return r;
        }
        default:
            return universe;
    }
}

}  // namespace

bool parse_params(const recs_param_token* tokens, uint32_t n, std::vector<Param>* out, std::string* err) {
    uint32_t pos = 0;
    out->clear();
    while (pos < n) {
        Param p;
        if (!parse_one(tokens, n, &pos, &p, false, err)) return false;
        out->push_back(p);
    }
    return true;
}

// component_accesses, flattened in parameter order (mod.rs:973-978; Query :1088-1090)
void params_accesses(const std::vector<Param>& params, std::vector<recs_gpu_access>* out) {
    for (const Param& p : params) {
        if (p.op == RECS_PARAM_READ || p.op == RECS_PARAM_WRITE) {
            recs_gpu_access a;
            a.component_id = p.comp;
            a.is_write = p.op == RECS_PARAM_WRITE ? 1 : 0;
            out->push_back(a);
        } else if (p.op == RECS_PARAM_QUERY_BEGIN) {
            params_accesses(p.children, out);
        }
    }
}

// supports_parallelization: no parameters -> false (mod.rs:954-957); a Query is parallelisable
// only if all of its accesses are reads (mod.rs:1101-1105); everything else defaults to true.
bool params_supports_parallelization(const std::vector<Param>& params) {
    if (params.empty()) return false;
    for (const Param& p : params) {
        if (p.op == RECS_PARAM_QUERY_BEGIN) {
            std::vector<recs_gpu_access> acc;
            params_accesses(p.children, &acc);
            for (const recs_gpu_access& a : acc)
                if (a.is_write) return false;
        }
    }
    return true;
}

// controls_iteration: Read / Write / Entity do, Query / filters / Commands do not.
bool params_controls_iteration(const std::vector<Param>& params) {
    for (const Param& p : params)
        if (p.op == RECS_PARAM_READ || p.op == RECS_PARAM_WRITE || p.op == RECS_PARAM_ENTITY) return true;
    return false;
}

// SystemParameters::get_archetype_indices (mod.rs:986-1000): universe = archetypes holding every
// base-signature component; result = universe ∩ each parameter's filter(universe).
// The reference collects a hash set with identity hashing into a Vec; the order returned here is
// ascending archetype index (unpinned by the reference's tests, single archetype in all configs).
std::vector<uint32_t> params_archetypes(const recs_world& w, const std::vector<Param>& params) {
    if (params.empty()) return std::vector<uint32_t>();  // impl SystemParameters for () (mod.rs:959-961)
    std::vector<uint64_t> signature;
    for (const Param& p : params) {
        uint64_t c;
        if (base_signature(p, &c)) signature.push_back(c);
    }
    const std::set<uint32_t> universe =
        world_archetype_indices(w, signature.data(), static_cast<uint32_t>(signature.size()));
    std::set<uint32_t> result = universe;
    for (const Param& p : params) result = set_and(result, filter(w, p, universe));
    return std::vector<uint32_t>(result.begin(), result.end());
}

}  // namespace host
}  // namespace recs

using namespace recs::host;

extern "C" {

int32_t recs_params_component_accesses(const recs_param_token* tokens, uint32_t n, recs_gpu_access* out, uint32_t cap,
                                       uint32_t* out_count) {
    std::vector<Param> params;
    std::string err;
    if (!parse_params(tokens, n, &params, &err)) return RECS_ERR_INVALID_ARGUMENT;
    std::vector<recs_gpu_access> acc;
    params_accesses(params, &acc);
    if (out_count) *out_count = static_cast<uint32_t>(acc.size());
    if (out)
        for (uint32_t i = 0; i < acc.size() && i < cap; ++i) out[i] = acc[i];
    return RECS_OK;
}

int32_t recs_params_supports_parallelization(const recs_param_token* tokens, uint32_t n, int32_t* out) {
    std::vector<Param> params;
    std::string err;
    if (!out || !parse_params(tokens, n, &params, &err)) return RECS_ERR_INVALID_ARGUMENT;
    *out = params_supports_parallelization(params) ? 1 : 0;
    return RECS_OK;
}

int32_t recs_params_controls_iteration(const recs_param_token* tokens, uint32_t n, int32_t* out) {
    std::vector<Param> params;
    std::string err;
    if (!out || !parse_params(tokens, n, &params, &err)) return RECS_ERR_INVALID_ARGUMENT;
    *out = params_controls_iteration(params) ? 1 : 0;
    return RECS_OK;
}

int32_t recs_params_get_archetype_indices(recs_world* world, const recs_param_token* tokens, uint32_t n,
                                          int32_t query_index, uint32_t* out, uint32_t cap, uint32_t* out_count) {
    if (!world) return RECS_ERR_INVALID_ARGUMENT;
    std::vector<Param> params;
    std::string err;
    if (!parse_params(tokens, n, &params, &err)) return world->fail(RECS_ERR_INVALID_ARGUMENT, err);
    const std::vector<Param>* use = &params;
    if (query_index >= 0) {
        int32_t seen = 0;
        use = nullptr;
        for (const Param& p : params)
            if (p.op == RECS_PARAM_QUERY_BEGIN && seen++ == query_index) use = &p.children;
        if (!use) return world->fail(RECS_ERR_INVALID_ARGUMENT, "no such nested Query parameter");
    }
    const std::vector<uint32_t> archs = params_archetypes(*world, *use);
    if (out_count) *out_count = static_cast<uint32_t>(archs.size());
    if (out)
        for (uint32_t i = 0; i < archs.size() && i < cap; ++i) out[i] = archs[i];
    return RECS_OK;
}

// calculate_auto_segment_size (mod.rs:397-403): MINIMUM_ENTITIES_PER_SEGMENT = 16, SEGMENTS_PER_THREAD = 4
uint64_t recs_auto_segment_size(uint64_t entity_count, uint32_t thread_count) {
    const uint64_t threads = thread_count ? thread_count : 1;
    const uint64_t size = entity_count / (threads * 4);
    return size < 16 ? 16 : size;
}

// calculate_segment_count (mod.rs:405-411)
uint64_t recs_segment_count(uint64_t entity_count, uint64_t segment_size) {
    if (entity_count != 0 && segment_size != 0) return 1 + (entity_count - 1) / segment_size;
    return 1;
}

// Read::split_borrowed_data (mod.rs:473-540); Write (:645-710) and Entity (:807-872) split the same way.
int32_t recs_split_segments(const uint64_t* lengths, uint32_t n_columns, uint64_t segment_size, recs_segment_slice* out,
                            uint64_t cap, uint64_t* out_slices, uint64_t* out_segments) {
    if (n_columns && !lengths) return RECS_ERR_INVALID_ARGUMENT;
    uint64_t n_out = 0;
    auto emit = [&](uint32_t segment, uint32_t column, uint64_t begin, uint64_t len) {
        if (out && n_out < cap) {
            out[n_out].segment = segment;
            out[n_out].column = column;
            out[n_out].begin = begin;
            out[n_out].len = len;
        }
        ++n_out;
    };
    if (segment_size == 0) {  // SegmentConfig::Single: one segment holding every column whole
        for (uint32_t c = 0; c < n_columns; ++c) emit(0, c, 0, lengths[c]);
        if (out_slices) *out_slices = n_out;
        if (out_segments) *out_segments = 1;
        return RECS_OK;
    }
    // `segments` (finished) and `current_segment` (open) of the reference; the open segment's
    // final index is only known when it is pushed, so its slices are buffered.
    struct Pending {
        uint32_t column;
        uint64_t begin, len;
    };
    std::vector<Pending> current;
    uint64_t current_size = 0;
    uint32_t n_segments = 0;
    auto push_current = [&]() {
        for (const Pending& p : current) emit(n_segments, p.column, p.begin, p.len);
        current.clear();
        current_size = 0;
        ++n_segments;
    };
    for (uint32_t c = 0; c < n_columns; ++c) {
        const uint64_t len = lengths[c];
        if (segment_size - current_size > len) {
            current_size += len;
            current.push_back(Pending{c, 0, len});
            continue;
        }
        const uint64_t left = segment_size - current_size;  // split_at
        uint64_t right_begin = left;
        if (left != 0) {
            current.push_back(Pending{c, 0, left});
            push_current();
        }
        const uint64_t right_len = len - left;
        const uint64_t n_chunks = right_len / segment_size;  // chunks_exact
        const uint64_t remainder = right_len % segment_size;
        if (remainder != 0) {
            current_size += remainder;
            current.push_back(Pending{c, right_begin + n_chunks * segment_size, remainder});
        }
        for (uint64_t k = 0; k < n_chunks; ++k) {
            emit(n_segments, c, right_begin + k * segment_size, segment_size);
            ++n_segments;
        }
    }
    if (current_size > 0 || n_segments == 0) push_current();
    if (out_slices) *out_slices = n_out;
    if (out_segments) *out_segments = n_segments;
    return RECS_OK;
}

}  // extern "C"
