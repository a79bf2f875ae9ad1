// Internal C++ declarations of the host-side ECS mirror (include/recs_ecs.h is the public C ABI).
#pragma once
#include <stdint.h>

#include <deque>
#include <map>
#include <set>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../../include/recs_ecs.h"

namespace recs {
namespace host {

inline uint64_t entity_key(recs_entity e) {
    // `impl Hash for Entity`: generation << 32 | id (crates/ecs/src/lib.rs:918-927)
    return (static_cast<uint64_t>(e.generation) << 32) | e.id;
}

struct ComponentInfo {
    uint32_t elem = 0;
    std::string name;
};

// One `RwLock<Vec<T>>` (crates/ecs/src/archetypes.rs:15).
struct ColumnStore {
    std::vector<uint8_t> data;
    uint32_t elem = 0;
    uint64_t len = 0;
    int readers = 0;
    bool writer = false;

    void push(const void* value);
    void swap_remove(uint64_t index, void* out_value /* may be null */);
};

// `struct Archetype` (crates/ecs/src/archetypes.rs:176-182).
struct Archetype {
    std::map<uint64_t, ColumnStore> columns;  // component_typeid_to_component_vec
    std::unordered_map<uint64_t, uint64_t> entity_to_component_index;
    std::vector<recs_entity> entities;
    uint64_t version = 0;
};

struct EntitySlot {
    bool alive = false;
    recs_entity entity{0, 0};
};

}  // namespace host
}  // namespace recs

// `struct World` (crates/ecs/src/lib.rs:665-700).
struct recs_world {
    std::vector<recs::host::EntitySlot> entities;           // Vec<Option<Entity>>
    std::deque<recs_entity> deleted_entities;               // VecDeque<Entity>
    std::unordered_map<uint64_t, uint32_t> entity_to_archetype_index;
    std::vector<recs::host::Archetype> archetypes;
    std::map<uint64_t, std::set<uint32_t>> component_typeid_to_archetype_indices;
    std::map<std::vector<uint64_t>, uint32_t> component_typeids_set_to_archetype_index;
    std::map<uint64_t, recs::host::ComponentInfo> registry;
    std::string error;

    int32_t fail(int32_t code, const std::string& msg) {
        error = msg;
        return code;
    }
};

namespace recs {
namespace host {

// ---- parameters ---------------------------------------------------------------------------------
struct Param {
    uint32_t op = 0;
    uint64_t comp = 0;
    std::vector<Param> children;  // QUERY: its parameters; NOT: 1 filter; AND/OR/XOR: 2 filters
};
bool parse_params(const recs_param_token* tokens, uint32_t n, std::vector<Param>* out, std::string* err);
void params_accesses(const std::vector<Param>& params, std::vector<recs_gpu_access>* out);
bool params_supports_parallelization(const std::vector<Param>& params);
bool params_controls_iteration(const std::vector<Param>& params);
std::vector<uint32_t> params_archetypes(const recs_world& w, const std::vector<Param>& params);
std::set<uint32_t> world_archetype_indices(const recs_world& w, const uint64_t* signature, uint32_t n);

// ---- schedule -----------------------------------------------------------------------------------
struct SysInfo {
    std::string name;
    std::vector<recs_gpu_access> access;
};
int precedence(const SysInfo& self, const SysInfo& other);

}  // namespace host
}  // namespace recs
