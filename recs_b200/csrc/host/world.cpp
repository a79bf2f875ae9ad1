// World / Archetype column storage: the indexing rules of crates/ecs/src/lib.rs:665-929 and
// crates/ecs/src/archetypes.rs:176-787 restated so that entity ids, archetype indices, component
// indices and column contents come out identical for the same sequence of operations.
#include <string.h>

#include <algorithm>

#include "host.h"

using namespace recs::host;

namespace recs {
namespace host {

void ColumnStore::push(const void* value) {
    if (elem) {
        const size_t old = data.size();
        data.resize(old + elem);
        if (value)
            memcpy(data.data() + old, value, elem);
        else
            memset(data.data() + old, 0, elem);
    }
    ++len;
}

// Vec::swap_remove: the last element takes the removed slot.
void ColumnStore::swap_remove(uint64_t index, void* out_value) {
    if (elem) {
        uint8_t* base = data.data();
        if (out_value) memcpy(out_value, base + index * elem, elem);
        if (index != len - 1) memcpy(base + index * elem, base + (len - 1) * elem, elem);
        data.resize((len - 1) * static_cast<size_t>(elem));
    }
    --len;
}

namespace {

std::string entity_str(recs_entity e) {
    return "Entity { id: " + std::to_string(e.id) + ", generation: " + std::to_string(e.generation) + " }";
}

// Archetype::store_entity (archetypes.rs:221-232)
int32_t store_entity(recs_world& w, Archetype& a, recs_entity e) {
    const uint64_t index = a.entity_to_component_index.size();
    auto ins = a.entity_to_component_index.insert(std::make_pair(entity_key(e), index));
    if (!ins.second) {
        ins.first->second = index;
        return w.fail(RECS_ERR_WORLD, "archetype already contains entity: " + entity_str(e));
    }
    a.entities.push_back(e);
    a.version++;
    return RECS_OK;
}

// Archetype::update_after_entity_removal (archetypes.rs:315-328)
int32_t update_after_entity_removal(recs_world& w, Archetype& a, recs_entity e) {
    auto it = a.entity_to_component_index.find(entity_key(e));
    if (it == a.entity_to_component_index.end())
        return w.fail(RECS_ERR_WORLD, "could not find component index of entity: " + entity_str(e));
    const uint64_t removed = it->second;
    if (!a.entities.empty()) a.entity_to_component_index[entity_key(a.entities.back())] = removed;
    // Vec::swap_remove
    a.entities[removed] = a.entities.back();
    a.entities.pop_back();
    a.entity_to_component_index.erase(entity_key(e));
    a.version++;
    return RECS_OK;
}

std::vector<uint64_t> component_types(const Archetype& a) {
    std::vector<uint64_t> ids;
    for (auto& kv : a.columns) ids.push_back(kv.first);
    return ids;  // std::map keeps them sorted, which is also the sorted key order
}

// World::get_exactly_matching_archetype (archetypes.rs:753-768)
bool exactly_matching(const recs_world& w, std::vector<uint64_t> ids, uint32_t* out) {
    if (ids.empty()) {
        *out = 0;  // EMPTY_ENTITY_ARCHETYPE_INDEX
        return true;
    }
    std::sort(ids.begin(), ids.end());
    auto it = w.component_typeids_set_to_archetype_index.find(ids);
    if (it == w.component_typeids_set_to_archetype_index.end()) return false;
    *out = it->second;
    return true;
}

// World::move_entity_components_between_archetypes (archetypes.rs:474-518)
int32_t move_between(recs_world& w, recs_entity e, uint32_t target_index, const std::vector<uint64_t>& to_move) {
    auto src_it = w.entity_to_archetype_index.find(entity_key(e));
    if (src_it == w.entity_to_archetype_index.end())
        return w.fail(RECS_ERR_WORLD, "could not find the given entity: " + entity_str(e));
    const uint32_t source_index = src_it->second;
    if (source_index == target_index)  // get_mut_at_two_indices: assert_ne!(first_index, second_index)
        return w.fail(RECS_ERR_WORLD, "assertion failed: source and target archetype are the same");
    Archetype& source = w.archetypes[source_index];
    Archetype& target = w.archetypes[target_index];
    const uint64_t source_component_index = source.entity_to_component_index.at(entity_key(e));
    std::vector<uint8_t> tmp;
    for (uint64_t type_id : to_move) {
        ColumnStore& from = source.columns.at(type_id);
        tmp.resize(from.elem ? from.elem : 1);
        from.swap_remove(source_component_index, tmp.data());  // ComponentVec::move_element
        auto it = target.columns.find(type_id);
        if (it == target.columns.end()) {
            ColumnStore fresh;
            fresh.elem = from.elem;
            it = target.columns.insert(std::make_pair(type_id, fresh)).first;
        }
        it->second.push(tmp.data());
    }
    source.version++;
    target.version++;
    int32_t st = store_entity(w, target, e);
    if (st != RECS_OK) return st;
    w.entity_to_archetype_index[entity_key(e)] = target_index;
    return RECS_OK;
}

// World::move_entity_components_to_new_archetype (archetypes.rs:520-543)
int32_t move_to_new(recs_world& w, recs_entity e, const std::vector<uint64_t>& to_move, uint32_t* out_index) {
    const uint32_t new_index = static_cast<uint32_t>(w.archetypes.size());
    w.archetypes.push_back(Archetype());
    for (uint64_t type_id : to_move) w.component_typeid_to_archetype_indices[type_id].insert(new_index);
    int32_t st = move_between(w, e, new_index, to_move);
    if (st != RECS_OK) return st;
    *out_index = new_index;
    return RECS_OK;
}

int32_t check_registered(recs_world& w, const uint64_t* ids, uint32_t n) {
    for (uint32_t i = 0; i < n; ++i)
        if (!w.registry.count(ids[i]))
            return w.fail(RECS_ERR_INVALID_ARGUMENT, "component " + std::to_string(ids[i]) + " is not registered");
    return RECS_OK;
}

// The id half of World::create_empty_entity (lib.rs:701-718): reuse the most recently deleted id with its generation
// bumped, else the next fresh one.
int32_t allocate_entity(recs_world& w, recs_entity* out) {
    if (w.archetypes.empty()) w.archetypes.push_back(Archetype());  // the "empty entity archetype"
    recs_entity e;
    if (!w.deleted_entities.empty()) {
        const recs_entity reuse = w.deleted_entities.front();
        w.deleted_entities.pop_front();
        e = reuse;
        e.generation += 1;  // keeps track of which ids have been reused
        if (reuse.id >= w.entities.size())
            return w.fail(RECS_ERR_WORLD, "could not find the given entity: " + entity_str(e));
        if (w.entities[reuse.id].alive)
            return w.fail(RECS_ERR_WORLD, "an entity in use should never be reused");
        w.entities[reuse.id].alive = true;
        w.entities[reuse.id].entity = e;
    } else {
        e.id = static_cast<uint32_t>(w.entities.size());
        e.generation = 0;
        EntitySlot slot;
        slot.alive = true;
        slot.entity = e;
        w.entities.push_back(slot);
    }
    *out = e;
    return RECS_OK;
}

int32_t create_empty_entity(recs_world& w, recs_entity* out) {
    recs_entity e;
    int32_t st = allocate_entity(w, &e);
    if (st != RECS_OK) return st;
    // store_entity_in_archetype(new_entity, EMPTY_ENTITY_ARCHETYPE_INDEX)
    st = store_entity(w, w.archetypes[0], e);
    if (st != RECS_OK) return st;
    w.entity_to_archetype_index[entity_key(e)] = 0;
    *out = e;
    return RECS_OK;
}

// World::add_components_to_entity (archetypes.rs:556-651)
int32_t add_components(recs_world& w, recs_entity e, const uint64_t* ids, const void* const* values, uint32_t n) {
    auto src_it = w.entity_to_archetype_index.find(entity_key(e));
    if (src_it == w.entity_to_archetype_index.end())
        return w.fail(RECS_ERR_WORLD, "could not find the given entity: " + entity_str(e));
    const uint32_t source_index = src_it->second;
    for (uint32_t i = 0; i < n; ++i)
        for (uint32_t j = 0; j < i; ++j)
            if (ids[i] == ids[j])
                return w.fail(RECS_ERR_WORLD, "component of same type " + std::to_string(ids[i]) +
                                                  " already exists for entity " + entity_str(e));
    // find_target_archetype(Addition) (archetypes.rs:394-472)
    const std::vector<uint64_t> source_ids = component_types(w.archetypes[source_index]);
    std::vector<uint64_t> target_ids = source_ids;
    for (uint32_t i = 0; i < n; ++i) {
        if (std::find(source_ids.begin(), source_ids.end(), ids[i]) != source_ids.end())
            return w.fail(RECS_ERR_WORLD, "a component mutation failed for one or more entities: component of same type " +
                                              std::to_string(ids[i]) + " already exists for entity " + entity_str(e));
        target_ids.push_back(ids[i]);
    }
    uint32_t target_index = 0;
    if (exactly_matching(w, target_ids, &target_index)) {
        int32_t st = move_between(w, e, target_index, source_ids);
        if (st != RECS_OK) return st;
    } else {
        int32_t st = move_to_new(w, e, source_ids, &target_index);
        if (st != RECS_OK) return st;
        std::vector<uint64_t> key = target_ids;
        std::sort(key.begin(), key.end());
        w.component_typeids_set_to_archetype_index[key] = target_index;
        Archetype& target = w.archetypes[target_index];
        for (uint32_t i = 0; i < n; ++i) {  // add_component_vec_for_any_component
            if (!target.columns.count(ids[i])) {
                ColumnStore fresh;
                fresh.elem = w.registry.at(ids[i]).elem;
                target.columns.insert(std::make_pair(ids[i], fresh));
            }
        }
        for (uint32_t i = 0; i < n; ++i) w.component_typeid_to_archetype_indices[ids[i]].insert(target_index);
    }
    Archetype& target = w.archetypes[target_index];
    for (uint32_t i = 0; i < n; ++i) {  // Archetype::add_components
        auto it = target.columns.find(ids[i]);
        if (it == target.columns.end())
            return w.fail(RECS_ERR_WORLD, "the requested component type is not present: " + std::to_string(ids[i]));
        it->second.push(values ? values[i] : nullptr);
    }
    target.version++;
    return update_after_entity_removal(w, w.archetypes[source_index], e);
}

// World::remove_component_types_from_entity (archetypes.rs:664-727)
int32_t remove_components(recs_world& w, recs_entity e, const uint64_t* ids, uint32_t n) {
    auto src_it = w.entity_to_archetype_index.find(entity_key(e));
    if (src_it == w.entity_to_archetype_index.end())
        return w.fail(RECS_ERR_WORLD, "could not find the given entity: " + entity_str(e));
    const uint32_t source_index = src_it->second;
    // find_target_archetype(Removal)
    std::vector<uint64_t> target_ids = component_types(w.archetypes[source_index]);
    for (uint32_t i = 0; i < n; ++i) {
        auto it = std::find(target_ids.begin(), target_ids.end(), ids[i]);
        if (it == target_ids.end())
            return w.fail(RECS_ERR_WORLD, "a component mutation failed for one or more entities: component of type " +
                                              std::to_string(ids[i]) + " does not exist for entity " + entity_str(e));
        target_ids.erase(it);
    }
    uint32_t target_index = 0;
    if (exactly_matching(w, target_ids, &target_index)) {
        int32_t st = move_between(w, e, target_index, target_ids);
        if (st != RECS_OK) return st;
    } else {
        int32_t st = move_to_new(w, e, target_ids, &target_index);
        if (st != RECS_OK) return st;
        std::vector<uint64_t> key = target_ids;
        std::sort(key.begin(), key.end());
        w.component_typeids_set_to_archetype_index[key] = target_index;
    }
    Archetype& source = w.archetypes[source_index];
    const uint64_t source_component_index = source.entity_to_component_index.at(entity_key(e));
    for (uint32_t i = 0; i < n; ++i) source.columns.at(ids[i]).swap_remove(source_component_index, nullptr);
    source.version++;
    return update_after_entity_removal(w, source, e);
}

// World::delete_entities for one entity (lib.rs:728-768)
int32_t delete_entity(recs_world& w, recs_entity e) {
    // delete_entity_index
    if (e.id >= w.entities.size()) return w.fail(RECS_ERR_WORLD, "could not find the given entity: " + entity_str(e));
    if (!w.entities[e.id].alive) return w.fail(RECS_ERR_WORLD, entity_str(e) + " has already been deleted");
    w.entities[e.id].alive = false;
    w.deleted_entities.push_front(w.entities[e.id].entity);
    // delete_entity_components -> Archetype::remove_entity (archetypes.rs:235-246)
    auto it = w.entity_to_archetype_index.find(entity_key(e));
    if (it == w.entity_to_archetype_index.end())
        return w.fail(RECS_ERR_WORLD, "could not find the given entity: " + entity_str(e));
    Archetype& a = w.archetypes[it->second];
    auto ci = a.entity_to_component_index.find(entity_key(e));
    if (ci == a.entity_to_component_index.end())
        return w.fail(RECS_ERR_WORLD, "could not find the given entity: " + entity_str(e));
    const uint64_t index = ci->second;
    for (auto& kv : a.columns) kv.second.swap_remove(index, nullptr);
    return update_after_entity_removal(w, a, e);
}

}  // namespace

// World::get_archetype_indices (lib.rs:842-865)
std::set<uint32_t> world_archetype_indices(const recs_world& w, const uint64_t* signature, uint32_t n) {
    std::set<uint32_t> result;
    if (n == 0) {
        for (uint32_t i = 0; i < w.archetypes.size(); ++i) result.insert(i);
        return result;
    }
    for (uint32_t i = 0; i < n; ++i) {
        auto it = w.component_typeid_to_archetype_indices.find(signature[i]);
        if (it == w.component_typeid_to_archetype_indices.end()) return std::set<uint32_t>();
        if (i == 0) {
            result = it->second;
        } else {
            std::set<uint32_t> next;
            for (uint32_t a : result)
                if (it->second.count(a)) next.insert(a);
            result.swap(next);
        }
    }
    return result;
}

}  // namespace host
}  // namespace recs

#define WORLD_ENTER(w) \
    if (!(w)) return RECS_ERR_INVALID_ARGUMENT;

extern "C" {

int32_t recs_world_create(recs_world** out) {
    if (!out) return RECS_ERR_INVALID_ARGUMENT;
    *out = new recs_world();
    return RECS_OK;
}

int32_t recs_world_destroy(recs_world* w) {
    WORLD_ENTER(w);
    delete w;
    return RECS_OK;
}

const char* recs_world_last_error(recs_world* w) { return w ? w->error.c_str() : "null world"; }

int32_t recs_world_register_component(recs_world* w, uint64_t id, uint32_t elem, const char* name) {
    WORLD_ENTER(w);
    auto it = w->registry.find(id);
    if (it != w->registry.end() && it->second.elem != elem)
        return w->fail(RECS_ERR_INVALID_ARGUMENT, "component registered twice with different sizes");
    ComponentInfo info;
    info.elem = elem;
    info.name = name ? name : ("component#" + std::to_string(id));
    w->registry[id] = info;
    return RECS_OK;
}

int32_t recs_world_create_empty_entity(recs_world* w, recs_entity* out) {
    WORLD_ENTER(w);
    recs_entity e;
    int32_t st = create_empty_entity(*w, &e);
    if (st == RECS_OK && out) *out = e;
    return st;
}

int32_t recs_world_create_entity(recs_world* w, const uint64_t* ids, const void* const* values, uint32_t n,
                                 recs_entity* out) {
    WORLD_ENTER(w);
    int32_t st = check_registered(*w, ids, n);
    if (st != RECS_OK) return st;
    recs_entity e;
    st = create_empty_entity(*w, &e);  // World::create_entity (lib.rs:720-726)
    if (st != RECS_OK) return st;
    st = add_components(*w, e, ids, values, n);
    if (st == RECS_OK && out) *out = e;
    return st;
}

int32_t recs_world_create_entities(recs_world* w, uint64_t count, const uint64_t* ids, const void* const* columns,
                                   uint32_t n, recs_entity* out) {
    WORLD_ENTER(w);
    int32_t st = check_registered(*w, ids, n);
    if (st != RECS_OK) return st;
    std::vector<const void*> values(n);
    std::vector<uint32_t> elem(n);
    for (uint32_t i = 0; i < n; ++i) elem[i] = w->registry.at(ids[i]).elem;
    bool distinct = n > 0;
    for (uint32_t i = 0; i < n; ++i)
        for (uint32_t j = 0; j < i; ++j) distinct = distinct && ids[i] != ids[j];
    const std::vector<uint64_t> id_set(ids, ids + n);
    for (uint64_t k = 0; k < count; ++k) {
        // Batch path, once the archetype of exactly this component set exists (the first entity of a new set creates it the
        // long way).  Entity by entity the reference stores the new entity in the empty-entity archetype and moves it out
        // again (lib.rs:720-726, archetypes.rs:556-651): there it is pushed last and swap-removed as the last element, which
        // leaves that archetype as it was.  What remains per entity is the id allocation, its row in the target archetype and
        // the two index maps; the component values of all remaining entities are appended column by column.  Same end state
        // (tests/test_world.py::test_batch_create_equals_single_creates).
        uint32_t target = 0;
        if (distinct && exactly_matching(*w, id_set, &target) && target != 0) {
            const uint64_t m = count - k;
            {
                Archetype& t = w->archetypes[target];
                t.entities.reserve(t.entities.size() + m);
                t.entity_to_component_index.reserve(t.entity_to_component_index.size() + m);
            }
            for (uint64_t r = 0; r < m; ++r) {
                recs_entity e;
                st = allocate_entity(*w, &e);
                if (st != RECS_OK) return st;
                Archetype& t = w->archetypes[target];
                const uint64_t index = t.entity_to_component_index.size();
                if (!t.entity_to_component_index.insert(std::make_pair(entity_key(e), index)).second)
                    return w->fail(RECS_ERR_WORLD, "archetype already contains entity: " + entity_str(e));
                t.entities.push_back(e);
                w->entity_to_archetype_index[entity_key(e)] = target;
                if (out) out[k + r] = e;
            }
            Archetype& t = w->archetypes[target];
            for (uint32_t i = 0; i < n; ++i) {
                ColumnStore& col = t.columns.at(ids[i]);
                if (col.elem) {
                    const size_t old = col.data.size();
                    col.data.resize(old + (size_t)m * col.elem);
                    if (columns && columns[i])
                        memcpy(col.data.data() + old, static_cast<const uint8_t*>(columns[i]) + k * elem[i], (size_t)m * col.elem);
                    else
                        memset(col.data.data() + old, 0, (size_t)m * col.elem);
                }
                col.len += m;
            }
            w->archetypes[0].version += 3 * m;  // store, move out, update_after_entity_removal -- per entity
            t.version += 3 * m;                 // move in, store, add_components
            return RECS_OK;
        }
        for (uint32_t i = 0; i < n; ++i)
            values[i] = columns && columns[i] ? static_cast<const uint8_t*>(columns[i]) + k * elem[i] : nullptr;
        recs_entity e;
        st = create_empty_entity(*w, &e);
        if (st != RECS_OK) return st;
        st = add_components(*w, e, ids, values.data(), n);
        if (st != RECS_OK) return st;
        if (out) out[k] = e;
    }
    return RECS_OK;
}

int32_t recs_world_delete_entity(recs_world* w, recs_entity e) {
    WORLD_ENTER(w);
    return delete_entity(*w, e);
}

int32_t recs_world_add_components(recs_world* w, recs_entity e, const uint64_t* ids, const void* const* values,
                                  uint32_t n) {
    WORLD_ENTER(w);
    int32_t st = check_registered(*w, ids, n);
    if (st != RECS_OK) return st;
    return add_components(*w, e, ids, values, n);
}

int32_t recs_world_remove_components(recs_world* w, recs_entity e, const uint64_t* ids, uint32_t n) {
    WORLD_ENTER(w);
    return remove_components(*w, e, ids, n);
}

int32_t recs_world_archetype_count(recs_world* w, uint32_t* out) {
    WORLD_ENTER(w);
    if (!out) return RECS_ERR_INVALID_ARGUMENT;
    *out = static_cast<uint32_t>(w->archetypes.size());
    return RECS_OK;
}

int32_t recs_world_entity_count(recs_world* w, uint64_t* out_alive, uint64_t* out_slots) {
    WORLD_ENTER(w);
    uint64_t alive = 0;
    for (auto& s : w->entities) alive += s.alive ? 1 : 0;
    if (out_alive) *out_alive = alive;
    if (out_slots) *out_slots = w->entities.size();
    return RECS_OK;
}

int32_t recs_world_entity_at(recs_world* w, uint32_t slot, recs_entity* out, int32_t* out_alive) {
    WORLD_ENTER(w);
    if (slot >= w->entities.size()) return w->fail(RECS_ERR_INVALID_ARGUMENT, "entity slot out of range");
    if (out) *out = w->entities[slot].entity;
    if (out_alive) *out_alive = w->entities[slot].alive ? 1 : 0;
    return RECS_OK;
}

int32_t recs_world_entity_archetype(recs_world* w, recs_entity e, uint32_t* out) {
    WORLD_ENTER(w);
    auto it = w->entity_to_archetype_index.find(entity_key(e));
    if (it == w->entity_to_archetype_index.end() || !out)
        return w->fail(RECS_ERR_WORLD, "could not find the given entity");
    *out = it->second;
    return RECS_OK;
}

int32_t recs_world_entity_component_index(recs_world* w, recs_entity e, uint64_t* out) {
    WORLD_ENTER(w);
    auto it = w->entity_to_archetype_index.find(entity_key(e));
    if (it == w->entity_to_archetype_index.end() || !out)
        return w->fail(RECS_ERR_WORLD, "could not find the given entity");
    const Archetype& a = w->archetypes[it->second];
    auto ci = a.entity_to_component_index.find(entity_key(e));
    if (ci == a.entity_to_component_index.end())
        return w->fail(RECS_ERR_WORLD, "could not find component index of entity");
    *out = ci->second;
    return RECS_OK;
}

int32_t recs_world_archetype_entities(recs_world* w, uint32_t arch, recs_entity* out, uint64_t cap, uint64_t* out_n) {
    WORLD_ENTER(w);
    if (arch >= w->archetypes.size()) return w->fail(RECS_ERR_WORLD, "could not find archetype with the given index");
    const Archetype& a = w->archetypes[arch];
    if (out_n) *out_n = a.entities.size();
    if (out)
        for (uint64_t i = 0; i < a.entities.size() && i < cap; ++i) out[i] = a.entities[i];
    return RECS_OK;
}

int32_t recs_world_archetype_components(recs_world* w, uint32_t arch, uint64_t* out, uint32_t cap, uint32_t* out_n) {
    WORLD_ENTER(w);
    if (arch >= w->archetypes.size()) return w->fail(RECS_ERR_WORLD, "could not find archetype with the given index");
    const std::vector<uint64_t> ids = component_types(w->archetypes[arch]);
    if (out_n) *out_n = static_cast<uint32_t>(ids.size());
    if (out)
        for (uint32_t i = 0; i < ids.size() && i < cap; ++i) out[i] = ids[i];
    return RECS_OK;
}

int32_t recs_world_column(recs_world* w, uint32_t arch, uint64_t comp, void** out_ptr, uint32_t* out_elem,
                          uint64_t* out_count) {
    WORLD_ENTER(w);
    if (arch >= w->archetypes.size()) return w->fail(RECS_ERR_WORLD, "could not find archetype with the given index");
    Archetype& a = w->archetypes[arch];
    auto it = a.columns.find(comp);
    if (it == a.columns.end()) return w->fail(RECS_ERR_MISSING_PARAMETER, "the requested component type is not present");
    if (out_ptr) *out_ptr = it->second.data.data();
    if (out_elem) *out_elem = it->second.elem;
    if (out_count) *out_count = it->second.len;
    return RECS_OK;
}

int32_t recs_world_borrow_column(recs_world* w, uint32_t arch, uint64_t comp, int32_t write) {
    WORLD_ENTER(w);
    if (arch >= w->archetypes.size()) return w->fail(RECS_ERR_WORLD, "could not find archetype with the given index");
    auto it = w->archetypes[arch].columns.find(comp);
    if (it == w->archetypes[arch].columns.end())
        return w->fail(RECS_ERR_MISSING_PARAMETER, "the requested component type is not present");
    ColumnStore& c = it->second;
    // RwLock::try_write / try_read failing -> panic_locked_component_vec (archetypes.rs:145-151)
    if (c.writer || (write && c.readers > 0)) {
        auto r = w->registry.find(comp);
        const std::string name = r == w->registry.end() ? std::to_string(comp) : r->second.name;
        return w->fail(RECS_ERR_BORROW_CONFLICT, "Lock of ComponentVec<" + name + "> is already taken!");
    }
    if (write)
        c.writer = true;
    else
        c.readers++;
    return RECS_OK;
}

int32_t recs_world_release_column(recs_world* w, uint32_t arch, uint64_t comp, int32_t write) {
    WORLD_ENTER(w);
    if (arch >= w->archetypes.size()) return w->fail(RECS_ERR_WORLD, "could not find archetype with the given index");
    auto it = w->archetypes[arch].columns.find(comp);
    if (it == w->archetypes[arch].columns.end())
        return w->fail(RECS_ERR_MISSING_PARAMETER, "the requested component type is not present");
    ColumnStore& c = it->second;
    if (write) {
        if (!c.writer) return w->fail(RECS_ERR_INVALID_ARGUMENT, "column is not write-borrowed");
        c.writer = false;
    } else {
        if (c.readers <= 0) return w->fail(RECS_ERR_INVALID_ARGUMENT, "column is not read-borrowed");
        c.readers--;
    }
    return RECS_OK;
}

int32_t recs_world_get_archetype_indices(recs_world* w, const uint64_t* signature, uint32_t n, uint32_t* out,
                                         uint32_t cap, uint32_t* out_n) {
    WORLD_ENTER(w);
    const std::set<uint32_t> s = world_archetype_indices(*w, signature, n);
    if (out_n) *out_n = static_cast<uint32_t>(s.size());
    uint32_t i = 0;
    if (out)
        for (uint32_t a : s) {
            if (i >= cap) break;
            out[i++] = a;
        }
    return RECS_OK;
}

int32_t recs_world_archetype_version(recs_world* w, uint32_t arch, uint64_t* out) {
    WORLD_ENTER(w);
    if (arch >= w->archetypes.size() || !out) return w->fail(RECS_ERR_WORLD, "could not find archetype with the given index");
    *out = w->archetypes[arch].version;
    return RECS_OK;
}

}  // extern "C"
