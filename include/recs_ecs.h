/*
 * recs_ecs.h -- C ABI of the host side of the hot path: the parts of crates/ecs and
 * crates/scheduler that decide WHICH rows a system sees and in WHAT order systems run, restated
 * in C++ (librecs_gpu.so) because no Rust toolchain exists in this environment.  In a real
 * drop-in the Rust crates keep doing this work and only include/recs_gpu.h is bound; this layer
 * exists so the GPU systems can be driven, tested and benchmarked behind the same
 * World / System / Schedule / Application contract, with bit-exact entity ids, archetype and
 * column indices, query membership, segment boundaries and per-tick system order.
 *
 * Every function returns an int32 status from recs_gpu.h (RECS_OK == 0) unless noted.
 */
#ifndef RECS_ECS_H
#define RECS_ECS_H

#include <stddef.h>
#include <stdint.h>

#include "recs_gpu.h"

#ifdef __cplusplus
extern "C" {
#endif

/* `Entity { id: u32, generation: u32 }` (crates/ecs/src/lib.rs:909-916). */
typedef struct recs_entity {
    uint32_t id;
    uint32_t generation;
} recs_entity;

/* ==========================================================================================
 * World: archetype column storage (crates/ecs/src/lib.rs:665-929, crates/ecs/src/archetypes.rs)
 * ======================================================================================== */
typedef struct recs_world recs_world;

int32_t recs_world_create(recs_world** out_world);
int32_t recs_world_destroy(recs_world* world);
const char* recs_world_last_error(recs_world* world);

/* Declares a component type: `component_id` stands in for `TypeId` (its numeric order is the
 * sort order of the `Vec<TypeId>` archetype key, archetypes.rs:753-768), `elem_size` is
 * size_of::<T>().  Zero-sized components (unit structs) are allowed. */
int32_t recs_world_register_component(recs_world* world, uint64_t component_id, uint32_t elem_size,
                                      const char* name);

/* World::create_empty_entity (archetypes.rs:338-373). */
int32_t recs_world_create_empty_entity(recs_world* world, recs_entity* out_entity);
/* World::create_entity (lib.rs:720-726): values[i] points at one value of component_ids[i]. */
int32_t recs_world_create_entity(recs_world* world, const uint64_t* component_ids, const void* const* values,
                                 uint32_t n_components, recs_entity* out_entity);
/* `count` consecutive create_entity calls; columns[i] holds `count` values of component_ids[i]
 * (Application::create_entities_with, lib.rs:190-207).  out_entities may be NULL. */
int32_t recs_world_create_entities(recs_world* world, uint64_t count, const uint64_t* component_ids,
                                   const void* const* columns, uint32_t n_components, recs_entity* out_entities);
/* World::delete_entities for one entity (lib.rs:728-768). */
int32_t recs_world_delete_entity(recs_world* world, recs_entity entity);
/* World::add_components_to_entity (archetypes.rs:556-651). */
int32_t recs_world_add_components(recs_world* world, recs_entity entity, const uint64_t* component_ids,
                                  const void* const* values, uint32_t n_components);
/* World::remove_component_types_from_entity (archetypes.rs:664-727). */
int32_t recs_world_remove_components(recs_world* world, recs_entity entity, const uint64_t* component_ids,
                                     uint32_t n_components);

/* Introspection (what the reference's unit tests look at). */
int32_t recs_world_archetype_count(recs_world* world, uint32_t* out_count);
int32_t recs_world_entity_count(recs_world* world, uint64_t* out_alive, uint64_t* out_slots);
int32_t recs_world_entity_at(recs_world* world, uint32_t slot, recs_entity* out_entity, int32_t* out_alive);
int32_t recs_world_entity_archetype(recs_world* world, recs_entity entity, uint32_t* out_archetype);
int32_t recs_world_entity_component_index(recs_world* world, recs_entity entity, uint64_t* out_index);
int32_t recs_world_archetype_entities(recs_world* world, uint32_t archetype, recs_entity* out, uint64_t cap,
                                      uint64_t* out_count);
int32_t recs_world_archetype_components(recs_world* world, uint32_t archetype, uint64_t* out_sorted_ids,
                                        uint32_t cap, uint32_t* out_count);
/* The column `Vec<T>` of one archetype: host pointer, element size, length. */
int32_t recs_world_column(recs_world* world, uint32_t archetype, uint64_t component_id, void** out_ptr,
                          uint32_t* out_elem_size, uint64_t* out_count);
/* RwLock::try_read / try_write on a column (archetypes.rs:106-151): a conflicting borrow returns
 * RECS_ERR_BORROW_CONFLICT with "Lock of ComponentVec<name> is already taken!". */
int32_t recs_world_borrow_column(recs_world* world, uint32_t archetype, uint64_t component_id, int32_t write);
int32_t recs_world_release_column(recs_world* world, uint32_t archetype, uint64_t component_id, int32_t write);
/* World::get_archetype_indices (lib.rs:842-865): archetypes containing every id of `signature`
 * (all archetypes when empty), ascending. */
int32_t recs_world_get_archetype_indices(recs_world* world, const uint64_t* signature, uint32_t n,
                                         uint32_t* out, uint32_t cap, uint32_t* out_count);
/* Structural-change counter of an archetype (bumped whenever its columns move or resize); the
 * device mirror compares it to decide when to re-bind (recs_gpu_invalidate). */
int32_t recs_world_archetype_version(recs_world* world, uint32_t archetype, uint64_t* out_version);

/* ==========================================================================================
 * System parameters (crates/ecs/src/systems/mod.rs:264-352, crates/ecs/src/filter.rs)
 * A parameter list is a prefix-notation token string, e.g.
 *   (Read<A>, Write<B>)                        -> READ A, WRITE B
 *   (Read<A>, Or<With<B>, Not<With<C>>>)        -> READ A, OR, WITH B, NOT, WITH C
 *   (Read<P>, Write<Acc>, Query<(Read<P>, Read<M>)>) -> READ P, WRITE Acc, QUERY_BEGIN, READ P, READ M, QUERY_END
 * ======================================================================================== */
typedef enum recs_param_op {
    RECS_PARAM_READ = 1,
    RECS_PARAM_WRITE = 2,
    RECS_PARAM_ENTITY = 3,
    RECS_PARAM_COMMANDS = 4,
    RECS_PARAM_WITH = 5,
    RECS_PARAM_NOT = 6,
    RECS_PARAM_AND = 7,
    RECS_PARAM_OR = 8,
    RECS_PARAM_XOR = 9,
    RECS_PARAM_QUERY_BEGIN = 10,
    RECS_PARAM_QUERY_END = 11
} recs_param_op;

typedef struct recs_param_token {
    uint32_t op;
    uint64_t component_id; /* READ / WRITE / WITH only */
} recs_param_token;

/* SystemParameters::component_accesses (mod.rs:973-978), flattened in parameter order. */
int32_t recs_params_component_accesses(const recs_param_token* tokens, uint32_t n, recs_gpu_access* out,
                                       uint32_t cap, uint32_t* out_count);
/* SystemParameters::supports_parallelization (mod.rs:980-984, 954-957, 1101-1105). */
int32_t recs_params_supports_parallelization(const recs_param_token* tokens, uint32_t n, int32_t* out_bool);
/* Whether any parameter controls iteration (else the system runs once, iteration.rs:43). */
int32_t recs_params_controls_iteration(const recs_param_token* tokens, uint32_t n, int32_t* out_bool);
/* SystemParameters::get_archetype_indices (mod.rs:986-1000) for the TOP-LEVEL parameters, or for
 * the `query_index`-th nested Query when query_index >= 0. */
int32_t recs_params_get_archetype_indices(recs_world* world, const recs_param_token* tokens, uint32_t n,
                                          int32_t query_index, uint32_t* out, uint32_t cap, uint32_t* out_count);

/* ==========================================================================================
 * Segmenting -- the par-iter split (mod.rs:394-411, 473-540; iteration.rs:191-207)
 * ======================================================================================== */
uint64_t recs_auto_segment_size(uint64_t entity_count, uint32_t thread_count);
uint64_t recs_segment_count(uint64_t entity_count, uint64_t segment_size);
/* One run of a segment: `len` items of column `column` starting at `begin`. */
typedef struct recs_segment_slice {
    uint32_t segment;
    uint32_t column; /* index into column_lengths (the matched archetypes in iteration order) */
    uint64_t begin;
    uint64_t len;
} recs_segment_slice;
/* split_borrowed_data with SegmentConfig::Size: cuts the concatenation of the columns into
 * segments of `segment_size`; slices are emitted in the order the reference pushes them.
 * segment_size == 0 means SegmentConfig::Single. */
int32_t recs_split_segments(const uint64_t* column_lengths, uint32_t n_columns, uint64_t segment_size,
                            recs_segment_slice* out, uint64_t cap, uint64_t* out_slices, uint64_t* out_segments);

/* ==========================================================================================
 * Schedule: PrecedenceGraph / OrderedPrecedenceGraph (crates/scheduler/src/schedule/mod.rs,
 * crates/scheduler/src/precedence.rs)
 * ======================================================================================== */
typedef struct recs_schedule recs_schedule;

typedef struct recs_system_info {
    const char* name;
    uint32_t n_access;
    const recs_gpu_access* access;
} recs_system_info;

/* Orderable::precedence_to (precedence.rs:39-66): -1 Before, 0 Equal, +1 After. */
int32_t recs_precedence(const recs_system_info* self, const recs_system_info* other);

/* Schedule::generate (schedule/mod.rs:116-149 incl. reduce_makespan :167-198; ordered != 0 selects
 * OrderedPrecedenceGraph :492-520). */
int32_t recs_schedule_generate(const recs_system_info* systems, uint32_t n, int32_t ordered,
                               recs_schedule** out_schedule);
int32_t recs_schedule_destroy(recs_schedule* schedule);
/* Final DAG edges in edge-index order: from[i] runs before to[i]. */
int32_t recs_schedule_edges(recs_schedule* schedule, uint32_t* from, uint32_t* to, uint32_t cap, uint32_t* out_count);

#define RECS_SCHEDULE_NEW_TICK 100    /* ScheduleError::NewTick                                 */
#define RECS_SCHEDULE_WOULD_BLOCK 101 /* the reference would block in Select (schedule/mod.rs:271-297) */
#define RECS_REACTION_RETURN_ERROR 0  /* NewTickReaction::ReturnError                           */
#define RECS_REACTION_RETURN_NEW_TICK 1
/* currently_executable_systems_with_reaction (schedule/mod.rs:151-157, 205-268).  Systems are
 * identified by their index in the `systems` array given to generate.  Completion of a system
 * (dropping its SystemExecutionGuard) is signalled with recs_schedule_complete. */
int32_t recs_schedule_next_batch(recs_schedule* schedule, int32_t reaction, uint32_t* out_systems, uint32_t cap,
                                 uint32_t* out_count);
int32_t recs_schedule_complete(recs_schedule* schedule, uint32_t system);
/* Which completed pending system the schedule acknowledges first when several have completed
 * (the reference blocks in crossbeam's Select, schedule/mod.rs:271-297):
 *   RECS_SELECT_CROSSBEAM  crossbeam-channel's thread-local xorshift shuffle, state as in a fresh
 *                          thread -- reproduces the batch sequences the reference's tests expect
 *   RECS_SELECT_FIRST      oldest pending completion first (deterministic; what recs_app uses) */
#define RECS_SELECT_CROSSBEAM 0
#define RECS_SELECT_FIRST 1
int32_t recs_schedule_set_select_policy(recs_schedule* schedule, int32_t policy);
/* Convenience: the batches of one whole tick when every dispatched system completes before the
 * next request (what a single-stream GPU chain does).  batch_of[i] = batch number of order[i]. */
int32_t recs_schedule_tick_order(recs_schedule* schedule, uint32_t* out_order, uint32_t* out_batch_of, uint32_t cap,
                                 uint32_t* out_count);

/* ==========================================================================================
 * Application: BasicApplication + run loop (crates/ecs/src/lib.rs:69-341) with an executor that
 * dispatches GPU systems on the context stream and CPU systems on the calling thread.
 * ======================================================================================== */
typedef struct recs_app recs_app;

/* A CPU system body: called once per matched archetype with the archetype's column pointers in
 * parameter order (READ/WRITE parameters only) -- the `Sequential` executor path
 * (lib.rs:439-465).  Systems without iterating parameters are called once with count == 0. */
typedef void (*recs_cpu_system_fn)(void* user, uint32_t archetype, uint64_t count, void* const* columns,
                                   uint32_t n_columns);

/* The par-iter form of a CPU system body (`SegmentIterable`, crates/ecs/src/systems/iteration.rs:151-245): called for
 * rows [row_begin, row_begin + row_count) of one archetype; `columns` are the archetype's column BASE pointers in
 * parameter order.  Under the Sequential executor it is called once per matched archetype with all rows; under the
 * WorkerPool executor the concatenation of the matched archetypes is cut into segments of SegmentSize::Auto =
 * max(16, N / (threads * 4)) entities (systems/mod.rs:394-411, 473-540), one task each, calls may come from any
 * worker thread concurrently (for disjoint row ranges). */
typedef void (*recs_cpu_segment_fn)(void* user, uint32_t archetype, uint64_t row_begin, uint64_t row_count,
                                    void* const* columns, uint32_t n_columns);

int32_t recs_app_create(recs_gpu_ctx* gpu /* may be NULL: CPU systems only */, recs_app** out_app);
int32_t recs_app_destroy(recs_app* app);
const char* recs_app_last_error(recs_app* app);
recs_world* recs_app_world(recs_app* app);
/* ApplicationBuilder::add_system (lib.rs:76-81).  Returns the registration index. */
int32_t recs_app_add_gpu_system(recs_app* app, recs_gpu_system_kind kind, const uint64_t* component_ids,
                                uint32_t n_component_ids, uint32_t* out_index);
/* A generic GPU system (recs_gpu_system_create_kernel) behind the same registration path: `tokens` is the
 * parameter list of the function -- Read / Write / Entity become the arguments of `body` in order, a Commands
 * parameter becomes `const recs_commands&` (remove() of the body's own entity; the rows feed the tick's command
 * playback like those of the built-in removal-deciding kinds), With / Not / And / Or / Xor only filter the archetype
 * match (crates/ecs/src/filter.rs) -- and type_names[i] is the C++ type of the i-th Read / Write / Entity argument in
 * `source`.  Element sizes come from the component registry. */
int32_t recs_app_add_gpu_kernel_system(recs_app* app, const char* name, const char* source, const char* body,
                                       const recs_param_token* tokens, uint32_t n_tokens, const char* const* type_names,
                                       const char* uniform_type, uint32_t uniform_bytes, uint32_t* out_index);
/* The same with `Commands::create` (crates/ecs/src/systems/command_buffers.rs:53-66): `create_components` lists the
 * components of the tuple the body hands to `commands.create(record)`; `create_record_type` names the POD struct in
 * `source` that holds them packed in that order (checked against the registry's element sizes at compile time).  The
 * records of a tick are played back with the tick's other commands, system by system in registration order, in the
 * order a sequential run would have recorded them -- rain's drop spawning (crates/rain-simulation/src/lib.rs:128-148)
 * runs on the device this way and the new rows reach the resident columns as appends. */
int32_t recs_app_add_gpu_kernel_system_creating(recs_app* app, const char* name, const char* source, const char* body,
                                                const recs_param_token* tokens, uint32_t n_tokens,
                                                const char* const* type_names, const char* uniform_type,
                                                uint32_t uniform_bytes, const uint64_t* create_components,
                                                uint32_t n_create_components, const char* create_record_type,
                                                uint32_t* out_index);
/* The nested-query form (recs_gpu_system_create_pair_kernel): `tokens` holds one Query<...> group; type_names lists the
 * C++ types of the outer Read / Write / Entity parameters and of the query's parameters in token order. */
int32_t recs_app_add_gpu_pair_kernel_system(recs_app* app, const char* name, const char* source, const char* acc_type,
                                            const char* pair_fn, const char* finish_fn, const recs_param_token* tokens,
                                            uint32_t n_tokens, const char* const* type_names, const char* uniform_type,
                                            uint32_t uniform_bytes, uint32_t* out_index);
int32_t recs_app_set_system_uniform(recs_app* app, uint32_t system_index, const void* data, uint32_t bytes);
/* The fold behind resident command playback, as a pure function: `rows[i]` is the row index passed to the i-th
 * `swap_remove` (crates/ecs/src/archetypes.rs:233-243) of a column that had len_before rows.  The result is the
 * equivalent set of moves "row dst[k] ends up holding what row src[k] held before the batch" (what
 * recs_gpu_archetype_move_rows applies) and the final length. */
int32_t recs_fold_swap_removes(uint64_t len_before, const uint32_t* rows, uint32_t n_rows, uint32_t* out_dst,
                               uint32_t* out_src, uint32_t cap, uint32_t* out_n_moves, uint64_t* out_len_after);
int32_t recs_app_add_cpu_system(recs_app* app, const char* name, const recs_param_token* tokens, uint32_t n_tokens,
                                recs_cpu_system_fn fn, void* user, uint32_t* out_index);
int32_t recs_app_add_cpu_segment_system(recs_app* app, const char* name, const recs_param_token* tokens, uint32_t n_tokens,
                                        recs_cpu_segment_fn fn, void* user, uint32_t* out_index);
/* Executor of recs_app_run: 0 = `Sequential` (lib.rs:439-465, the default), n > 0 = `WorkerPool` with n worker
 * threads (crates/scheduler/src/executor/mod.rs:101-291): the CPU systems of a schedule batch run concurrently, par-iter
 * systems as segment tasks; GPU systems stay single tasks on the context stream. */
int32_t recs_app_set_executor(recs_app* app, uint32_t worker_threads);
/* Application::run for `n_ticks` ticks (lib.rs:314-341): schedule generation on first use
 * (into_tickable, :386-400), then per tick execute_once + playback (structural commands are
 * applied by the caller between calls).  ordered != 0 selects OrderedPrecedenceGraph. */
int32_t recs_app_run(recs_app* app, uint32_t n_ticks, int32_t ordered);
/* `Commands` for CPU systems (crates/ecs/src/systems/command_buffers.rs:53-112): may be called from
 * inside a recs_cpu_system_fn; recorded and played back after the tick like
 * `CommandPlayer::playback_commands` (:377-435): all creates first (record order), then removes.
 * GPU systems that decide removals (RECS_SYS_NBODY_COLLISION, RECS_SYS_RAIN_GROUND_COLLISION) feed
 * the same playback with the rows they flagged.  Before entities are removed or created the host
 * columns are made current (recs_app_sync_to_host); the device mirror re-binds on the next tick. */
int32_t recs_app_command_create(recs_app* app, const uint64_t* component_ids, const void* const* values,
                                uint32_t n_components);
int32_t recs_app_command_remove(recs_app* app, recs_entity entity);
/* Number of entities created / removed by the playback of the last tick. */
int32_t recs_app_last_playback(recs_app* app, uint64_t* out_created, uint64_t* out_removed);
/* Order in which systems were dispatched during the last tick, and their batch numbers. */
int32_t recs_app_last_tick_order(recs_app* app, uint32_t* out_order, uint32_t* out_batch_of, uint32_t cap,
                                 uint32_t* out_count);
/* Makes the host columns current (downloads every device-dirty column). */
int32_t recs_app_sync_to_host(recs_app* app);
/* Application::add_component / remove_component (crates/ecs/src/lib.rs:213-228) between ticks.  The entity moves to
 * another archetype (crates/ecs/src/archetypes.rs:474-518, 556-651); device-resident columns of both archetypes follow
 * without a re-upload: one row per column comes down where the device held the newest copy, the source archetype
 * replays its swap_remove as a row move, the target's resident columns get the pushed row appended.  (Calling
 * recs_world_add_components directly instead leaves the mirror to the invalidate / re-bind path.) */
int32_t recs_app_add_components(recs_app* app, recs_entity entity, const uint64_t* component_ids, const void* const* values,
                                uint32_t n);
int32_t recs_app_remove_components(recs_app* app, recs_entity entity, const uint64_t* component_ids, uint32_t n);
/* Tells the app the host columns of an archetype were modified outside of systems. */
int32_t recs_app_mark_host_dirty(recs_app* app, uint32_t archetype);

#ifdef __cplusplus
}
#endif
#endif /* RECS_ECS_H */
