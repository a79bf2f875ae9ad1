/*
 * recs_gpu.h -- C ABI of librecs_gpu.so: the B200 (sm_100a) data-parallel hot path of the
 * recs ECS engine (n-body systems, rain streaming systems, plain query iteration) behind the
 * reference's own System / Schedule / Executor plugin contract.
 *
 * The reference (martinjonsson01/recs) has no FFI of its own; every entry point below cites the
 * Rust interface (file:line under the reference checkout) that a thin `impl System` shim on
 * the Rust side would forward to.  INTEGRATION.md shows that shim.
 *
 * Conventions
 *   - every function returns an int32 status (RECS_OK == 0); nothing throws or aborts across
 *     the ABI.  Errors are sticky per context until recs_gpu_clear_error().
 *   - every function may be called from any host thread (the reference runs a system on an
 *     arbitrary worker thread each tick, crates/scheduler/src/executor/worker.rs:143-154); the
 *     context serialises calls internally and selects its device on every call.
 *   - plain pointers and sizes only; component columns are described as (archetype index,
 *     component id, element size, element count) exactly like the reference's
 *     `Archetype::component_typeid_to_component_vec` (crates/ecs/src/archetypes.rs:178-182).
 *   - there is NO CPU fallback: without a CUDA device every compute call fails with
 *     RECS_ERR_NO_DEVICE.
 */
#ifndef RECS_GPU_H
#define RECS_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RECS_GPU_ABI_VERSION 1

/* ------------------------------------------------------------------------------------------
 * Status codes.  Mirrors SystemError / ExecutionError
 * (crates/ecs/src/systems/mod.rs:28-45, crates/ecs/src/lib.rs:419-432).
 * ---------------------------------------------------------------------------------------- */
enum {
    RECS_OK = 0,
    RECS_ERR_INVALID_ARGUMENT = 1,  /* null pointer, bad enum, bad handle                      */
    RECS_ERR_NO_DEVICE = 2,         /* no CUDA device / driver: there is no CPU fallback       */
    RECS_ERR_CUDA = 3,              /* a CUDA runtime call failed (sticky)                     */
    RECS_ERR_NCCL = 4,              /* an NCCL call failed (sticky)                            */
    RECS_ERR_MISSING_PARAMETER = 5, /* SystemError::MissingParameter: a column the system needs
                                       is not bound for a matched archetype                    */
    RECS_ERR_STALE_BINDING = 6,     /* column bound before recs_gpu_invalidate() and not rebound*/
    RECS_ERR_SIZE_MISMATCH = 7,     /* element size / count disagree between columns           */
    RECS_ERR_BORROW_CONFLICT = 8,   /* "Lock of ComponentVec<T> is already taken!"
                                       (crates/ecs/src/archetypes.rs:145-151)                  */
    RECS_ERR_SCHEDULE = 9,          /* ScheduleError (crates/ecs/src/lib.rs:470-493)           */
    RECS_ERR_WORLD = 10,            /* WorldError (crates/ecs/src/lib.rs:600-645)              */
    RECS_ERR_UNSUPPORTED = 11,
    RECS_ERR_COMPILE = 12           /* the body of a generic GPU system did not compile (NVRTC log in last_error) */
};

typedef struct recs_gpu_ctx recs_gpu_ctx;

/* Context creation parameters.  Replaces the compile-time constants of the reference:
 * FIXED_TIME_STEP (crates/n-body/src/lib.rs:14), GRAVITATIONAL_CONSTANT (:123), f32::EPSILON
 * (:118), rain's FIXED_TIME_STEP / GRAVITY_ACCELERATION / TERMINAL_VELOCITY / WIND_ACCELERATION
 * (crates/rain-simulation/src/lib.rs:15,59-60,72).  recs_gpu_default_config() fills in exactly
 * those values. */
typedef struct recs_gpu_config {
    int32_t device;            /* CUDA device ordinal                                          */
    int32_t rank;              /* this process' shard index (0 when single GPU)                */
    int32_t world_size;        /* number of shards (1 when single GPU)                         */
    int32_t use_cuda_graph;    /* capture recs_gpu_tick chains into a CUDA graph               */
    float nbody_dt;            /* 1/20000                                                      */
    float nbody_g;             /* 6.67e-11                                                     */
    float nbody_epsilon;       /* f32::EPSILON = 1.1920929e-7                                  */
    float rain_dt;             /* 1/20                                                         */
    float rain_gravity;        /* 9.82                                                         */
    float rain_terminal;       /* 9.0                                                          */
    float rain_wind_accel;     /* 10.0                                                         */
    float rain_wind_deg_per_s; /* 10.0 (WIND_ROTATION_SPEED)                                   */
    float rain_ground_height;  /* 0.0  (GROUND_HEIGHT, crates/rain-simulation/src/lib.rs:150)  */
    float nbody_collision_radius; /* 0.1 (COLLISION_RADIUS, crates/n-body/examples/n_body_demo.rs:16) */
} recs_gpu_config;

void recs_gpu_default_config(recs_gpu_config* cfg);

int32_t recs_gpu_abi_version(void);

/* Lifecycle.  Replaces `E::default()` / `Drop for WorkerPool`
 * (crates/ecs/src/lib.rs:386-400, crates/scheduler/src/executor/mod.rs:84-124). */
int32_t recs_gpu_create(const recs_gpu_config* cfg, recs_gpu_ctx** out_ctx);
int32_t recs_gpu_destroy(recs_gpu_ctx* ctx);
const char* recs_gpu_last_error(recs_gpu_ctx* ctx); /* never NULL; "" when no error           */
int32_t recs_gpu_clear_error(recs_gpu_ctx* ctx);
int32_t recs_gpu_synchronize(recs_gpu_ctx* ctx);

/* ------------------------------------------------------------------------------------------
 * Component columns: device-resident mirrors of `RwLock<Vec<T>>`
 * (crates/ecs/src/archetypes.rs:15, 208-263; World::borrow_component_vecs(_mut),
 * crates/ecs/src/lib.rs:807-833).
 *
 * The host keeps ownership of `host_ptr`; it must stay valid until the next bind / invalidate
 * for that column (the Rust shim holds the column's write guard for that long).  The device
 * copy has the SAME byte layout as the host Vec<T> (Position = 3 x f32, 12 bytes).
 * ---------------------------------------------------------------------------------------- */
int32_t recs_gpu_bind_column(recs_gpu_ctx* ctx, uint32_t archetype_index, uint64_t component_id,
                             void* host_ptr, uint32_t elem_size, uint64_t count);
/* host -> device / device -> host for one bound column (async on the context stream unless
 * RECS_GPU_SYNC is passed to the system call that follows; download always completes before it
 * returns). */
int32_t recs_gpu_upload(recs_gpu_ctx* ctx, uint32_t archetype_index, uint64_t component_id);
int32_t recs_gpu_download(recs_gpu_ctx* ctx, uint32_t archetype_index, uint64_t component_id);
/* The same for rows [row_begin, row_end) only.  In an index-sharded run a rank owns the rows
 * recs_gpu_shard_rows returns: it needs no other row of Velocity / Acceleration at all, and the other ranks'
 * positions and masses reach it through the all-gather, not through its host. */
int32_t recs_gpu_upload_rows(recs_gpu_ctx* ctx, uint32_t archetype_index, uint64_t component_id,
                             uint64_t row_begin, uint64_t row_end);
int32_t recs_gpu_download_rows(recs_gpu_ctx* ctx, uint32_t archetype_index, uint64_t component_id,
                               uint64_t row_begin, uint64_t row_end);
/* Structural change happened on the host (command-buffer playback,
 * crates/ecs/src/lib.rs:321-323): all bindings of that archetype become stale. */
int32_t recs_gpu_invalidate(recs_gpu_ctx* ctx, uint32_t archetype_index);
/* Structural changes WITHOUT a round trip (command-buffer playback against resident columns,
 * crates/ecs/src/systems/command_buffers.rs:377-435): the host World stays the authority on entity ids and
 * row indices and replays on the device what it did to its own columns.
 *  - recs_gpu_column_append: `Vec::push` of new rows (archetypes.rs:200-231).  Rows [old count, new_count) are
 *    copied from host_ptr (the column's possibly re-allocated base); earlier rows are NOT touched, the device
 *    buffer grows geometrically with a device-to-device copy.
 *  - recs_gpu_archetype_move_rows: a batch of `swap_remove`s (archetypes.rs:233-243) folded into
 *    row[dst[k]] = row[src[k]] (sources as they were before the batch), applied to the listed resident columns
 *    of the archetype, which then have new_count rows.
 *  - recs_gpu_column_rebind_host: the host Vec moved (re-allocation) but its contents did not change.
 *  - recs_gpu_get_transfer_bytes: column bytes copied host->device / device->host so far (tests assert that a
 *    tick with structural changes moves only the new rows). */
int32_t recs_gpu_column_append(recs_gpu_ctx* ctx, uint32_t archetype_index, uint64_t component_id,
                               void* host_ptr, uint64_t new_count);
int32_t recs_gpu_archetype_move_rows(recs_gpu_ctx* ctx, uint32_t archetype_index,
                                     const uint64_t* component_ids, uint32_t n_components,
                                     const uint32_t* dst_rows, const uint32_t* src_rows, uint32_t n_moves,
                                     uint64_t new_count);
int32_t recs_gpu_column_rebind_host(recs_gpu_ctx* ctx, uint32_t archetype_index, uint64_t component_id,
                                    void* host_ptr);
int32_t recs_gpu_get_transfer_bytes(recs_gpu_ctx* ctx, uint64_t* out_h2d_bytes, uint64_t* out_d2h_bytes);
/* Raw device pointer of a column (for zero-copy interop, e.g. torch / NCCL plumbing). */
int32_t recs_gpu_column_device_ptr(recs_gpu_ctx* ctx, uint32_t archetype_index,
                                   uint64_t component_id, void** out_dev_ptr, uint64_t* out_count);

/* ------------------------------------------------------------------------------------------
 * GPU systems.  A GPU system is the device-side body of one reference system function plus the
 * `System` trait data the schedule needs (crates/ecs/src/systems/mod.rs:48-62).
 * ---------------------------------------------------------------------------------------- */
typedef enum recs_gpu_system_kind {
    /* crates/n-body/src/lib.rs:104-135 -- gravity(Read<Position>, Write<Acceleration>,
     *                                             Query<(Read<Position>, Read<Mass>)>)
     * components = {Position, Acceleration, Mass}                                            */
    RECS_SYS_NBODY_GRAVITY = 1,
    /* crates/n-body/src/lib.rs:87-93 -- movement(Write<Position>, Read<Velocity>)
     * components = {Position, Velocity}                                                      */
    RECS_SYS_NBODY_MOVEMENT = 2,
    /* crates/n-body/src/lib.rs:95-102 -- acceleration(Write<Velocity>, Read<Acceleration>)
     * components = {Velocity, Acceleration}                                                  */
    RECS_SYS_NBODY_ACCELERATION = 3,
    /* crates/rain-simulation/src/lib.rs:65-70 -- gravity(Write<Velocity>); components = {Velocity} */
    RECS_SYS_RAIN_GRAVITY = 4,
    /* crates/rain-simulation/src/lib.rs:86-92 -- wind(Write<Velocity>); components = {Velocity}    */
    RECS_SYS_RAIN_WIND = 5,
    /* crates/rain-simulation/src/lib.rs:94-96 -- movement(Write<Position>, Read<Velocity>)          */
    RECS_SYS_RAIN_MOVEMENT = 6,
    /* crates/rain-simulation/src/lib.rs:78-84 -- wind_direction(): no component access; rotates
     * the context's wind direction (host side, passed to the wind kernel as an argument)     */
    RECS_SYS_RAIN_WIND_DIRECTION = 7,
    /* crates/ecs_bench_suite/src/recs/simple_iter.rs:47-49 -- |Write<Position>, Read<Velocity>|
     * position.0 += velocity.0; also the generic (Write<A>, Read<B>) f32-vector par-iter query.
     * components = {A (written), B (read)}                                                   */
    RECS_SYS_QUERY_ADD = 8,
    /* crates/n-body/examples/n_body_demo.rs:54-71 -- collision(Entity, Read<Position>,
     * Query<(Entity, Read<Position>)>, Commands): an entity is removed when another entity of the
     * query is closer than COLLISION_RADIUS.  components = {Position}.  The system only DECIDES;
     * the removals are fetched with recs_gpu_system_removals and played back on the host World. */
    RECS_SYS_NBODY_COLLISION = 9,
    /* crates/rain-simulation/src/lib.rs:152-156 -- ground_collision(Entity, Read<Position>, Commands):
     * removed when position.y <= GROUND_HEIGHT.  components = {Position}                          */
    RECS_SYS_RAIN_GROUND_COLLISION = 10,
    /* a generic system created with recs_gpu_system_create_kernel (not accepted by recs_gpu_system_create) */
    RECS_SYS_KERNEL = 11
} recs_gpu_system_kind;

/* One entry of `System::component_accesses()` (crates/ecs/src/systems/mod.rs:88-94). */
typedef struct recs_gpu_access {
    uint64_t component_id;
    uint8_t is_write;
} recs_gpu_access;

#define RECS_GPU_MAX_ACCESSES 8
/* What the schedule needs to know about a system: `System::name()` and
 * `System::component_accesses()` in parameter order (crates/ecs/src/systems/mod.rs:973-978). */
typedef struct recs_gpu_system_desc {
    const char* name; /* owned by the context */
    uint32_t n_access;
    recs_gpu_access access[RECS_GPU_MAX_ACCESSES];
} recs_gpu_system_desc;

/* Registers a GPU system.  `component_ids` are the host's ids (TypeId stand-ins) for the
 * components listed next to the kind above, in that order. */
int32_t recs_gpu_system_create(recs_gpu_ctx* ctx, recs_gpu_system_kind kind,
                               const uint64_t* component_ids, uint32_t n_component_ids,
                               uint32_t* out_system_id);
int32_t recs_gpu_system_describe(recs_gpu_ctx* ctx, uint32_t system_id, recs_gpu_system_desc* out);
/* Result of `SystemParameters::get_archetype_indices` (crates/ecs/src/systems/mod.rs:986-1000)
 * computed by the host for the system's outer parameters, in iteration order. */
int32_t recs_gpu_system_set_archetypes(recs_gpu_ctx* ctx, uint32_t system_id,
                                       const uint32_t* archetype_indices, uint32_t n);
/* Same for the nested `Query<...>` parameter of gravity (its own archetype match,
 * crates/ecs/src/systems/mod.rs:1042). */
int32_t recs_gpu_system_set_query_archetypes(recs_gpu_ctx* ctx, uint32_t system_id,
                                             const uint32_t* archetype_indices, uint32_t n);

/* The `Commands::remove(entity)` calls a removal-deciding system issued during its last run, as
 * ascending component indices (rows) of one of its outer archetypes.  Synchronises the stream. */
int32_t recs_gpu_system_removals(recs_gpu_ctx* ctx, uint32_t system_id, uint32_t archetype_index,
                                 uint32_t* out_rows, uint32_t cap, uint32_t* out_count);

/* `Commands::create(components)` calls of a generic system's last run (rain spawns its drops this way,
 * crates/rain-simulation/src/lib.rs:128-148): *out_count records of *out_record_bytes bytes each, in the order a
 * sequential run of the system would have recorded them (entity order, then call order within an entity), so
 * the entity ids the host World hands out at playback do not depend on thread timing.  out_records may be NULL
 * to ask for the count.  Synchronises the stream.  More records than the buffer holds is an error
 * (RECS_ERR_SIZE_MISMATCH): raise it with recs_gpu_system_set_create_capacity (default 65 536 per run). */
int32_t recs_gpu_system_creates(recs_gpu_ctx* ctx, uint32_t system_id, void* out_records, uint32_t cap_records,
                                uint32_t* out_count, uint32_t* out_record_bytes);
int32_t recs_gpu_system_set_create_capacity(recs_gpu_ctx* ctx, uint32_t system_id, uint32_t max_records);

/* ------------------------------------------------------------------------------------------
 * Generic GPU systems: any per-entity function over `Read<T>` / `Write<T>` / `Entity` parameters
 * (crates/ecs/src/systems/mod.rs:457-566 Read, 627-739 Write, 868-912 Entity), iterated over the
 * concatenation of its matched archetypes' columns exactly once per entity -- the contract of
 * `SequentiallyIterable::run` / `SegmentIterable::segments` (crates/ecs/src/systems/iteration.rs:26-57,
 * 151-245; segments may span archetypes, crates/ecs/src/systems/mod.rs:473-540, 645-710).
 *
 * The body is CUDA C++ source defining the component types and one function
 *     void body(Write params as T&, Read/Entity params as const T&, Commands as const recs_commands&, ...
 *               [, const Uniform& u]);
 * in parameter order.  It is compiled for sm_100a with NVRTC when the system is registered
 * (--fmad=false: rustc never contracts a*b+c), wrapped in a kernel that finds each entity's
 * archetype through a prefix table and, when the rows fit, stages every column chunk through
 * shared memory so that HBM sees only 128-bit accesses whatever sizeof(T) is.  Archetype
 * membership (filters included) stays with the host: recs_gpu_system_set_archetypes.
 * ---------------------------------------------------------------------------------------- */
#define RECS_KPARAM_READ 0u
#define RECS_KPARAM_WRITE 1u
#define RECS_KPARAM_ENTITY 2u
#define RECS_KPARAM_COMMANDS 3u /* `Commands` (crates/ecs/src/systems/command_buffers.rs:53-112): the body receives
                                 * `const recs_commands&` and may call .remove() for its own entity (rows fetched with
                                 * recs_gpu_system_removals like for the built-in removal-deciding kinds) and, when the
                                 * parameter names a record type (type_name, elem_size = sizeof), .create(record): the
                                 * tuple of components of a new entity as one POD struct, fetched with
                                 * recs_gpu_system_creates and played back by the host World */
/* Component id under which the host binds an archetype's entity column for `Entity` parameters:
 * one uint64 per row, generation << 32 | id, the value `Entity::hash` feeds
 * (crates/ecs/src/lib.rs:918-927). */
#define RECS_GPU_ENTITY_COMPONENT 0xFFFFFFFFFFFFFFFEull
typedef struct recs_gpu_kernel_param {
    uint64_t component_id; /* ignored for RECS_KPARAM_ENTITY (RECS_GPU_ENTITY_COMPONENT is used)  */
    uint32_t elem_size;    /* sizeof(T) on the host == sizeof(type_name) in `source`              */
    uint32_t kind;         /* RECS_KPARAM_READ / _WRITE / _ENTITY / _COMMANDS (type_name + elem_size: the create()
                            * record, or NULL / 0) / _QUERY_READ / _QUERY_ENTITY (pair kernels only)          */
    const char* type_name; /* C++ type of this parameter in `source`                              */
} recs_gpu_kernel_param;
/* Registers a generic system.  `uniform_type` (may be NULL) names a POD struct of at most 256 bytes
 * passed by value to every call, set with recs_gpu_system_set_uniform (the per-tick scalars a Rust
 * closure would capture, e.g. a time step or the wind direction).  Compile errors return
 * RECS_ERR_COMPILE with the NVRTC log in recs_gpu_last_error. */
int32_t recs_gpu_system_create_kernel(recs_gpu_ctx* ctx, const char* name, const char* source,
                                      const char* body, const recs_gpu_kernel_param* params,
                                      uint32_t n_params, const char* uniform_type,
                                      uint32_t uniform_bytes, uint32_t* out_system_id);
int32_t recs_gpu_system_set_uniform(recs_gpu_ctx* ctx, uint32_t system_id, const void* data,
                                    uint32_t bytes);
/* Generic systems with a nested `Query<(...)>` parameter (crates/ecs/src/systems/mod.rs:915-934, 1028-1140) -- the
 * shape of the reference's `gravity` (crates/n-body/src/lib.rs:104-135) and `collision`
 * (crates/n-body/examples/n_body_demo.rs:54-71).  Parameters of kind RECS_KPARAM_QUERY_READ / _QUERY_ENTITY describe the
 * nested query (read-only, as `supports_parallelization` demands, systems/mod.rs:1101-1105); its archetypes are set
 * with recs_gpu_system_set_query_archetypes.  `source` defines an accumulator type and two functions:
 *     void pair(ACC& acc, <outer Read / Entity params as const T&>, <query params as const T&> [, const U& u]);
 *     void finish(<all outer params: Write as T&, Read / Entity as const T&, Commands as const recs_commands&>,
 *                 const ACC& acc [, const U& u]);
 * For every outer entity: ACC acc{}; pair(...) for every query entity IN QUERY ORDER (archetype order, then
 * row order); finish(...).  One thread per outer entity, query entities staged through shared memory in tiles, so a
 * body that accumulates in call order reproduces the reference's sequential fold bit for bit (--fmad=false, IEEE
 * division and square root).  In a sharded context every rank folds the whole query for the outer rows it owns; the
 * query's columns must be complete on every rank (recs_gpu_column_allgather after a sharded write).
 * Optional speculative fast fold: if `source` also defines
 *     bool <pair_fn>_fast(<the parameters of pair>);
 * -- a body that does to `acc` exactly what pair does whenever it returns true -- every tile of the nested query is
 * folded with it first and, only if some call returned false, folded again from the tile's starting accumulator with
 * pair.  The bodies see helper functions for the purpose: recs_div_rn_seq(a, b) / recs_sqrt_rn_seq(x) are `a / b` and
 * sqrtf(x) as the branch-free in-range instruction sequences, bit-identical to the operators while
 * recs_div_in_range(a) & recs_div_in_range(b) / recs_sqrt_in_range(x) hold (operands within 2^-63 .. 2^63 /
 * 2^-101 .. 2^127).  Results never depend on whether the fast body exists. */
#define RECS_KPARAM_QUERY_READ 4u
#define RECS_KPARAM_QUERY_ENTITY 5u
int32_t recs_gpu_system_create_pair_kernel(recs_gpu_ctx* ctx, const char* name, const char* source,
                                           const char* acc_type, const char* pair_fn, const char* finish_fn,
                                           const recs_gpu_kernel_param* params, uint32_t n_params,
                                           const char* uniform_type, uint32_t uniform_bytes,
                                           uint32_t* out_system_id);
int32_t recs_gpu_pair_kernel_system_compile(const char* source, const char* acc_type, const char* pair_fn,
                                            const char* finish_fn, const recs_gpu_kernel_param* params,
                                            uint32_t n_params, const char* uniform_type, char* log,
                                            uint32_t log_cap, uint64_t* out_cubin_bytes);
/* Generation + compilation only: needs neither a context nor a GPU (NVRTC cross-compiles), so a
 * host can validate its systems at build time.  `force_direct` != 0 selects the unstaged kernel
 * form.  Writes the NVRTC log (NUL terminated, truncated to log_cap), the cubin size and the
 * staged chunk size (0 = direct form). */
int32_t recs_gpu_kernel_system_compile(const char* source, const char* body,
                                       const recs_gpu_kernel_param* params, uint32_t n_params,
                                       const char* uniform_type, int32_t force_direct, char* log,
                                       uint32_t log_cap, uint64_t* out_cubin_bytes,
                                       uint32_t* out_staged_chunk);

#define RECS_GPU_ASYNC 0u /* enqueue only: every dependent is a GPU system of this context    */
#define RECS_GPU_SYNC 1u  /* return after the system's writes are visible to the host thread  */
/* `SequentiallyIterable::run(&self, world)` (crates/ecs/src/systems/iteration.rs:9-14) for a
 * GPU system: launches its kernels on the context stream.  Completion in the reference is the
 * drop of `SystemExecutionGuard::finished_sender` after `run` returns
 * (crates/ecs/src/lib.rs:525-557); with RECS_GPU_SYNC the shim may drop it on return. */
int32_t recs_gpu_run_system(recs_gpu_ctx* ctx, uint32_t system_id, uint32_t flags);

/* The per-tick order produced by the schedule (PrecedenceGraph batches flattened,
 * crates/scheduler/src/schedule/mod.rs:205-333) for the GPU systems of this context. */
int32_t recs_gpu_set_tick_order(recs_gpu_ctx* ctx, const uint32_t* system_ids, uint32_t n);
/* `Executor::execute_once` x n_ticks (crates/scheduler/src/executor/mod.rs:191-236) for a chain
 * made only of GPU systems: one stream-ordered chain (a CUDA graph when enabled).
 * When the chain is exactly gravity -> movement -> acceleration over one body archetype of at most
 * 4 096 rows (the reference's own benchmark size, crates/ecs_bench_suite/benches/n_body.rs), the
 * ticks of a call run in ONE cooperative launch that keeps the position mirror in shared memory and
 * the rows in registers (one grid barrier per tick; preceded by one tick of the ordinary chain when
 * the mirror has to be packed first); columns come out bit-identical to the tick-by-tick chain.  RECS_GPU_RESIDENT_MAX_ROWS=0 (environment) switches
 * that path off. */
int32_t recs_gpu_tick(recs_gpu_ctx* ctx, uint32_t n_ticks);

/* Rain: current wind direction (the reference's `static WIND_DIRECTION`,
 * crates/rain-simulation/src/lib.rs:75). */
int32_t recs_gpu_set_wind_direction(recs_gpu_ctx* ctx, const float xyz[3]);
int32_t recs_gpu_get_wind_direction(recs_gpu_ctx* ctx, float xyz[3]);

/* ------------------------------------------------------------------------------------------
 * Index-sharded multi-GPU n-body (one process per GPU).  Rank r owns one contiguous slice of the
 * rows of the body archetype (recs_gpu_shard_plan) and needs every body's position and mass for
 * its nested query (crates/n-body/src/lib.rs:104-135).  Two exchanges of the 16-byte
 * {position, G*m} mirror entries:
 *
 *  (a) peer push (default once recs_gpu_exchange_import succeeded): every rank's mirror lives in
 *      one device allocation that all ranks map through CUDA IPC.  The integration tail of the
 *      tick (or the pack kernel after a host upload) stores a rank's refreshed rows straight into
 *      EVERY rank's next mirror buffer over NVLink and then raises a generation flag there
 *      (release, system scope); the next tick's interaction kernel -- ONE persistent launch per
 *      tick -- starts on its own tiles at once and acquires a peer's flag only before the first
 *      tile of that peer's slice.  No collective kernel competes for SMs, nothing is exposed.
 *  (b) NCCL: ncclAllGather of the mirror on a second stream between an own-tile launch and a
 *      launch over the other ranks' tiles (needs only recs_gpu_comm_init).
 *
 * The result of a row does not depend on the exchange, the rank count or the grid: tile sums are
 * added in a fixed tree (chunks of consecutive tiles, then chunks in ascending order).
 * ---------------------------------------------------------------------------------------- */
#define RECS_GPU_NCCL_ID_BYTES 128
int32_t recs_gpu_comm_get_unique_id(recs_gpu_ctx* ctx, uint8_t id_bytes[RECS_GPU_NCCL_ID_BYTES]);
int32_t recs_gpu_comm_init(recs_gpu_ctx* ctx, const uint8_t id_bytes[RECS_GPU_NCCL_ID_BYTES]);
/* Peer push, step 1 (every rank): allocates the exchange buffers for up to `max_bodies` bodies
 * and returns the CUDA IPC handle of the allocation.  The host distributes the handles (any
 * channel: the one that carried the NCCL id will do). */
#define RECS_GPU_IPC_HANDLE_BYTES 64
int32_t recs_gpu_exchange_export(recs_gpu_ctx* ctx, uint64_t max_bodies,
                                 uint8_t handle_bytes[RECS_GPU_IPC_HANDLE_BYTES]);
/* Peer push, step 2 (every rank): `handles` = world_size handles in rank order (the own one is
 * ignored).  Maps every peer's buffers (cudaIpcOpenMemHandle, peer access over NVLink). */
int32_t recs_gpu_exchange_import(recs_gpu_ctx* ctx, const uint8_t* handles, uint32_t n_handles);
/* 0 = single GPU, 1 = NCCL all-gather, 2 = peer push. */
int32_t recs_gpu_exchange_mode(recs_gpu_ctx* ctx);
/* Delivers every rank's own rows of a column (its recs_gpu_shard_rows slice) to all ranks: grouped in-place
 * ncclBroadcasts on the context stream (needs recs_gpu_comm_init).  Row-sharded systems write only the rows a
 * rank owns; a generic system's nested Query<...> reads whole columns, so a column such a system wrote is
 * gathered before the next nested query over it.  No-op on single-GPU contexts. */
int32_t recs_gpu_column_allgather(recs_gpu_ctx* ctx, uint32_t archetype_index, uint64_t component_id);
/* Rows [row_begin, row_end) of gravity / movement / acceleration that this rank computes. */
int32_t recs_gpu_shard_rows(recs_gpu_ctx* ctx, uint64_t n_rows, uint64_t* row_begin,
                            uint64_t* row_end);
/* The same partition as a pure function (no device needed): every rank owns one slice of
 * `*slice_rows` rows of the padded body range (whole chunks of tiles); the exchange moves 16 bytes
 * per slice row and peer. */
int32_t recs_gpu_shard_plan(uint64_t n_rows, int32_t rank, int32_t world_size, uint64_t* row_begin,
                            uint64_t* row_end, uint64_t* slice_rows);

/* ------------------------------------------------------------------------------------------
 * Measurement hooks (used by bench.py; all times are CUDA-event times on the context stream).
 * ---------------------------------------------------------------------------------------- */
typedef struct recs_gpu_timing {
    float last_tick_chain_ms;    /* sum over recs_gpu_tick calls since reset (whole chains)   */
    float gravity_ms;            /* sum over launches of the interaction kernel               */
    float gravity_reduce_ms;
    float movement_ms;
    float acceleration_ms;
    float pack_ms;
    float other_ms;
    float gather_ms;             /* NCCL position all-gather (comm stream; includes waiting for peers) */
    uint64_t kernel_launches;    /* kernels launched by this library since create / reset     */
    uint64_t gravity_launches;
} recs_gpu_timing;
int32_t recs_gpu_enable_timing(recs_gpu_ctx* ctx, int32_t enabled);
int32_t recs_gpu_get_timing(recs_gpu_ctx* ctx, recs_gpu_timing* out);
int32_t recs_gpu_reset_timing(recs_gpu_ctx* ctx);
/* Generic CUDA-event bracket on the context stream: start records an event, stop records a second
 * one, waits for it and returns the elapsed device time. */
int32_t recs_gpu_timer_start(recs_gpu_ctx* ctx);
int32_t recs_gpu_timer_stop(recs_gpu_ctx* ctx, float* out_ms);
/* Issue rate of the FP32 FMA pipe measured on this GPU with a register-resident FFMA2 loop:
 * out_tflops = 2 * FMAs / s; out_sm_mhz = the SM clock derived from clock64 over the run. */
int32_t recs_gpu_measure_fp32_peak(recs_gpu_ctx* ctx, int32_t packed, float* out_tflops,
                                   float* out_sm_mhz);
/* Writes a buffer larger than L2, then streams a second one through it with loads so that no dirty
 * lines are left for the next kernel to write back (timing hygiene between iterations). */
int32_t recs_gpu_flush_l2(recs_gpu_ctx* ctx);
/* Pinned host memory for the end-to-end path. */
int32_t recs_gpu_host_alloc(recs_gpu_ctx* ctx, size_t bytes, void** out_ptr);
int32_t recs_gpu_host_free(recs_gpu_ctx* ctx, void* ptr);

#ifdef __cplusplus
}
#endif
#endif /* RECS_GPU_H */
