// BasicApplication + run loop (crates/ecs/src/lib.rs:69-341) with an executor that dispatches GPU
// systems on the context stream and CPU systems on the calling thread -- the "GPU system kind" of
// the north star behind the reference's System / Schedule / Executor contract.
//
//   * systems are registered in order (ApplicationBuilder::add_system, lib.rs:76-81);
//   * the schedule is generated from their names + component accesses exactly like
//     `S::generate(systems)` in into_tickable (lib.rs:386-400);
//   * per tick the executor asks the schedule for batches (Executor::execute_once,
//     crates/scheduler/src/executor/mod.rs:191-236) and completes each dispatched system, which for
//     a GPU system means "enqueued on the stream" (every dependent GPU system is stream-ordered
//     behind it) and for a CPU system means "returned";
//   * component columns live in the host World and are mirrored to HBM on demand: a column is
//     (re)bound when its archetype changed structurally, uploaded when the host copy is newer,
//     downloaded before a CPU system touches a column the device wrote.
#include <string.h>

#include <chrono>
#include <condition_variable>
#include <deque>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <thread>
#include <unordered_set>

#include "host.h"

using namespace recs::host;

namespace {

struct AppSystem {
    bool gpu = false;
    uint32_t gpu_id = 0;
    recs_gpu_system_kind kind = RECS_SYS_QUERY_ADD;
    std::string name;
    std::vector<Param> params;
    std::vector<recs_gpu_access> access;
    recs_cpu_system_fn fn = nullptr;
    recs_cpu_segment_fn segment_fn = nullptr;  // par-iter form: may be cut into segments (SegmentIterable)
    bool decides_removals = false;             // generic GPU system with a Commands parameter
    // generic GPU system whose Commands parameter creates entities: the components of a create() record, in record order
    std::vector<uint64_t> create_comps;
    std::vector<uint32_t> create_sizes;
    void* user = nullptr;
};

// The system whose function is running on this thread: commands are recorded into ITS buffer, and playback drains the
// buffers in system registration order (receive_all_commands, crates/ecs/src/lib.rs:325-336).
thread_local uint32_t t_current_system = 0xffffffffu;
struct CurrentSystem {
    uint32_t saved;
    explicit CurrentSystem(uint32_t index) : saved(t_current_system) { t_current_system = index; }
    ~CurrentSystem() { t_current_system = saved; }
};

// `WorkerPool` (crates/scheduler/src/executor/mod.rs:101-188, worker.rs:141-168) reduced to what the executor
// contract needs: N worker threads draining one shared injector queue of tasks; the dispatching thread blocks
// until the tasks of the current batch are done (the reference blocks on the systems' finished-channels).
class WorkerPool {
  public:
    explicit WorkerPool(uint32_t n) {
        for (uint32_t i = 0; i < n; ++i) workers_.emplace_back([this] { loop(); });
    }
    ~WorkerPool() {
        {
            std::lock_guard<std::mutex> lock(mu_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto& t : workers_) t.join();  // Drop for WorkerPool joins its threads (executor/mod.rs:84-99)
    }
    uint32_t size() const { return (uint32_t)workers_.size(); }
    void submit(std::function<void()> task) {
        {
            std::lock_guard<std::mutex> lock(mu_);
            queue_.push_back(std::move(task));
            ++pending_;
        }
        cv_.notify_one();
    }
    void wait_idle() {
        std::unique_lock<std::mutex> lock(mu_);
        done_cv_.wait(lock, [this] { return pending_ == 0; });
    }

  private:
    void loop() {
        for (;;) {
            std::function<void()> task;
            {
                std::unique_lock<std::mutex> lock(mu_);
                cv_.wait(lock, [this] { return stop_ || !queue_.empty(); });
                if (queue_.empty()) return;
                task = std::move(queue_.front());
                queue_.pop_front();
            }
            task();
            {
                std::lock_guard<std::mutex> lock(mu_);
                if (--pending_ == 0) done_cv_.notify_all();
            }
        }
    }
    std::vector<std::thread> workers_;
    std::deque<std::function<void()>> queue_;
    std::mutex mu_;
    std::condition_variable cv_, done_cv_;
    uint64_t pending_ = 0;
    bool stop_ = false;
};

struct Mirror {
    uint64_t bound_version = ~0ull;  // archetype version the device buffer was bound at
    bool host_dirty = true;          // host copy newer than device
    bool device_dirty = false;       // device copy newer than host
};

typedef std::pair<uint32_t, uint64_t> ColKey;

recs_param_token tok(uint32_t op, uint64_t comp = 0) {
    recs_param_token t;
    t.op = op;
    t.component_id = comp;
    return t;
}

// Parameter list of each GPU system kind == the signature of the reference function it replaces.
std::vector<recs_param_token> gpu_kind_tokens(recs_gpu_system_kind kind, const uint64_t* c) {
    switch (kind) {
        case RECS_SYS_NBODY_GRAVITY:  // crates/n-body/src/lib.rs:106-110
            return {tok(RECS_PARAM_READ, c[0]), tok(RECS_PARAM_WRITE, c[1]), tok(RECS_PARAM_QUERY_BEGIN),
                    tok(RECS_PARAM_READ, c[0]), tok(RECS_PARAM_READ, c[2]), tok(RECS_PARAM_QUERY_END)};
        case RECS_SYS_NBODY_MOVEMENT:      // :89
        case RECS_SYS_RAIN_MOVEMENT:       // crates/rain-simulation/src/lib.rs:94
        case RECS_SYS_NBODY_ACCELERATION:  // :97
        case RECS_SYS_QUERY_ADD:           // crates/ecs_bench_suite/src/recs/simple_iter.rs:47
            return {tok(RECS_PARAM_WRITE, c[0]), tok(RECS_PARAM_READ, c[1])};
        case RECS_SYS_RAIN_GRAVITY:  // crates/rain-simulation/src/lib.rs:65
        case RECS_SYS_RAIN_WIND:     // :86
            return {tok(RECS_PARAM_WRITE, c[0])};
        case RECS_SYS_RAIN_WIND_DIRECTION:  // :78
            return {};
        case RECS_SYS_NBODY_COLLISION:  // crates/n-body/examples/n_body_demo.rs:54-59
            return {tok(RECS_PARAM_ENTITY), tok(RECS_PARAM_READ, c[0]), tok(RECS_PARAM_QUERY_BEGIN), tok(RECS_PARAM_ENTITY),
                    tok(RECS_PARAM_READ, c[0]), tok(RECS_PARAM_QUERY_END), tok(RECS_PARAM_COMMANDS)};
        case RECS_SYS_RAIN_GROUND_COLLISION:  // crates/rain-simulation/src/lib.rs:152
            return {tok(RECS_PARAM_ENTITY), tok(RECS_PARAM_READ, c[0]), tok(RECS_PARAM_COMMANDS)};
        case RECS_SYS_KERNEL:  // generic systems bring their own parameter list (recs_app_add_gpu_kernel_system)
            return {};
    }
    return {};
}

}  // namespace

struct CreateCommand {
    std::vector<uint64_t> ids;
    std::vector<std::vector<uint8_t>> values;
    uint32_t system = 0xffffffffu;  // index of the recording system (commands recorded outside a system come last)
    // A batch: `count` entities of the same component set recorded by one GPU system in one run, component k of all of
    // them contiguous in values[k] (count == 0: a single entity, values[k] = its component k)
    uint64_t count = 0;
};

struct recs_app {
    std::vector<CreateCommand> creates;   // EntityCommand::Create, record order
    std::vector<recs_entity> removes;     // EntityCommand::Remove, record order
    std::vector<std::pair<uint32_t, std::vector<uint32_t>>> gpu_removal_sources;  // (system index, outer archetypes) run this tick
    std::vector<uint32_t> gpu_create_sources;  // generic GPU systems with a create() record run this tick
    uint64_t last_created = 0, last_removed = 0;
    // RECS_APP_PROFILE: milliseconds by phase, summed over ticks
    double prof[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // systems+flush, fetch removals, fetch creates, world creates, device appends, world removes, device moves+rebind, ticks
    recs_gpu_ctx* gpu = nullptr;
    recs_world* world = nullptr;
    std::vector<AppSystem> systems;
    recs_schedule* schedule = nullptr;
    int32_t schedule_ordered = -1;
    size_t scheduled_systems = 0;
    std::vector<uint32_t> last_order, last_batch;
    std::map<ColKey, Mirror> mirror;
    // entity columns for `Entity` parameters of generic GPU systems: generation << 32 | id per row
    struct EntityColumn {
        std::vector<uint64_t> keys;
        uint64_t version = ~0ull;
    };
    std::map<uint32_t, EntityColumn> entity_columns;
    std::unique_ptr<WorkerPool> pool;  // null: Sequential executor
    // GPU systems of the current tick that are prepared but not yet launched: consecutive GPU systems go to the
    // context as ONE chain (recs_gpu_set_tick_order + recs_gpu_tick), so its fusions apply through the Application too
    // (gravity -> movement -> acceleration in the reduce tail, wind -> gravity -> movement in one pass).
    std::vector<uint32_t> gpu_chain;
    uint32_t chains_flushed = 0;  // in the current tick
    uint32_t last_chain_len = 0;
    std::mutex command_mu;             // Commands may be recorded from several worker threads at once
    std::string error;

    int32_t fail(int32_t code, const std::string& msg) {
        error = msg;
        return code;
    }
};

namespace {

int32_t gpu_fail(recs_app* app, int32_t st) { return app->fail(st, recs_gpu_last_error(app->gpu)); }

// Makes the device copy of (archetype, component) current before a GPU system uses it.
int32_t ensure_on_device(recs_app* app, uint32_t arch, uint64_t comp) {
    recs_world* w = app->world;
    if (arch >= w->archetypes.size()) return app->fail(RECS_ERR_WORLD, "could not find archetype with the given index");
    Archetype& a = w->archetypes[arch];
    auto it = a.columns.find(comp);
    if (it == a.columns.end()) return app->fail(RECS_ERR_MISSING_PARAMETER, "could not execute system due to missing parameter");
    Mirror& m = app->mirror[ColKey(arch, comp)];
    if (m.bound_version != a.version) {
        if (m.device_dirty) return app->fail(RECS_ERR_STALE_BINDING, "archetype changed structurally while its device copy was newer; call recs_app_sync_to_host before mutating the world");
        int32_t st = recs_gpu_bind_column(app->gpu, arch, comp, it->second.data.data(), it->second.elem ? it->second.elem : 1,
                                          it->second.elem ? it->second.len : 0);
        if (st != RECS_OK) return gpu_fail(app, st);
        m.bound_version = a.version;
        m.host_dirty = true;
    }
    if (m.host_dirty) {
        int32_t st = recs_gpu_upload(app->gpu, arch, comp);
        if (st != RECS_OK) return gpu_fail(app, st);
        m.host_dirty = false;
    }
    return RECS_OK;
}

int32_t ensure_on_host(recs_app* app, uint32_t arch, uint64_t comp) {
    auto it = app->mirror.find(ColKey(arch, comp));
    if (it == app->mirror.end() || !it->second.device_dirty) return RECS_OK;
    int32_t st = recs_gpu_download(app->gpu, arch, comp);
    if (st != RECS_OK) return gpu_fail(app, st);
    it->second.device_dirty = false;
    return RECS_OK;
}

// `Entity` parameter of a generic GPU system: the archetype's entity list as a device column.
int32_t ensure_entities_on_device(recs_app* app, uint32_t arch) {
    recs_world* w = app->world;
    if (arch >= w->archetypes.size()) return app->fail(RECS_ERR_WORLD, "could not find archetype with the given index");
    const Archetype& a = w->archetypes[arch];
    recs_app::EntityColumn& col = app->entity_columns[arch];
    if (col.version == a.version) return RECS_OK;
    col.keys.resize(a.entities.size());
    for (size_t i = 0; i < a.entities.size(); ++i) col.keys[i] = entity_key(a.entities[i]);
    int32_t st = recs_gpu_bind_column(app->gpu, arch, RECS_GPU_ENTITY_COMPONENT, col.keys.data(), 8, col.keys.size());
    if (st != RECS_OK) return gpu_fail(app, st);
    if ((st = recs_gpu_upload(app->gpu, arch, RECS_GPU_ENTITY_COMPONENT)) != RECS_OK) return gpu_fail(app, st);
    col.version = a.version;
    return RECS_OK;
}

// Launches the pending chain of GPU systems in order.  Called before anything on the host may look at or change a
// column (CPU systems with component parameters, playback, sync_to_host) and at the end of every tick.
int32_t flush_gpu_chain(recs_app* app) {
    if (app->gpu_chain.empty()) return RECS_OK;
    int32_t st = recs_gpu_set_tick_order(app->gpu, app->gpu_chain.data(), (uint32_t)app->gpu_chain.size());
    if (st == RECS_OK) st = recs_gpu_tick(app->gpu, 1);
    app->chains_flushed++;
    app->last_chain_len = (uint32_t)app->gpu_chain.size();
    app->gpu_chain.clear();
    return st == RECS_OK ? RECS_OK : gpu_fail(app, st);
}

// Query membership of the system + residency of every column it touches; the launch itself is deferred (gpu_chain).
int32_t run_gpu_system(recs_app* app, AppSystem& s) {
    if (!app->gpu) return app->fail(RECS_ERR_NO_DEVICE, "GPU system registered on an application without a GPU context");
    std::vector<uint32_t> archs = params_archetypes(*app->world, s.params);
    std::vector<uint32_t> qarchs;
    for (const Param& p : s.params)
        if (p.op == RECS_PARAM_QUERY_BEGIN) qarchs = params_archetypes(*app->world, p.children);
    int32_t st;
    for (const Param& p : s.params) {
        if (p.op == RECS_PARAM_READ || p.op == RECS_PARAM_WRITE)
            for (uint32_t a : archs)
                if ((st = ensure_on_device(app, a, p.comp)) != RECS_OK) return st;
        if (p.op == RECS_PARAM_ENTITY && s.kind == RECS_SYS_KERNEL)
            for (uint32_t a : archs)
                if ((st = ensure_entities_on_device(app, a)) != RECS_OK) return st;
        if (p.op == RECS_PARAM_QUERY_BEGIN)
            for (const Param& q : p.children) {
                if (q.op == RECS_PARAM_READ || q.op == RECS_PARAM_WRITE)
                    for (uint32_t a : qarchs)
                        if ((st = ensure_on_device(app, a, q.comp)) != RECS_OK) return st;
                if (q.op == RECS_PARAM_ENTITY && s.kind == RECS_SYS_KERNEL)
                    for (uint32_t a : qarchs)
                        if ((st = ensure_entities_on_device(app, a)) != RECS_OK) return st;
            }
    }
    if ((st = recs_gpu_system_set_archetypes(app->gpu, s.gpu_id, archs.data(), (uint32_t)archs.size())) != RECS_OK)
        return gpu_fail(app, st);
    bool has_query = false;
    for (const Param& p : s.params) has_query = has_query || p.op == RECS_PARAM_QUERY_BEGIN;
    if (!qarchs.empty() || s.kind == RECS_SYS_NBODY_GRAVITY || has_query)
        if ((st = recs_gpu_system_set_query_archetypes(app->gpu, s.gpu_id, qarchs.data(), (uint32_t)qarchs.size())) != RECS_OK)
            return gpu_fail(app, st);
    app->gpu_chain.push_back(s.gpu_id);  // launched by flush_gpu_chain
    if (s.kind == RECS_SYS_NBODY_COLLISION || s.kind == RECS_SYS_RAIN_GROUND_COLLISION || s.decides_removals)
        app->gpu_removal_sources.push_back(std::make_pair((uint32_t)(&s - app->systems.data()), archs));
    if (!s.create_comps.empty()) app->gpu_create_sources.push_back((uint32_t)(&s - app->systems.data()));
    for (const Param& p : s.params)
        if (p.op == RECS_PARAM_WRITE)
            for (uint32_t a : archs) app->mirror[ColKey(a, p.comp)].device_dirty = true;
    return RECS_OK;
}

// A CPU system of the current batch, borrowed and cut into tasks.
struct CpuRun {
    AppSystem* sys = nullptr;
    std::vector<uint32_t> archs;
    std::vector<std::vector<void*>> cols;  // per matched archetype: column base pointers in parameter order
    std::vector<uint64_t> counts;
};

void release_borrows(recs_app* app, const AppSystem& s, uint32_t arch, const Param* upto) {
    for (const Param& q : s.params) {
        if (&q == upto) break;
        if (q.op == RECS_PARAM_READ || q.op == RECS_PARAM_WRITE)
            recs_world_release_column(app->world, arch, q.comp, q.op == RECS_PARAM_WRITE);
    }
}

// Borrow phase of `SequentiallyIterable::run` / `SegmentIterable::segments`
// (crates/ecs/src/systems/iteration.rs:26-57, 151-245): archetype match, host residency, RwLock try_read /
// try_write on every column for the duration of the system (archetypes.rs:106-138).
int32_t prepare_cpu_system(recs_app* app, AppSystem& s, CpuRun* run) {
    run->sys = &s;
    if (!params_controls_iteration(s.params)) return RECS_OK;
    // the host is about to read or write columns: everything the device still owes must have been enqueued, or a
    // later upload could overtake a pending kernel that has to see the old contents
    int32_t flushed = flush_gpu_chain(app);
    if (flushed != RECS_OK) return flushed;
    run->archs = params_archetypes(*app->world, s.params);
    for (uint32_t a : run->archs) {
        std::vector<void*> cols;
        for (const Param& p : s.params) {
            if (p.op != RECS_PARAM_READ && p.op != RECS_PARAM_WRITE) continue;
            int32_t st = ensure_on_host(app, a, p.comp);
            if (st != RECS_OK) return st;
            auto it = app->world->archetypes[a].columns.find(p.comp);
            if (it == app->world->archetypes[a].columns.end())
                return app->fail(RECS_ERR_MISSING_PARAMETER, "could not execute system due to missing parameter");
            st = recs_world_borrow_column(app->world, a, p.comp, p.op == RECS_PARAM_WRITE);
            if (st != RECS_OK) {
                release_borrows(app, s, a, &p);  // what was taken so far in this archetype
                for (size_t k = 0; k + 1 <= run->cols.size(); ++k) release_borrows(app, s, run->archs[k], nullptr);
                return app->fail(st, recs_world_last_error(app->world));
            }
            cols.push_back(it->second.data.data());
        }
        run->cols.push_back(cols);
        run->counts.push_back(app->world->archetypes[a].entities.size());
    }
    return RECS_OK;
}

void finish_cpu_system(recs_app* app, CpuRun& run) {
    for (size_t k = 0; k < run.cols.size(); ++k) {
        const uint32_t a = run.archs[k];
        release_borrows(app, *run.sys, a, nullptr);
        for (const Param& p : run.sys->params)
            if (p.op == RECS_PARAM_WRITE) app->mirror[ColKey(a, p.comp)].host_dirty = true;
    }
}

// The tasks of one CPU system.  Sequential executor / plain systems: one task.  WorkerPool + par-iter system: one
// task per segment of SegmentSize::Auto = max(16, N / (threads * 4)) entities (crates/ecs/src/systems/mod.rs:394-411),
// the concatenation of the matched archetypes' columns cut by split_borrowed_data (:473-540).
void cpu_system_tasks(recs_app* app, CpuRun& run, std::vector<std::function<void()>>* tasks) {
    AppSystem& s = *run.sys;
    const uint32_t self = (uint32_t)(&s - app->systems.data());
    if (!params_controls_iteration(s.params)) {
        // no iterating parameter: the function runs exactly once (iterate_once, iteration.rs:43)
        if (s.segment_fn)
            tasks->push_back([&s, self] {
                CurrentSystem cur(self);
                s.segment_fn(s.user, 0, 0, 0, nullptr, 0);
            });
        else
            tasks->push_back([&s, self] {
                CurrentSystem cur(self);
                s.fn(s.user, 0, 0, nullptr, 0);
            });
        return;
    }
    if (!s.segment_fn) {
        tasks->push_back([&s, &run, self] {
            CurrentSystem cur(self);
            for (size_t k = 0; k < run.cols.size(); ++k)
                s.fn(s.user, run.archs[k], run.counts[k], run.cols[k].data(), (uint32_t)run.cols[k].size());
        });
        return;
    }
    uint64_t total = 0;
    for (uint64_t c : run.counts) total += c;
    const uint64_t segment_size = app->pool ? recs_auto_segment_size(total, app->pool->size()) : 0;  // 0 = Single
    std::vector<recs_segment_slice> slices(run.counts.size() + (segment_size ? total / segment_size + 2 : 1) + 1);
    uint64_t n_slices = 0, n_segments = 0;
    recs_split_segments(run.counts.data(), (uint32_t)run.counts.size(), segment_size, slices.data(), slices.size(), &n_slices,
                        &n_segments);
    slices.resize(n_slices);
    for (uint64_t seg = 0; seg < n_segments; ++seg) {
        std::vector<recs_segment_slice> mine;
        for (const recs_segment_slice& sl : slices)
            if (sl.segment == seg && sl.len) mine.push_back(sl);
        if (mine.empty()) continue;
        tasks->push_back([&s, &run, mine, self] {
            CurrentSystem cur(self);
            for (const recs_segment_slice& sl : mine)
                s.segment_fn(s.user, run.archs[sl.column], sl.begin, sl.len, run.cols[sl.column].data(),
                             (uint32_t)run.cols[sl.column].size());
        });
    }
}

// `SequentiallyIterable::run` for a CPU system on the calling thread.
int32_t run_cpu_system(recs_app* app, AppSystem& s) {
    CpuRun run;
    int32_t st = prepare_cpu_system(app, s, &run);
    if (st != RECS_OK) return st;
    std::vector<std::function<void()>> tasks;
    cpu_system_tasks(app, run, &tasks);
    for (auto& t : tasks) t();
    finish_cpu_system(app, run);
    return RECS_OK;
}

// A batch of `swap_remove(row)` calls (crates/ecs/src/archetypes.rs:233-243) folded into row moves: `origin` maps a
// current row to the row it held before the batch; rows that were never a target are absent.
struct SwapRemoveFold {
    std::unordered_map<uint32_t, uint32_t> origin;
    void remove(uint32_t row, uint32_t last) {  // `last` = index of the final row at the time of this removal
        auto lit = origin.find(last);
        const uint32_t from = lit == origin.end() ? last : lit->second;
        if (lit != origin.end()) origin.erase(lit);
        if (row != last) origin[row] = from;
    }
    void moves(std::vector<uint32_t>* dst, std::vector<uint32_t>* src) const {
        for (auto& m : origin)
            if (m.first != m.second) {
                dst->push_back(m.first);
                src->push_back(m.second);
            }
    }
};

// CommandPlayer::playback_commands (crates/ecs/src/systems/command_buffers.rs:377-435): creates, then
// removes (a set: duplicates collapse; entities that no longer exist are skipped like the reference's
// logged-and-ignored removal errors).  Removal order here is record order (GPU systems: system order,
// archetype order, ascending component index); the reference iterates a hash set, so the final row
// ORDER inside an archetype after several removals in one playback is not pinned by it.
// RECS_APP_PROFILE=1: where a tick's host time goes (printed by recs_app_destroy)
struct PhaseClock {
    double* slot;
    std::chrono::steady_clock::time_point t0;
    explicit PhaseClock(double* s) : slot(s), t0(std::chrono::steady_clock::now()) {}
    ~PhaseClock() { *slot += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); }
};

int32_t playback_commands(recs_app* app) {
    app->last_created = app->last_removed = 0;
    std::vector<recs_entity> to_remove = app->removes;
    {
    PhaseClock pc(&app->prof[1]);
    for (auto& src : app->gpu_removal_sources) {
        AppSystem& s = app->systems[src.first];
        for (uint32_t arch : src.second) {
            uint32_t n = 0;
            int32_t st = recs_gpu_system_removals(app->gpu, s.gpu_id, arch, nullptr, 0, &n);
            if (st != RECS_OK) return gpu_fail(app, st);
            if (!n) continue;
            std::vector<uint32_t> rows(n);
            if ((st = recs_gpu_system_removals(app->gpu, s.gpu_id, arch, rows.data(), n, &n)) != RECS_OK) return gpu_fail(app, st);
            const Archetype& a = app->world->archetypes[arch];
            for (uint32_t r : rows)
                if (r < a.entities.size()) to_remove.push_back(a.entities[r]);
        }
    }
    }
    app->gpu_removal_sources.clear();
    app->removes.clear();
    PhaseClock* pc2 = new PhaseClock(&app->prof[2]);
    // `Commands::create` calls of generic GPU systems: records fetched in sequential-run order, split into components
    for (uint32_t index : app->gpu_create_sources) {
        AppSystem& s = app->systems[index];
        uint32_t n = 0, bytes = 0;
        int32_t st = recs_gpu_system_creates(app->gpu, s.gpu_id, nullptr, 0, &n, &bytes);
        if (st != RECS_OK) return gpu_fail(app, st);
        if (!n) continue;
        std::vector<uint8_t> records((size_t)n * bytes);
        if ((st = recs_gpu_system_creates(app->gpu, s.gpu_id, records.data(), n, &n, &bytes)) != RECS_OK) return gpu_fail(app, st);
        // one batch per system and run: the records (array of structs) become one column per component
        CreateCommand cmd;
        cmd.system = index;
        cmd.count = n;
        size_t field = 0;
        for (size_t k = 0; k < s.create_comps.size(); ++k) {
            const size_t elem = s.create_sizes[k];
            cmd.ids.push_back(s.create_comps[k]);
            std::vector<uint8_t> column((size_t)n * elem);
            for (uint32_t r = 0; r < n; ++r) memcpy(column.data() + (size_t)r * elem, records.data() + (size_t)r * bytes + field, elem);
            cmd.values.push_back(std::move(column));
            field += elem;
        }
        app->creates.push_back(std::move(cmd));
    }
    app->gpu_create_sources.clear();
    delete pc2;
    // command buffers are drained system by system, in registration order (crates/ecs/src/lib.rs:325-336)
    std::stable_sort(app->creates.begin(), app->creates.end(),
                     [](const CreateCommand& a, const CreateCommand& b) { return a.system < b.system; });
    if (app->creates.empty() && to_remove.empty()) return RECS_OK;
    // The host World stays the authority on entity ids and row indices; resident device columns follow it WITHOUT a
    // round trip: appended rows are uploaded alone, swap-removes are replayed on the device as row moves.  A column
    // whose newest copy is on the device therefore stays there (its host bytes are stale but structurally in step).
    recs_world* w = app->world;
    struct Touched {
        uint64_t version_before = 0;
        uint64_t len_before = 0;                       // rows before this playback
        SwapRemoveFold fold;                            // rows as they were after the appends
        bool removed = false;
    };
    std::map<uint32_t, Touched> touched;
    auto touch = [&](uint32_t arch) -> Touched& {
        auto it = touched.find(arch);
        if (it == touched.end()) {
            Touched t;
            t.version_before = w->archetypes[arch].version;
            t.len_before = w->archetypes[arch].entities.size();
            it = touched.insert(std::make_pair(arch, t)).first;
        }
        return it->second;
    };
    int32_t st;
    const uint32_t archetypes_before = (uint32_t)w->archetypes.size();
    for (uint32_t a = 0; a < archetypes_before; ++a) touch(a);  // cheap: a handful of archetypes
    PhaseClock* pc3 = new PhaseClock(&app->prof[3]);
    for (CreateCommand& cmd : app->creates) {
        std::vector<const void*> ptrs(cmd.ids.size());
        for (size_t i = 0; i < cmd.ids.size(); ++i) ptrs[i] = cmd.values[i].empty() ? nullptr : cmd.values[i].data();
        if (cmd.count) {
            st = recs_world_create_entities(w, cmd.count, cmd.ids.data(), ptrs.data(), (uint32_t)cmd.ids.size(), nullptr);
            app->last_created += st == RECS_OK ? cmd.count : 0;
        } else {
            st = recs_world_create_entity(w, cmd.ids.data(), ptrs.data(), (uint32_t)cmd.ids.size(), nullptr);
            app->last_created += st == RECS_OK ? 1 : 0;
        }
        if (st != RECS_OK) return app->fail(st, recs_world_last_error(w));
    }
    app->creates.clear();
    delete pc3;
    // resident columns of grown archetypes: upload the new rows only
    std::set<ColKey> resident;  // columns kept in step on the device during this playback
    PhaseClock* pc4 = new PhaseClock(&app->prof[4]);
    if (app->gpu)
        for (auto& kv : touched) {
            const uint32_t arch = kv.first;
            Archetype& a = w->archetypes[arch];
            for (auto& col : a.columns) {
                auto mit = app->mirror.find(ColKey(arch, col.first));
                if (mit == app->mirror.end() || mit->second.bound_version != kv.second.version_before || mit->second.host_dirty)
                    continue;  // never bound, structurally out of date, or the host copy is the newer one: regular re-bind
                resident.insert(ColKey(arch, col.first));
                // zero-sized components (marker types) are bound with no rows: nothing to append, nothing to move
                if (a.entities.size() > kv.second.len_before && col.second.elem != 0) {
                    st = recs_gpu_column_append(app->gpu, arch, col.first, col.second.data.data(), a.entities.size());
                    if (st != RECS_OK) return gpu_fail(app, st);
                }
            }
        }
    delete pc4;
    PhaseClock* pc5 = new PhaseClock(&app->prof[5]);
    std::unordered_set<uint64_t> seen;
    seen.reserve(to_remove.size() * 2);
    for (recs_entity e : to_remove) {
        if (!seen.insert(entity_key(e)).second) continue;
        uint32_t arch = 0;
        uint64_t row = 0;
        if (recs_world_entity_archetype(w, e, &arch) != RECS_OK || recs_world_entity_component_index(w, e, &row) != RECS_OK)
            continue;  // already gone: the reference logs and ignores removal errors
        const uint32_t last = (uint32_t)w->archetypes[arch].entities.size() - 1;
        if (recs_world_delete_entity(w, e) != RECS_OK) continue;
        app->last_removed++;
        if (arch >= archetypes_before) continue;  // archetype created in this playback: nothing resident yet
        Touched& t = touch(arch);
        t.removed = true;
        t.fold.remove((uint32_t)row, last);
    }
    delete pc5;
    PhaseClock pc6(&app->prof[6]);
    if (app->gpu)
        for (auto& kv : touched) {
            const uint32_t arch = kv.first;
            Archetype& a = w->archetypes[arch];
            if (a.version == kv.second.version_before) continue;
            std::vector<uint64_t> comps;
            for (auto& col : a.columns)
                if (resident.count(ColKey(arch, col.first)) && col.second.elem != 0) comps.push_back(col.first);
            if (kv.second.removed && !comps.empty()) {
                std::vector<uint32_t> dst, src;
                kv.second.fold.moves(&dst, &src);
                st = recs_gpu_archetype_move_rows(app->gpu, arch, comps.data(), (uint32_t)comps.size(), dst.data(), src.data(),
                                                  (uint32_t)dst.size(), a.entities.size());
                if (st != RECS_OK) return gpu_fail(app, st);
            }
            for (auto& col : a.columns) {
                if (!resident.count(ColKey(arch, col.first))) continue;
                if ((st = recs_gpu_column_rebind_host(app->gpu, arch, col.first, col.second.data.data())) != RECS_OK)
                    return gpu_fail(app, st);
                app->mirror[ColKey(arch, col.first)].bound_version = a.version;
            }
        }
    return RECS_OK;
}

int32_t ensure_schedule(recs_app* app, int32_t ordered) {
    if (app->schedule && app->schedule_ordered == ordered && app->scheduled_systems == app->systems.size()) return RECS_OK;
    if (app->schedule) recs_schedule_destroy(app->schedule);
    app->schedule = nullptr;
    std::vector<recs_system_info> infos(app->systems.size());
    for (size_t i = 0; i < app->systems.size(); ++i) {
        infos[i].name = app->systems[i].name.c_str();
        infos[i].n_access = (uint32_t)app->systems[i].access.size();
        infos[i].access = app->systems[i].access.data();
    }
    int32_t st = recs_schedule_generate(infos.data(), (uint32_t)infos.size(), ordered, &app->schedule);
    if (st != RECS_OK) return app->fail(st, "schedule generation failed");
    recs_schedule_set_select_policy(app->schedule, 1);  // deterministic: oldest pending completion first
    app->schedule_ordered = ordered;
    app->scheduled_systems = app->systems.size();
    return RECS_OK;
}

}  // namespace

extern "C" {

int32_t recs_app_create(recs_gpu_ctx* gpu, recs_app** out) {
    if (!out) return RECS_ERR_INVALID_ARGUMENT;
    recs_app* app = new recs_app();
    app->gpu = gpu;
    recs_world_create(&app->world);
    *out = app;
    return RECS_OK;
}

int32_t recs_app_destroy(recs_app* app) {
    if (!app) return RECS_ERR_INVALID_ARGUMENT;
    if (getenv("RECS_APP_PROFILE") && app->prof[7] > 0)
        fprintf(stderr, "recs_app profile over %.0f ticks (ms per tick): systems+flush %.4f | fetch removals %.4f | fetch creates %.4f | "
                "world creates %.4f | device appends %.4f | world removes %.4f | device moves+rebind %.4f\n", app->prof[7],
                app->prof[0] / app->prof[7], app->prof[1] / app->prof[7], app->prof[2] / app->prof[7], app->prof[3] / app->prof[7],
                app->prof[4] / app->prof[7], app->prof[5] / app->prof[7], app->prof[6] / app->prof[7]);
    if (app->schedule) recs_schedule_destroy(app->schedule);
    recs_world_destroy(app->world);
    delete app;
    return RECS_OK;
}

const char* recs_app_last_error(recs_app* app) { return app ? app->error.c_str() : "null app"; }

recs_world* recs_app_world(recs_app* app) { return app ? app->world : nullptr; }

int32_t recs_app_add_gpu_system(recs_app* app, recs_gpu_system_kind kind, const uint64_t* ids, uint32_t n_ids,
                                uint32_t* out_index) {
    if (!app) return RECS_ERR_INVALID_ARGUMENT;
    if (!app->gpu) return app->fail(RECS_ERR_NO_DEVICE, "no GPU context: GPU systems cannot be registered (there is no CPU fallback)");
    AppSystem s;
    s.gpu = true;
    s.kind = kind;
    int32_t st = recs_gpu_system_create(app->gpu, kind, ids, n_ids, &s.gpu_id);
    if (st != RECS_OK) return gpu_fail(app, st);
    recs_gpu_system_desc desc;
    if ((st = recs_gpu_system_describe(app->gpu, s.gpu_id, &desc)) != RECS_OK) return gpu_fail(app, st);
    s.name = desc.name;
    s.access.assign(desc.access, desc.access + desc.n_access);
    const std::vector<recs_param_token> tokens = gpu_kind_tokens(kind, ids);
    std::string err;
    if (!parse_params(tokens.data(), (uint32_t)tokens.size(), &s.params, &err)) return app->fail(RECS_ERR_INVALID_ARGUMENT, err);
    app->systems.push_back(s);
    if (out_index) *out_index = (uint32_t)app->systems.size() - 1;
    return RECS_OK;
}

int32_t recs_app_add_gpu_kernel_system(recs_app* app, const char* name, const char* source, const char* body,
                                       const recs_param_token* tokens, uint32_t n_tokens, const char* const* type_names,
                                       const char* uniform_type, uint32_t uniform_bytes, uint32_t* out_index) {
    return recs_app_add_gpu_kernel_system_creating(app, name, source, body, tokens, n_tokens, type_names, uniform_type,
                                                   uniform_bytes, nullptr, 0, nullptr, out_index);
}

int32_t recs_app_add_gpu_kernel_system_creating(recs_app* app, const char* name, const char* source, const char* body,
                                                const recs_param_token* tokens, uint32_t n_tokens,
                                                const char* const* type_names, const char* uniform_type,
                                                uint32_t uniform_bytes, const uint64_t* create_components,
                                                uint32_t n_create_components, const char* create_record_type,
                                                uint32_t* out_index) {
    if (!app || !name || !source || !body || (n_tokens && !tokens) || !type_names) return RECS_ERR_INVALID_ARGUMENT;
    if (n_create_components && (!create_components || !create_record_type)) return RECS_ERR_INVALID_ARGUMENT;
    if (!app->gpu) return app->fail(RECS_ERR_NO_DEVICE, "no GPU context: GPU systems cannot be registered (there is no CPU fallback)");
    AppSystem s;
    s.gpu = true;
    s.kind = RECS_SYS_KERNEL;
    s.name = name;
    std::string err;
    if (!parse_params(tokens, n_tokens, &s.params, &err)) return app->fail(RECS_ERR_INVALID_ARGUMENT, err);
    // Read / Write / Entity become kernel arguments in order; filters only shape the archetype match
    std::vector<recs_gpu_kernel_param> kp;
    size_t n_typed = 0;  // Read / Write / Entity parameters seen so far (index into type_names)
    for (const Param& p : s.params) {
        recs_gpu_kernel_param k;
        k.component_id = p.comp;
        k.type_name = nullptr;
        if (p.op == RECS_PARAM_READ || p.op == RECS_PARAM_WRITE || p.op == RECS_PARAM_ENTITY) k.type_name = type_names[n_typed];
        if (p.op == RECS_PARAM_READ || p.op == RECS_PARAM_WRITE) {
            auto it = app->world->registry.find(p.comp);
            if (it == app->world->registry.end()) return app->fail(RECS_ERR_INVALID_ARGUMENT, "kernel system: component is not registered");
            k.elem_size = it->second.elem;
            k.kind = p.op == RECS_PARAM_WRITE ? RECS_KPARAM_WRITE : RECS_KPARAM_READ;
        } else if (p.op == RECS_PARAM_ENTITY) {
            k.elem_size = 8;
            k.kind = RECS_KPARAM_ENTITY;
        } else if (p.op == RECS_PARAM_COMMANDS) {  // no type name consumed from type_names
            k.component_id = 0;
            k.elem_size = 0;
            k.kind = RECS_KPARAM_COMMANDS;
            k.type_name = nullptr;
            // Commands::create: the record is the tuple of components of the new entity, packed in the given order
            for (uint32_t ci = 0; ci < n_create_components; ++ci) {
                auto it = app->world->registry.find(create_components[ci]);
                if (it == app->world->registry.end())
                    return app->fail(RECS_ERR_INVALID_ARGUMENT, "kernel system: a component of the create record is not registered");
                s.create_comps.push_back(create_components[ci]);
                s.create_sizes.push_back(it->second.elem);
                k.elem_size += it->second.elem;
            }
            if (n_create_components) k.type_name = create_record_type;
            kp.push_back(k);
            s.decides_removals = true;
            continue;
        } else if (p.op == RECS_PARAM_QUERY_BEGIN) {
            return app->fail(RECS_ERR_UNSUPPORTED, "kernel system: nested Query<...> parameters are not available in generic GPU systems");
        } else {
            continue;  // With / Not / And / Or / Xor
        }
        if (!k.type_name) return app->fail(RECS_ERR_INVALID_ARGUMENT, "kernel system: missing type name");
        kp.push_back(k);
        ++n_typed;
    }
    if (n_create_components && s.create_comps.empty())
        return app->fail(RECS_ERR_INVALID_ARGUMENT, "kernel system: a create record needs a Commands parameter");
    int32_t st = recs_gpu_system_create_kernel(app->gpu, name, source, body, kp.data(), (uint32_t)kp.size(), uniform_type,
                                               uniform_bytes, &s.gpu_id);
    if (st != RECS_OK) return gpu_fail(app, st);
    params_accesses(s.params, &s.access);
    app->systems.push_back(s);
    if (out_index) *out_index = (uint32_t)app->systems.size() - 1;
    return RECS_OK;
}

int32_t recs_fold_swap_removes(uint64_t len_before, const uint32_t* rows, uint32_t n_rows, uint32_t* out_dst,
                               uint32_t* out_src, uint32_t cap, uint32_t* out_n_moves, uint64_t* out_len_after) {
    if ((n_rows && !rows) || !out_n_moves) return RECS_ERR_INVALID_ARGUMENT;
    SwapRemoveFold fold;
    uint64_t len = len_before;
    for (uint32_t i = 0; i < n_rows; ++i) {
        if (len == 0 || rows[i] >= len) return RECS_ERR_INVALID_ARGUMENT;
        fold.remove(rows[i], (uint32_t)(len - 1));
        --len;
    }
    std::vector<uint32_t> dst, src;
    fold.moves(&dst, &src);
    *out_n_moves = (uint32_t)dst.size();
    if (out_len_after) *out_len_after = len;
    for (uint32_t i = 0; i < dst.size() && i < cap; ++i) {
        if (out_dst) out_dst[i] = dst[i];
        if (out_src) out_src[i] = src[i];
    }
    return RECS_OK;
}

int32_t recs_app_add_gpu_pair_kernel_system(recs_app* app, const char* name, const char* source, const char* acc_type,
                                            const char* pair_fn, const char* finish_fn, const recs_param_token* tokens,
                                            uint32_t n_tokens, const char* const* type_names, const char* uniform_type,
                                            uint32_t uniform_bytes, uint32_t* out_index) {
    if (!app || !name || !source || !acc_type || !pair_fn || !finish_fn || (n_tokens && !tokens) || !type_names)
        return RECS_ERR_INVALID_ARGUMENT;
    if (!app->gpu) return app->fail(RECS_ERR_NO_DEVICE, "no GPU context: GPU systems cannot be registered (there is no CPU fallback)");
    AppSystem s;
    s.gpu = true;
    s.kind = RECS_SYS_KERNEL;
    s.name = name;
    std::string err;
    if (!parse_params(tokens, n_tokens, &s.params, &err)) return app->fail(RECS_ERR_INVALID_ARGUMENT, err);
    std::vector<recs_gpu_kernel_param> kp;
    size_t n_typed = 0;
    bool have_query = false;
    auto typed = [&](const Param& p, bool nested) -> int32_t {
        recs_gpu_kernel_param k;
        k.component_id = p.comp;
        k.type_name = type_names[n_typed];
        if (!k.type_name) return app->fail(RECS_ERR_INVALID_ARGUMENT, "kernel system: missing type name");
        if (p.op == RECS_PARAM_ENTITY) {
            k.elem_size = 8;
            k.kind = nested ? RECS_KPARAM_QUERY_ENTITY : RECS_KPARAM_ENTITY;
        } else {
            auto it = app->world->registry.find(p.comp);
            if (it == app->world->registry.end()) return app->fail(RECS_ERR_INVALID_ARGUMENT, "kernel system: component is not registered");
            k.elem_size = it->second.elem;
            if (nested && p.op == RECS_PARAM_WRITE)
                return app->fail(RECS_ERR_UNSUPPORTED, "kernel system: a nested Query may only read (systems/mod.rs:1101-1105)");
            k.kind = nested ? RECS_KPARAM_QUERY_READ : (p.op == RECS_PARAM_WRITE ? RECS_KPARAM_WRITE : RECS_KPARAM_READ);
        }
        kp.push_back(k);
        ++n_typed;
        return RECS_OK;
    };
    for (const Param& p : s.params) {
        int32_t st = RECS_OK;
        if (p.op == RECS_PARAM_READ || p.op == RECS_PARAM_WRITE || p.op == RECS_PARAM_ENTITY) {
            st = typed(p, false);
        } else if (p.op == RECS_PARAM_COMMANDS) {
            recs_gpu_kernel_param k;
            k.component_id = 0;
            k.elem_size = 0;
            k.kind = RECS_KPARAM_COMMANDS;
            k.type_name = nullptr;
            kp.push_back(k);
            s.decides_removals = true;
        } else if (p.op == RECS_PARAM_QUERY_BEGIN) {
            if (have_query) return app->fail(RECS_ERR_UNSUPPORTED, "kernel system: one nested Query per system");
            have_query = true;
            for (const Param& q : p.children)
                if (q.op == RECS_PARAM_READ || q.op == RECS_PARAM_WRITE || q.op == RECS_PARAM_ENTITY)
                    if ((st = typed(q, true)) != RECS_OK) break;
        }
        if (st != RECS_OK) return st;
    }
    int32_t st = recs_gpu_system_create_pair_kernel(app->gpu, name, source, acc_type, pair_fn, finish_fn, kp.data(),
                                                    (uint32_t)kp.size(), uniform_type, uniform_bytes, &s.gpu_id);
    if (st != RECS_OK) return gpu_fail(app, st);
    params_accesses(s.params, &s.access);
    app->systems.push_back(s);
    if (out_index) *out_index = (uint32_t)app->systems.size() - 1;
    return RECS_OK;
}

int32_t recs_app_set_system_uniform(recs_app* app, uint32_t index, const void* data, uint32_t bytes) {
    if (!app || index >= app->systems.size() || !app->systems[index].gpu) return RECS_ERR_INVALID_ARGUMENT;
    int32_t st = recs_gpu_system_set_uniform(app->gpu, app->systems[index].gpu_id, data, bytes);
    return st == RECS_OK ? RECS_OK : gpu_fail(app, st);
}

int32_t recs_app_add_cpu_system(recs_app* app, const char* name, const recs_param_token* tokens, uint32_t n_tokens,
                                recs_cpu_system_fn fn, void* user, uint32_t* out_index) {
    if (!app || !fn) return RECS_ERR_INVALID_ARGUMENT;
    AppSystem s;
    s.gpu = false;
    s.name = name ? name : "system";
    s.fn = fn;
    s.user = user;
    std::string err;
    if (!parse_params(tokens, n_tokens, &s.params, &err)) return app->fail(RECS_ERR_INVALID_ARGUMENT, err);
    params_accesses(s.params, &s.access);
    app->systems.push_back(s);
    if (out_index) *out_index = (uint32_t)app->systems.size() - 1;
    return RECS_OK;
}

int32_t recs_app_add_cpu_segment_system(recs_app* app, const char* name, const recs_param_token* tokens, uint32_t n_tokens,
                                        recs_cpu_segment_fn fn, void* user, uint32_t* out_index) {
    if (!app || !fn) return RECS_ERR_INVALID_ARGUMENT;
    AppSystem s;
    s.gpu = false;
    s.name = name ? name : "system";
    s.segment_fn = fn;
    s.user = user;
    std::string err;
    if (!parse_params(tokens, n_tokens, &s.params, &err)) return app->fail(RECS_ERR_INVALID_ARGUMENT, err);
    if (!params_supports_parallelization(s.params))
        return app->fail(RECS_ERR_UNSUPPORTED, "segment systems need parameters that all support parallelization (a nested Query only with read accesses, systems/mod.rs:1101-1105)");
    params_accesses(s.params, &s.access);
    app->systems.push_back(s);
    if (out_index) *out_index = (uint32_t)app->systems.size() - 1;
    return RECS_OK;
}

int32_t recs_app_set_executor(recs_app* app, uint32_t worker_threads) {
    if (!app) return RECS_ERR_INVALID_ARGUMENT;
    if (worker_threads == 0)
        app->pool.reset();
    else if (!app->pool || app->pool->size() != worker_threads)
        app->pool.reset(new WorkerPool(worker_threads));
    return RECS_OK;
}

namespace {
bool whole_tick_is_one_gpu_chain(recs_app* app) {
    if (!app->gpu || app->systems.empty() || app->chains_flushed != 1 || app->last_chain_len != app->systems.size()) return false;
    for (const AppSystem& s : app->systems) {
        if (!s.gpu || s.decides_removals || !s.create_comps.empty()) return false;
        if (s.kind == RECS_SYS_NBODY_COLLISION || s.kind == RECS_SYS_RAIN_GROUND_COLLISION) return false;
        if (s.kind == RECS_SYS_RAIN_WIND_DIRECTION) return false;  // a host-side per-tick value
    }
    std::lock_guard<std::mutex> lock(app->command_mu);
    return app->creates.empty() && app->removes.empty();
}
}  // namespace

int32_t recs_app_run(recs_app* app, uint32_t n_ticks, int32_t ordered) {
    if (!app) return RECS_ERR_INVALID_ARGUMENT;
    int32_t st = ensure_schedule(app, ordered ? 1 : 0);
    if (st != RECS_OK) return st;
    std::vector<uint32_t> batch(app->systems.size() + 1);
    for (uint32_t tick = 0; tick < n_ticks; ++tick) {
        app->gpu_chain.clear();  // nothing survives a tick (or an error return)
        app->chains_flushed = 0;
        app->gpu_create_sources.clear();
        app->last_order.clear();
        app->last_batch.clear();
        uint32_t batch_no = 0;
        app->prof[7] += 1;
        PhaseClock* pc0 = new PhaseClock(&app->prof[0]);
        for (;;) {  // Executor::execute_once
            uint32_t got = 0;
            st = recs_schedule_next_batch(app->schedule, RECS_REACTION_RETURN_ERROR, batch.data(), (uint32_t)batch.size(), &got);
            if (st == RECS_SCHEDULE_NEW_TICK) break;
            if (st != RECS_OK) return app->fail(st, "schedule error");
            if (!app->pool) {
                for (uint32_t i = 0; i < got; ++i) {
                    AppSystem& s = app->systems[batch[i]];
                    st = s.gpu ? run_gpu_system(app, s) : run_cpu_system(app, s);
                    if (st != RECS_OK) return st;
                    app->last_order.push_back(batch[i]);
                    app->last_batch.push_back(batch_no);
                }
            } else {
                // WorkerPool::execute_once (executor/mod.rs:191-236): the systems of a batch do not conflict, so
                // their tasks share the injector queue; GPU systems are single tasks issued from this thread.
                std::vector<CpuRun> runs(got);
                size_t n_runs = 0;
                for (uint32_t i = 0; i < got && st == RECS_OK; ++i) {
                    AppSystem& s = app->systems[batch[i]];
                    if (!s.gpu) st = prepare_cpu_system(app, s, &runs[n_runs++]);
                }
                if (st == RECS_OK) {
                    std::vector<std::function<void()>> tasks;
                    for (size_t k = 0; k < n_runs; ++k) cpu_system_tasks(app, runs[k], &tasks);
                    for (auto& t : tasks) app->pool->submit(std::move(t));
                    for (uint32_t i = 0; i < got && st == RECS_OK; ++i) {
                        AppSystem& s = app->systems[batch[i]];
                        if (s.gpu) st = run_gpu_system(app, s);
                    }
                    app->pool->wait_idle();
                }
                for (size_t k = 0; k < n_runs; ++k)
                    if (runs[k].sys) finish_cpu_system(app, runs[k]);
                if (st != RECS_OK) return st;
                for (uint32_t i = 0; i < got; ++i) {
                    app->last_order.push_back(batch[i]);
                    app->last_batch.push_back(batch_no);
                }
            }
            // completion == dropping the SystemExecutionGuard (lib.rs:525-557)
            for (uint32_t i = 0; i < got; ++i) recs_schedule_complete(app->schedule, batch[i]);
            ++batch_no;
        }
        st = flush_gpu_chain(app);
        delete pc0;
        if (st != RECS_OK) return st;
        // loop { runner.tick(); runner.playback_commands(); } (lib.rs:321-323)
        if ((st = playback_commands(app)) != RECS_OK) return st;
        // A tick that was ONE chain of GPU systems holding every system of the application, none of which can record a
        // command: nothing on the host changes from tick to tick (same batches, same archetypes, no playback), so the
        // remaining ticks are the same chain again -- handed to the context in one call, which captures it in a CUDA graph
        // or, for a small n-body world, runs all of them in one cooperative launch (recs_gpu_tick).
        if (tick + 1 < n_ticks && whole_tick_is_one_gpu_chain(app)) {
            st = recs_gpu_tick(app->gpu, n_ticks - tick - 1);  // (the tick order is the one flush_gpu_chain just set)
            return st == RECS_OK ? RECS_OK : gpu_fail(app, st);
        }
    }
    return RECS_OK;
}

int32_t recs_app_command_create(recs_app* app, const uint64_t* ids, const void* const* values, uint32_t n) {
    if (!app || (n && !ids)) return RECS_ERR_INVALID_ARGUMENT;
    CreateCommand cmd;
    for (uint32_t i = 0; i < n; ++i) {
        auto it = app->world->registry.find(ids[i]);
        if (it == app->world->registry.end()) return app->fail(RECS_ERR_INVALID_ARGUMENT, "command_create: component is not registered");
        cmd.ids.push_back(ids[i]);
        std::vector<uint8_t> bytes(it->second.elem);
        if (it->second.elem && values && values[i]) memcpy(bytes.data(), values[i], it->second.elem);
        cmd.values.push_back(bytes);
    }
    cmd.system = t_current_system;
    std::lock_guard<std::mutex> lock(app->command_mu);
    app->creates.push_back(cmd);
    return RECS_OK;
}

int32_t recs_app_command_remove(recs_app* app, recs_entity entity) {
    if (!app) return RECS_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lock(app->command_mu);
    app->removes.push_back(entity);
    return RECS_OK;
}

int32_t recs_app_last_playback(recs_app* app, uint64_t* out_created, uint64_t* out_removed) {
    if (!app) return RECS_ERR_INVALID_ARGUMENT;
    if (out_created) *out_created = app->last_created;
    if (out_removed) *out_removed = app->last_removed;
    return RECS_OK;
}

int32_t recs_app_last_tick_order(recs_app* app, uint32_t* out_order, uint32_t* out_batch_of, uint32_t cap, uint32_t* out_n) {
    if (!app) return RECS_ERR_INVALID_ARGUMENT;
    if (out_n) *out_n = (uint32_t)app->last_order.size();
    for (uint32_t i = 0; i < app->last_order.size() && i < cap; ++i) {
        if (out_order) out_order[i] = app->last_order[i];
        if (out_batch_of) out_batch_of[i] = app->last_batch[i];
    }
    return RECS_OK;
}

int32_t recs_app_sync_to_host(recs_app* app) {
    if (!app) return RECS_ERR_INVALID_ARGUMENT;
    if (app->gpu) {
        int32_t flushed = flush_gpu_chain(app);
        if (flushed != RECS_OK) return flushed;
    }
    for (auto& kv : app->mirror) {
        if (!kv.second.device_dirty) continue;
        int32_t st = recs_gpu_download(app->gpu, kv.first.first, kv.first.second);
        if (st != RECS_OK) return gpu_fail(app, st);
        kv.second.device_dirty = false;
    }
    if (app->gpu) {
        int32_t st = recs_gpu_synchronize(app->gpu);
        if (st != RECS_OK) return gpu_fail(app, st);
    }
    return RECS_OK;
}

// Application::add_component / remove_component (crates/ecs/src/lib.rs:213-228): the entity moves to another archetype
// (crates/ecs/src/archetypes.rs:474-518, 556-651: swap_remove from the source, push onto the target).  Resident device
// columns of both archetypes follow WITHOUT a re-upload: the moved row is fetched from the device where the device holds
// the newest copy (one row per column), the source archetype replays its swap_remove as a row move, the target's
// resident columns receive the pushed row as an append.
namespace {
int32_t move_entity_between_archetypes(recs_app* app, recs_entity e, const std::function<int32_t()>& host_op) {
    recs_world* w = app->world;
    if (!app->gpu) {
        const int32_t st = host_op();
        return st == RECS_OK ? RECS_OK : app->fail(st, recs_world_last_error(w));
    }
    int32_t st = flush_gpu_chain(app);
    if (st != RECS_OK) return st;
    uint32_t src = 0;
    uint64_t row = 0;
    if (recs_world_entity_archetype(w, e, &src) != RECS_OK || recs_world_entity_component_index(w, e, &row) != RECS_OK)
        return app->fail(RECS_ERR_WORLD, recs_world_last_error(w));
    std::vector<uint64_t> version_before(w->archetypes.size());
    for (size_t a = 0; a < w->archetypes.size(); ++a) version_before[a] = w->archetypes[a].version;
    auto is_resident = [&](uint32_t arch, uint64_t comp, uint32_t elem) {
        auto mit = app->mirror.find(ColKey(arch, comp));
        return elem != 0 && mit != app->mirror.end() && arch < version_before.size() &&
               mit->second.bound_version == version_before[arch] && !mit->second.host_dirty;
    };
    // the row about to move: make the host copy current where the device copy is the newer one
    const uint64_t last = w->archetypes[src].entities.size() - 1;
    std::vector<uint64_t> src_resident;
    for (auto& col : w->archetypes[src].columns) {
        if (!is_resident(src, col.first, col.second.elem)) continue;
        src_resident.push_back(col.first);
        if (app->mirror[ColKey(src, col.first)].device_dirty)
            if ((st = recs_gpu_download_rows(app->gpu, src, col.first, row, row + 1)) != RECS_OK) return gpu_fail(app, st);
    }
    if ((st = host_op()) != RECS_OK) return app->fail(st, recs_world_last_error(w));
    uint32_t dst = 0;
    if (recs_world_entity_archetype(w, e, &dst) != RECS_OK) return app->fail(RECS_ERR_WORLD, recs_world_last_error(w));
    // source: swap_remove(row) replayed on the device
    Archetype& sa = w->archetypes[src];
    if (!src_resident.empty()) {
        const uint32_t d = (uint32_t)row, s_ = (uint32_t)last;
        st = recs_gpu_archetype_move_rows(app->gpu, src, src_resident.data(), (uint32_t)src_resident.size(), &d, &s_, row != last ? 1u : 0u,
                                          sa.entities.size());
        if (st != RECS_OK) return gpu_fail(app, st);
        for (uint64_t comp : src_resident) {
            if ((st = recs_gpu_column_rebind_host(app->gpu, src, comp, sa.columns[comp].data.data())) != RECS_OK) return gpu_fail(app, st);
            app->mirror[ColKey(src, comp)].bound_version = sa.version;
        }
    }
    // target: the pushed row goes up alone (an archetype created by this move has nothing resident yet)
    if (dst < version_before.size()) {
        Archetype& da = w->archetypes[dst];
        for (auto& col : da.columns) {
            if (!is_resident(dst, col.first, col.second.elem)) continue;
            if ((st = recs_gpu_column_append(app->gpu, dst, col.first, col.second.data.data(), da.entities.size())) != RECS_OK)
                return gpu_fail(app, st);
            app->mirror[ColKey(dst, col.first)].bound_version = da.version;
        }
    }
    return RECS_OK;
}
}  // namespace

int32_t recs_app_add_components(recs_app* app, recs_entity entity, const uint64_t* ids, const void* const* values, uint32_t n) {
    if (!app || (n && !ids)) return RECS_ERR_INVALID_ARGUMENT;
    return move_entity_between_archetypes(app, entity, [&] { return recs_world_add_components(app->world, entity, ids, values, n); });
}

int32_t recs_app_remove_components(recs_app* app, recs_entity entity, const uint64_t* ids, uint32_t n) {
    if (!app || (n && !ids)) return RECS_ERR_INVALID_ARGUMENT;
    return move_entity_between_archetypes(app, entity, [&] { return recs_world_remove_components(app->world, entity, ids, n); });
}

int32_t recs_app_mark_host_dirty(recs_app* app, uint32_t archetype) {
    if (!app) return RECS_ERR_INVALID_ARGUMENT;
    for (auto& kv : app->mirror)
        if (kv.first.first == archetype) kv.second.host_dirty = true;
    return RECS_OK;
}

}  // extern "C"
