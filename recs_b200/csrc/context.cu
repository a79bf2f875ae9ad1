// librecs_gpu.so -- context, device column mirrors, GPU systems and the C ABI (include/recs_gpu.h).
//
// Host-side logic only; the kernels are in gravity.cu / streaming.cu.  A context owns one CUDA
// stream (plus a communication stream when sharded), the device mirrors of the bound component
// columns and the workspace of the gravity system (pair-interleaved posm mirror, fragments).
#include <cuda_runtime.h>
#include <math.h>
#include <nvtx3/nvToolsExt.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "../../include/recs_gpu.h"
#include "host/kernel_gen.h"
#include "kernels.cuh"
#include "nccl_dyn.h"

namespace recs {
namespace {

struct Column {
    void* host = nullptr;
    void* dev = nullptr;
    uint32_t elem = 0;
    uint64_t count = 0;
    size_t cap = 0;  // bytes allocated at dev
    bool stale = false;
    uint64_t version = 0;  // bumped on every write (upload or system)
};

// State of a generic (NVRTC-compiled) system, recs_gpu_system_create_kernel.
struct KernelSys {
    std::vector<KernelParamSpec> params;
    std::vector<uint64_t> comp;            // component id per parameter
    KernelPlan plan;
    cudaLibrary_t lib = nullptr;
    cudaKernel_t kernel = nullptr;
    std::vector<unsigned char> uniform;    // bytes passed by value to every launch
    bool has_uniform = false;
    bool has_commands = false;             // one RECS_KPARAM_COMMANDS parameter: the system decides removals
    // Commands::create: device append buffer {count, cap, record_bytes, bad | records}, see kernel_gen.cpp
    uint32_t create_bytes = 0;             // sizeof of the record create() takes (0: the system only removes)
    uint32_t create_cap = 65536;           // records per launch (recs_gpu_system_set_create_capacity)
    unsigned char* create_buf = nullptr;
    size_t create_buf_cap = 0;
    bool create_valid = false;             // the buffer holds the records of the last run
    uint32_t create_stride() const { return 16u + ((create_bytes + 7u) & ~7u); }
    // nested Query<(...)> form (recs_gpu_system_create_pair_kernel)
    bool is_pair = false;
    std::vector<KernelParamSpec> qparams;
    std::vector<uint64_t> qcomp;
    unsigned long long* qtable = nullptr;
    size_t qtable_cap = 0;
    std::vector<unsigned long long> qtable_host;
    uint32_t n_qarch = 0;
    // archetype table on the device, rebuilt when bindings / membership change
    unsigned long long* table = nullptr;
    size_t table_cap = 0;
    std::vector<unsigned long long> table_host;
    uint64_t table_epoch = ~0ull;
    uint64_t n_items = 0;
    uint32_t n_arch = 0;
    uint64_t footprint_bytes = 0;          // bytes of all columns one launch touches
    std::vector<uint32_t> table_archs;     // archetypes with rows, in table order
    ~KernelSys() {
        if (table) cudaFree(table);
        if (qtable) cudaFree(qtable);
        if (create_buf) cudaFree(create_buf);
        if (lib) cudaLibraryUnload(lib);
    }
};

struct System {
    recs_gpu_system_kind kind;
    std::shared_ptr<KernelSys> ks;
    uint64_t comp[3] = {0, 0, 0};
    uint32_t n_comp = 0;
    std::string name;
    std::vector<recs_gpu_access> access;
    std::vector<uint32_t> archs;   // outer parameters' archetypes, iteration order
    std::vector<uint32_t> qarchs;  // nested Query archetypes (gravity, collision)
    // removal-deciding systems: per outer archetype, flags + compacted row list of the last run
    struct Removals {
        uint32_t* flags = nullptr;
        uint32_t* rows = nullptr;
        uint32_t* counter = nullptr;
        size_t cap_rows = 0;
        uint32_t n_rows = 0;
        bool valid = false;
    };
    std::map<uint32_t, Removals> removals;
};

enum TimeSlot { T_GRAVITY = 0, T_REDUCE, T_MOVEMENT, T_ACCEL, T_PACK, T_OTHER, T_CHAIN, T_GATHER, T_COUNT };

struct TimedSpan {
    cudaEvent_t a, b;
    int slot;
};

typedef std::pair<uint32_t, uint64_t> ColKey;

}  // namespace
}  // namespace recs

using namespace recs;

struct recs_gpu_ctx {
    std::mutex mu;
    recs_gpu_config cfg;
    int32_t status = RECS_OK;
    std::string error;
    cudaStream_t stream = nullptr;
    cudaStream_t comm_stream = nullptr;
    cudaEvent_t ev_pack = nullptr, ev_gather = nullptr;
    int sm_count = 148;
    int grav_ctas_per_sm[kGravVariantCount] = {0};
    bool grav_variant_forced = false;
    bool posm_scaled = false;
    int grav_variant = kGravVariantDefault;

    std::map<ColKey, Column> columns;
    std::vector<System> systems;
    std::vector<uint32_t> order;
    uint64_t epoch = 0;  // bumped whenever bindings / archetype lists / order change

    // gravity workspace
    float* posm = nullptr;
    size_t posm_cap = 0;
    // second mirror buffer + grid-barrier word of the many-ticks-in-one-launch path (small single-archetype worlds)
    float* posm_alt = nullptr;
    size_t posm_alt_cap = 0;
    uint32_t* resident_barrier = nullptr;
    uint32_t resident_barrier_count = 0;  // arrivals counted so far (the counter only grows: no reset between launches)
    uint64_t posm_pack_serial = 0;        // bumped by every pack of the mirror
    uint64_t posm_alt_serial = ~0ull;     // pack whose G*m entries and null bodies the second buffer holds
    uint64_t resident_max_rows = 4096;  // RECS_GPU_RESIDENT_MAX_ROWS (0 switches the path off)
    uint64_t resident_launches = 0;
    uint64_t posm_bodies = 0;  // padded body count currently laid out
    std::vector<std::pair<ColKey, uint64_t>> posm_sources;  // (column, version) the mirror was packed from
    float4* frags = nullptr;
    size_t frags_cap = 0;
    int grav_grid_forced = 0;           // RECS_GPU_GRAVITY_GRID: test knob (results must not depend on it)
    bool grav_no_heavy = false;         // RECS_GPU_GRAVITY_NO_LAST_SPLIT: A/B and test knob (ditto)
    int emulate_ranks = 0, emulate_rank = 0;  // RECS_GPU_GRAVITY_EMULATE_RANK="R,r": rank r's phase split on one GPU (test knob)

    // sharded runs, peer-pushed mirror: one allocation [mirror buffer 0 | mirror buffer 1 | flags | done counter],
    // mapped into every rank through CUDA IPC (recs_gpu_exchange_export / _import)
    struct Exchange {
        unsigned char* base = nullptr;
        uint64_t cap_bodies = 0;        // bodies per mirror buffer
        bool imported = false;
        unsigned char* peer[kGravMaxPeers] = {nullptr};  // every rank's base in this process (own rank: base)
        uint32_t gen = 0;               // newest generation pushed by this rank
        float* buffer(int rank_slot, uint32_t g) const { return reinterpret_cast<float*>(peer[rank_slot] + (size_t)(g & 1u) * cap_bodies * 16); }
        uint32_t* flags(int rank_slot) const { return reinterpret_cast<uint32_t*>(peer[rank_slot] + 2 * (size_t)cap_bodies * 16); }
        uint32_t* done_counter() const { return reinterpret_cast<uint32_t*>(base + 2 * (size_t)cap_bodies * 16 + 256); }
        size_t bytes() const { return 2 * (size_t)cap_bodies * 16 + 512; }
    } xchg;
    uint32_t* dev_error_host = nullptr;  // mapped pinned word the kernels raise when a peer wait timed out
    uint32_t* dev_error = nullptr;
    bool exchange_nccl_forced = false;   // RECS_GPU_EXCHANGE=nccl

    // rain
    float wind[3] = {1.f, 0.f, 0.f};

    // CUDA graph of one tick
    cudaGraphExec_t graph = nullptr;
    uint64_t graph_epoch = ~0ull;
    uint32_t graph_ticks = 1;        // ticks held by the captured graph
    uint32_t graph_ticks_wanted = 8; // RECS_GPU_GRAPH_TICKS: 1-2 us per tick less than one graph launch per tick at 5 000 .. 16 384 bodies
    bool capturing = false;
    bool graph_elided_pack = false;  // the captured tick found the mirror current and holds no pack kernel

    // timing
    bool timing = false;
    std::vector<TimedSpan> spans;
    std::vector<cudaEvent_t> event_pool;
    uint64_t kernel_launches = 0;
    uint64_t gravity_launches = 0;
    float chain_ms = 0.f;

    // NCCL
    NcclApi nccl;
    NcclApi::comm_t comm = nullptr;

    // misc
    cudaEvent_t timer_a = nullptr, timer_b = nullptr;
    float* l2_flush = nullptr;
    size_t l2_flush_bytes = 0;
    bool l2_flush_zeroed = false;
    // structural changes on resident columns
    uint32_t* move_idx = nullptr;
    size_t move_idx_cap = 0;
    unsigned char* move_tmp = nullptr;
    size_t move_tmp_cap = 0;
    uint64_t h2d_bytes = 0, d2h_bytes = 0;  // component-column traffic over PCIe since create / reset
    // small host -> device copies (archetype tables, row-move lists, appended rows) go through a pinned two-half
    // staging buffer, so the source may be pageable / short-lived and the stream is never synchronised for them
    unsigned char* stage = nullptr;
    size_t stage_half = 0, stage_used = 0;
    int stage_cur = 0;
    cudaEvent_t stage_ev[2] = {nullptr, nullptr};
    bool stage_busy[2] = {false, false};
};

namespace {

int32_t fail(recs_gpu_ctx* c, int32_t code, const std::string& msg) {
    if (c->status == RECS_OK) {
        c->status = code;
        c->error = msg;
    }
    return code;
}

#define RECS_CUDA(c, expr)                                                                       \
    do {                                                                                         \
        cudaError_t e__ = (expr);                                                                \
        if (e__ != cudaSuccess)                                                                  \
            return fail((c), RECS_ERR_CUDA,                                                      \
                        std::string(#expr) + ": " + cudaGetErrorString(e__));                   \
    } while (0)

#define RECS_ENTER(c)                                                                            \
    if (!(c)) return RECS_ERR_INVALID_ARGUMENT;                                                  \
    std::lock_guard<std::mutex> lock__((c)->mu);                                                 \
    if ((c)->status != RECS_OK) return (c)->status;                                              \
    {                                                                                            \
        cudaError_t e__ = cudaSetDevice((c)->cfg.device);                                        \
        if (e__ != cudaSuccess) return fail((c), RECS_ERR_CUDA, cudaGetErrorString(e__));        \
    }

cudaEvent_t get_event(recs_gpu_ctx* c) {
    if (!c->event_pool.empty()) {
        cudaEvent_t e = c->event_pool.back();
        c->event_pool.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

// RAII bracket of one kernel category with CUDA events on the context stream.
struct Timed {
    recs_gpu_ctx* c;
    TimedSpan s;
    bool on;
    Timed(recs_gpu_ctx* ctx, int slot) : c(ctx), on(ctx->timing && !ctx->capturing) {
        if (on) {
            s.a = get_event(c);
            s.b = get_event(c);
            s.slot = slot;
            cudaEventRecord(s.a, c->stream);
        }
    }
    ~Timed() {
        if (on) {
            cudaEventRecord(s.b, c->stream);
            c->spans.push_back(s);
        }
    }
};

Column* find_column(recs_gpu_ctx* c, uint32_t arch, uint64_t comp) {
    auto it = c->columns.find(ColKey(arch, comp));
    return it == c->columns.end() ? nullptr : &it->second;
}

int32_t need_column(recs_gpu_ctx* c, uint32_t arch, uint64_t comp, uint32_t elem, Column** out) {
    Column* col = find_column(c, arch, comp);
    char buf[160];
    if (!col || !col->dev) {
        snprintf(buf, sizeof buf, "could not execute system due to missing parameter: archetype %u component %llu is not bound",
                 arch, (unsigned long long)comp);
        return fail(c, RECS_ERR_MISSING_PARAMETER, buf);
    }
    if (col->stale) {
        snprintf(buf, sizeof buf, "archetype %u component %llu was invalidated and not rebound", arch,
                 (unsigned long long)comp);
        return fail(c, RECS_ERR_STALE_BINDING, buf);
    }
    if (col->elem != elem) {
        snprintf(buf, sizeof buf, "archetype %u component %llu has element size %u, system expects %u", arch,
                 (unsigned long long)comp, col->elem, elem);
        return fail(c, RECS_ERR_SIZE_MISMATCH, buf);
    }
    *out = col;
    return RECS_OK;
}

int32_t same_count(recs_gpu_ctx* c, const Column* a, const Column* b) {
    if (a->count != b->count) return fail(c, RECS_ERR_SIZE_MISMATCH, "columns of one archetype differ in length");
    return RECS_OK;
}

template <typename T>
int32_t ensure(recs_gpu_ctx* c, T** p, size_t* cap, size_t bytes) {
    if (*cap >= bytes && *p) return RECS_OK;
    if (c->capturing) return fail(c, RECS_ERR_CUDA, "workspace growth during graph capture");
    if (*p) RECS_CUDA(c, cudaFree(*p));
    *p = nullptr;
    *cap = 0;
    RECS_CUDA(c, cudaMalloc(reinterpret_cast<void**>(p), bytes));
    *cap = bytes;
    return RECS_OK;
}

// Asynchronous copy of a small host block whose lifetime ends with the call.
int32_t staged_h2d(recs_gpu_ctx* c, void* dst, const void* src, size_t bytes) {
    if (bytes == 0) return RECS_OK;
    const size_t kHalf = 1u << 20;
    if (!c->stage) {
        RECS_CUDA(c, cudaHostAlloc((void**)&c->stage, 2 * kHalf, cudaHostAllocDefault));
        RECS_CUDA(c, cudaEventCreateWithFlags(&c->stage_ev[0], cudaEventDisableTiming));
        RECS_CUDA(c, cudaEventCreateWithFlags(&c->stage_ev[1], cudaEventDisableTiming));
        c->stage_half = kHalf;
    }
    const size_t need = (bytes + 255) & ~(size_t)255;
    if (need > c->stage_half || c->capturing) {  // too large to stage: copy straight from the caller's memory and wait
        RECS_CUDA(c, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
        RECS_CUDA(c, cudaStreamSynchronize(c->stream));
        return RECS_OK;
    }
    if (c->stage_used + need > c->stage_half) {
        // this half is full: mark where its copies end, move to the other half once ITS copies are done
        RECS_CUDA(c, cudaEventRecord(c->stage_ev[c->stage_cur], c->stream));
        c->stage_busy[c->stage_cur] = true;
        c->stage_cur ^= 1;
        c->stage_used = 0;
        if (c->stage_busy[c->stage_cur]) {
            RECS_CUDA(c, cudaEventSynchronize(c->stage_ev[c->stage_cur]));
            c->stage_busy[c->stage_cur] = false;
        }
    }
    unsigned char* at = c->stage + (size_t)c->stage_cur * c->stage_half + c->stage_used;
    memcpy(at, src, bytes);
    c->stage_used += need;
    RECS_CUDA(c, cudaMemcpyAsync(dst, at, bytes, cudaMemcpyHostToDevice, c->stream));
    return RECS_OK;
}

// Equal slices of the padded body range, one per rank: whole chunks of the summation tree (kernels.cuh), so a rank's
// own tiles and the other ranks' tiles are two runs of whole chunks and the result does not depend on the rank count.
void shard_rows_for(int rank, int world_size, uint64_t n, uint64_t* slice, uint64_t* rb, uint64_t* re) {
    const uint64_t R = world_size > 0 ? (uint64_t)world_size : 1;
    const uint64_t K = gravity_chunk_tiles(n);
    const uint64_t tiles = (n + kGravTile - 1) / kGravTile;
    uint64_t slice_tiles = ((tiles + R - 1) / R + K - 1) / K * K;
    if (slice_tiles == 0) slice_tiles = K;
    const uint64_t s = slice_tiles * kGravTile;
    *slice = s;
    uint64_t b = s * (uint64_t)rank, e = b + s;
    if (b > n) b = n;
    if (e > n) e = n;
    *rb = b;
    *re = e;
}
void shard_rows(const recs_gpu_config& cfg, uint64_t n, uint64_t* slice, uint64_t* rb, uint64_t* re) {
    shard_rows_for(cfg.rank, cfg.world_size, n, slice, rb, re);
}

bool posm_is_current(recs_gpu_ctx* c);

// ---- systems -----------------------------------------------------------------------------------
// `movement` / `acceleration`: the systems fused into the tail of the reduce kernel, or null.
int32_t run_gravity(recs_gpu_ctx* c, System& sys, System* movement = nullptr, System* acceleration = nullptr) {
    const uint64_t POS = sys.comp[0], ACC = sys.comp[1], MASS = sys.comp[2];
    const bool sharded = c->cfg.world_size > 1;
    const bool p2p = sharded && c->xchg.imported && !c->exchange_nccl_forced;
    // 1. the nested Query<(Read<Position>, Read<Mass>)>: concatenation of its archetypes
    std::vector<std::pair<Column*, Column*>> q;
    uint64_t total_j = 0;
    std::vector<std::pair<ColKey, uint64_t>> sources;
    for (uint32_t a : sys.qarchs) {
        Column *p, *m;
        int32_t st;
        if ((st = need_column(c, a, POS, 12, &p)) != RECS_OK) return st;
        if ((st = need_column(c, a, MASS, 4, &m)) != RECS_OK) return st;
        if ((st = same_count(c, p, m)) != RECS_OK) return st;
        q.push_back(std::make_pair(p, m));
        total_j += p->count;
        sources.push_back(std::make_pair(ColKey(a, POS), p->version));
        sources.push_back(std::make_pair(ColKey(a, MASS), m->version));
    }
    if (sharded && (sys.qarchs.size() != 1 || sys.archs.size() != 1 || sys.qarchs[0] != sys.archs[0]))
        return fail(c, RECS_ERR_UNSUPPORTED, "sharded gravity needs exactly one body archetype");
    if (total_j >= (1ull << 31)) return fail(c, RECS_ERR_UNSUPPORTED, "more than 2^31 bodies");
    if (sharded && c->cfg.world_size > kGravMaxPeers) return fail(c, RECS_ERR_UNSUPPORTED, "more than 16 ranks");

    // the mirror's tile layout: whole chunks of the summation tree, equal slices per rank
    const int lay_R = sharded ? c->cfg.world_size : (c->emulate_ranks > 1 ? c->emulate_ranks : 1);
    const int lay_r = sharded ? c->cfg.rank : (c->emulate_ranks > 1 ? c->emulate_rank : 0);
    uint64_t slice = 0, rb = 0, re = 0;
    shard_rows_for(lay_r, lay_R, total_j, &slice, &rb, &re);
    if (!sharded) rb = 0, re = total_j;  // (an emulated rank still computes every row)
    const uint32_t K = gravity_chunk_tiles(total_j);
    const uint64_t n_jpad = slice * (uint64_t)lay_R;
    const uint32_t n_tiles = (uint32_t)(n_jpad / kGravTile);
    const uint32_t slice_tiles = (uint32_t)(slice / kGravTile);
    if (n_tiles / K > 64) return fail(c, RECS_ERR_UNSUPPORTED, "summation tree has more than 64 chunks");  // (kGravMaxChunks of the reduce kernel)

    int32_t st = RECS_OK;
    if (p2p) {
        if (n_jpad > c->xchg.cap_bodies)
            return fail(c, RECS_ERR_SIZE_MISMATCH, "the exported exchange buffer is smaller than the body count: export / import again");
    } else if ((st = ensure(c, &c->posm, &c->posm_cap, std::max<size_t>(n_jpad, kGravTile) * 16)) != RECS_OK) {
        return st;
    }

    // kernel shape by outer-row count (profiles/r01_gravity_small_n_sweep.txt, r01_gravity_variant_sweep5.txt):
    // small i-blocks give the stream-K split more units, and below 5 120 bodies the redo of the tile that holds a
    // warp's own bodies is a large enough share that masking every pair exactly is cheaper
    int variant = c->grav_variant;
    if (!c->grav_variant_forced) {
        uint64_t max_rows = 0;
        for (uint32_t a : sys.archs) {
            Column* p = find_column(c, a, POS);
            if (p) max_rows = std::max<uint64_t>(max_rows, sharded ? re - rb : p->count);
        }
        // The j-packed small shape has its own arithmetic (two bodies j per lane pair, exact mask on every pair), the
        // i-packed shapes share theirs: the small shape is chosen by the size of the nested query alone, so that a rank of
        // a sharded run uses the arithmetic the single-GPU run uses and results stay bit-identical across rank counts.
        if (total_j <= kGravSmallRows) variant = kGravVariantSmall;
        else if (max_rows <= kGravMediumRows) variant = kGravVariantMedium;
    }
    const bool scaled = gravity_variant(variant).scaled;
    const float pscale = scaled ? kGravPosScale : 1.f;

    auto peer_push = [&](uint32_t gen) {
        PeerPush pp = no_peer_push();
        pp.n = (uint32_t)c->cfg.world_size;
        pp.gen = gen;
        for (int r = 0; r < c->cfg.world_size; ++r) {
            pp.posm[r] = c->xchg.buffer(r, gen);
            pp.flag[r] = c->xchg.flags(r) + c->cfg.rank;
        }
        pp.local_flags = c->xchg.flags(c->cfg.rank);
        pp.done_counter = c->xchg.done_counter();
        pp.error = c->dev_error;
        pp.wait = pp.signal = 1u;
        return pp;
    };

    // 2. refresh the posm mirror if any source column changed since it was packed
    const bool dirty = sources != c->posm_sources || c->posm_bodies != n_jpad || c->posm_scaled != scaled;
    if (dirty) {
        Timed t(c, T_PACK);
        if (!sharded) {
            uint64_t off = 0;
            for (size_t qi = 0; qi < q.size(); ++qi) {
                auto& pm = q[qi];
                const uint64_t pad = qi + 1 == q.size() ? n_jpad - total_j : 0;
                RECS_CUDA(c, launch_pack_posm((const float*)pm.first->dev, (const float*)pm.second->dev, c->cfg.nbody_g,
                                              c->posm, 0, pm.first->count, pad, off, pscale, no_peer_push(), c->stream));
                off += pm.first->count;
                c->kernel_launches += pm.first->count + pad ? 1 : 0;
            }
        } else {
            // own rows (and the null bodies that fill the slice) only: the other slices arrive from their owners
            const uint64_t own_end = slice * ((uint64_t)c->cfg.rank + 1);
            const uint64_t own_begin = slice * (uint64_t)c->cfg.rank;
            const PeerPush pp = p2p ? peer_push(++c->xchg.gen) : no_peer_push();
            RECS_CUDA(c, launch_pack_posm((const float*)q[0].first->dev, (const float*)q[0].second->dev, c->cfg.nbody_g,
                                          c->posm, rb, re - rb, own_end - std::max(re, own_begin), std::max(rb, own_begin),
                                          pscale, pp, c->stream));
            c->kernel_launches++;
        }
        c->posm_sources = sources;
        c->posm_bodies = n_jpad;
        c->posm_scaled = scaled;
        c->posm_pack_serial++;
    } else if (c->capturing) {
        c->graph_elided_pack = true;
    }
    if (sharded && !p2p && !c->comm)
        return fail(c, RECS_ERR_NCCL, "world_size > 1 but neither recs_gpu_exchange_import nor recs_gpu_comm_init was called");
    const float* posm_now = p2p ? c->xchg.buffer(c->cfg.rank, c->xchg.gen) : c->posm;
    // NCCL exchange: the all-gather is enqueued AFTER the own-tile launch: an NCCL kernel that is resident
    // while it waits for a late peer would keep persistent gravity CTAs off its SMs.
    auto enqueue_gather = [&]() -> int32_t {
        RECS_CUDA(c, cudaEventRecord(c->ev_pack, c->stream));  // after the own-tile launch (and the pack before it)
        RECS_CUDA(c, cudaStreamWaitEvent(c->comm_stream, c->ev_pack, 0));
        float* send = c->posm + (size_t)slice * (uint64_t)c->cfg.rank * 4;
        TimedSpan span;
        const bool timed = c->timing && !c->capturing;
        if (timed) {
            span.a = get_event(c);
            span.b = get_event(c);
            span.slot = T_GATHER;
            cudaEventRecord(span.a, c->comm_stream);
        }
        int r = c->nccl.AllGather(send, c->posm, (size_t)slice * 4, /*ncclFloat32*/ 7, c->comm, c->comm_stream);
        if (r != 0) return fail(c, RECS_ERR_NCCL, std::string("ncclAllGather: ") + c->nccl.GetErrorString(r));
        if (timed) {
            cudaEventRecord(span.b, c->comm_stream);
            c->spans.push_back(span);
        }
        RECS_CUDA(c, cudaEventRecord(c->ev_gather, c->comm_stream));
        return RECS_OK;
    };

    // 3. outer iteration: Read<Position>, Write<Acceleration> over the system's archetypes.  Every archetype's
    //    interaction launch comes before any reduce: the fused tail of one archetype moves bodies that the nested
    //    query of the next one still has to see where they were (gravity reads Position, movement writes it).
    struct Outer {
        uint32_t arch;
        Column *p, *acc;
        uint64_t row_begin, row_count;
        int variant;
        uint32_t iblock_rows;
        GravityWork work;
        size_t frag_base;  // in float4
        uint32_t self_offset;
    };
    std::vector<Outer> outers;
    size_t frag_float4 = 0;
    auto add_outer = [&](uint32_t arch, Column* p, Column* acc, uint64_t row_begin, uint64_t row_count, int v, uint32_t self_offset) -> int32_t {
        Outer o;
        o.arch = arch;
        o.p = p;
        o.acc = acc;
        o.row_begin = row_begin;
        o.row_count = row_count;
        o.variant = v;
        o.iblock_rows = (uint32_t)gravity_variant(v).rows();
        o.self_offset = self_offset;
        GravityWork& w = o.work;
        memset(&w, 0, sizeof w);
        w.n_iblocks = (uint32_t)((row_count + o.iblock_rows - 1) / o.iblock_rows);
        w.n_tiles_total = n_tiles;
        w.chunk_tiles = K;
        if ((uint64_t)w.n_iblocks * n_tiles >= (1ull << 32)) return fail(c, RECS_ERR_UNSUPPORTED, "work-unit count overflows 32 bits");
        // tile ranges: the rank's own slice first, then the other ranks' slices starting behind it (one range on one GPU)
        struct TileRange {
            uint32_t n_vtiles, tile_offset;
        } ranges[2];
        uint32_t n_ranges = 1;
        if (lay_R > 1) {
            n_ranges = 2;
            ranges[0] = {slice_tiles, slice_tiles * (uint32_t)lay_r};
            ranges[1] = {n_tiles - slice_tiles, (slice_tiles * (uint32_t)lay_r + slice_tiles) % n_tiles};
        } else {
            ranges[0] = {n_tiles, 0u};
        }
        // i-block ranges: a partly filled last i-block of the 1536-row shapes is dealt by itself (its warps without rows
        // skip the arithmetic: cheaper units, spread evenly over all CTAs).  Worth its extra chunk-sum launch when one
        // i-block is a visible share of the rank's work: 0.7 % of the tick at 65 536 bodies and on a rank of an 8-GPU run.
        const uint64_t rest = row_count % o.iblock_rows;
        const bool own_last = !c->grav_no_heavy && gravity_variant(v).rows_per_thread == 4 && gravity_variant(v).scaled && rest != 0 &&
                              rest <= (uint64_t)o.iblock_rows * 3 / 4 && w.n_iblocks >= 16 && w.n_iblocks <= 256;
        // (its rows have up to K - 1 tile fragments per chunk; the reduce kernel's one-thread-per-row walk is too serial
        // for them even at K = 8 -- +48 us at 65 536 bodies -- so their chunk sums come from nbody_gravity_chunksum_kernel)
        w.heavy_iblock = own_last ? w.n_iblocks - 1 : 0xffffffffu;
        w.n_phases = 0;
        w.n_own_phases = 0;
        for (uint32_t tr = 0; tr < n_ranges; ++tr) {
            for (uint32_t part = 0; part < (own_last ? 2u : 1u); ++part) {
                GravityPhase& P = w.phase[w.n_phases++];
                P.n_vtiles = ranges[tr].n_vtiles;
                P.tile_offset = ranges[tr].tile_offset;
                P.iblock_begin = part == 0 ? 0 : w.n_iblocks - 1;
                P.n_iblocks = own_last ? (part == 0 ? w.n_iblocks - 1 : 1) : w.n_iblocks;
            }
            if (tr == 0) w.n_own_phases = w.n_phases;
        }
        uint64_t max_units = 1;
        for (uint32_t ph = 0; ph < w.n_phases; ++ph) max_units = std::max<uint64_t>(max_units, (uint64_t)w.phase[ph].n_iblocks * w.phase[ph].n_vtiles);
        w.grid = (uint32_t)std::min<uint64_t>((uint64_t)c->sm_count * c->grav_ctas_per_sm[v], max_units);
        if (c->grav_grid_forced > 0) w.grid = (uint32_t)std::min<uint64_t>((uint64_t)c->grav_grid_forced, max_units);
        o.frag_base = frag_float4;
        frag_float4 += (size_t)gravity_assign_frag_slots(w) * o.iblock_rows;
        outers.push_back(o);
        return RECS_OK;
    };
    for (uint32_t a : sys.archs) {
        Column *p, *acc;
        if ((st = need_column(c, a, POS, 12, &p)) != RECS_OK) return st;
        if ((st = need_column(c, a, ACC, 12, &acc)) != RECS_OK) return st;
        if ((st = same_count(c, p, acc)) != RECS_OK) return st;
        const uint64_t row_begin = sharded ? rb : 0;
        const uint64_t row_count = sharded ? re - rb : p->count;
        if (row_count == 0 && !sharded) continue;
        if (row_count >= (1ull << 31)) return fail(c, RECS_ERR_UNSUPPORTED, "more than 2^31 rows");
        // mirror index of outer row 0 when this archetype is part of the nested query
        uint32_t self_offset = 0xffffffffu;
        uint64_t off = 0;
        for (size_t qi = 0; qi < sys.qarchs.size(); ++qi) {
            if (sys.qarchs[qi] == a) {
                self_offset = (uint32_t)off;
                break;
            }
            off += q[qi].first->count;
        }
        // (A partly filled last i-block is handled inside add_outer: phases of its own.  Giving its rows to a second, small
        // launch of the 256-row shape was measured first and dropped -- profiles/r02_remainder_launch.txt.)
        if ((st = add_outer(a, p, acc, row_begin, row_count, variant, self_offset)) != RECS_OK) return st;
    }
    if ((st = ensure(c, &c->frags, &c->frags_cap, std::max<size_t>(frag_float4, 1) * sizeof(float4))) != RECS_OK) return st;

    // one launch for all phases, except with the NCCL exchange: own tiles, all-gather, the other tiles
    const bool split = sharded && !p2p;
    for (uint32_t l = 0; l < (split ? 2u : 1u); ++l) {
        if (split && l == 1) {
            if ((st = enqueue_gather()) != RECS_OK) return st;
            RECS_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_gather, 0));
        }
        for (Outer& o : outers) {
            if (o.row_count == 0) continue;
            GravityArgs ga;
            memset(&ga, 0, sizeof ga);
            ga.posm = reinterpret_cast<const float4*>(posm_now);
            ga.ipos = (const float*)o.p->dev;
            ga.frags = c->frags + o.frag_base;
            ga.row_begin = (uint32_t)o.row_begin;
            ga.row_count = (uint32_t)o.row_count;
            ga.eps = c->cfg.nbody_epsilon;
            ga.self_offset = o.self_offset;
            ga.work = o.work;
            ga.variant = o.variant;
            ga.tiles_per_rank = slice_tiles;
            ga.own_rank = (uint32_t)lay_r;
            ga.error = c->dev_error;
            if (p2p) {
                ga.flags = c->xchg.flags(c->cfg.rank);
                ga.wait_gen = c->xchg.gen;
            }
            ga.phase_begin = split && l == 1 ? o.work.n_own_phases : 0;
            ga.phase_end = split && l == 0 ? o.work.n_own_phases : o.work.n_phases;
            Timed t(c, T_GRAVITY);
            RECS_CUDA(c, launch_gravity(ga, c->stream));
            c->kernel_launches++;
            c->gravity_launches++;
        }
    }
    // the tail may be fused when it cannot disturb a later read of this tick: every interaction launch is enqueued
    uint32_t push_gen = 0;
    for (size_t oi = 0; oi < outers.size(); ++oi) {
        Outer& o = outers[oi];
        GravityReduceArgs ra;
        memset(&ra, 0, sizeof ra);
        ra.frags = c->frags + o.frag_base;
        ra.acc = (float*)o.acc->dev;
        ra.row_begin = (uint32_t)o.row_begin;
        ra.row_count = (uint32_t)o.row_count;
        ra.iblock_rows = o.iblock_rows;
        ra.work = o.work;
        Column *tail_pos = nullptr, *tail_vel = nullptr;
        if (movement) {
            Column* mass_col = nullptr;
            if ((st = need_column(c, o.arch, movement->comp[0], 12, &tail_pos)) != RECS_OK) return st;
            if ((st = need_column(c, o.arch, movement->comp[1], 12, &tail_vel)) != RECS_OK) return st;
            if ((st = same_count(c, tail_pos, tail_vel)) != RECS_OK) return st;
            ra.tail.pos = (float*)tail_pos->dev;
            ra.tail.vel = (float*)tail_vel->dev;
            ra.tail.dt = c->cfg.nbody_dt;
            ra.tail.g = c->cfg.nbody_g;
            ra.tail.pscale = pscale;
            // refresh the mirror in place when this archetype is part of the nested query
            if (o.self_offset != 0xffffffffu && movement->comp[0] == POS) {
                if ((st = need_column(c, o.arch, MASS, 4, &mass_col)) != RECS_OK) return st;
                ra.tail.mass = (const float*)mass_col->dev;
                ra.tail.posm_offset = o.self_offset;
                if (p2p) {
                    // the refreshed rows go to every rank's buffer of the next generation: ONE generation per tick, the
                    // first pushing launch waits for the peers, the last one raises the flag
                    if (!push_gen) push_gen = ++c->xchg.gen;
                    ra.push = peer_push(push_gen);
                    ra.push.wait = oi == 0 ? 1u : 0u;
                    ra.push.signal = oi + 1 == outers.size() ? 1u : 0u;
                    ra.tail.posm = ra.push.posm[c->cfg.rank];
                } else {
                    ra.tail.posm = c->posm;
                }
            }
        }
        {
            Timed t(c, T_REDUCE);
            if (o.work.heavy_iblock != 0xffffffffu && o.row_count) {
                RECS_CUDA(c, launch_gravity_chunksum(ra, c->stream));
                c->kernel_launches++;
            }
            RECS_CUDA(c, launch_gravity_reduce(ra, c->stream));
            c->kernel_launches++;
        }
        if (oi + 1 == outers.size() || outers[oi + 1].acc != o.acc) o.acc->version++;
        if (movement && (oi + 1 == outers.size() || outers[oi + 1].arch != o.arch)) {
            tail_pos->version++;
            tail_vel->version++;
        }
    }
    if (movement && movement->comp[0] == POS && sys.archs == sys.qarchs) {
        // every source of the mirror was just refreshed by the tail: record it as current
        for (auto& src : c->posm_sources) src.second = find_column(c, src.first.first, src.first.second)->version;
    }
    if (sharded && !p2p && outers.empty()) {  // nothing launched here: still take part in the collective
        if ((st = enqueue_gather()) != RECS_OK) return st;
        RECS_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_gather, 0));
    }
    return RECS_OK;
}

// (Write<Y>, Read<X>) with y += x * a over each matched archetype's rows.
int32_t run_axpy(recs_gpu_ctx* c, System& sys, uint64_t Y, uint64_t X, float a, int slot, bool shard) {
    for (uint32_t arch : sys.archs) {
        Column *y, *x;
        int32_t st;
        if ((st = need_column(c, arch, Y, 12, &y)) != RECS_OK) return st;
        if ((st = need_column(c, arch, X, 12, &x)) != RECS_OK) return st;
        if ((st = same_count(c, y, x)) != RECS_OK) return st;
        uint64_t rb = 0, re = y->count, slice;
        if (shard && c->cfg.world_size > 1) shard_rows(c->cfg, y->count, &slice, &rb, &re);
        if (re == rb) continue;
        Timed t(c, slot);
        RECS_CUDA(c, launch_axpy_rn((float*)y->dev + 3 * rb, (const float*)x->dev + 3 * rb, a, 3 * (re - rb),
                                    c->stream));
        c->kernel_launches++;
        y->version++;
    }
    return RECS_OK;
}

void rotate_wind(recs_gpu_ctx* c) {
    // crates/rain-simulation/src/lib.rs:78-84 with cgmath 0.18 semantics:
    // Quaternion::from_axis_angle(unit_y, Deg(10) * dt).rotate_vector(dir)
    const float deg = c->cfg.rain_wind_deg_per_s * c->cfg.rain_dt;
    const float rad = deg * (float)(3.14159265358979323846 / 180.0);
    const float half = rad * 0.5f;
    const float s = sinf(half), qs = cosf(half);
    const float qv[3] = {0.f * s, 1.f * s, 0.f * s};
    const float* v = c->wind;
    // tmp = qv x v + v * qs
    const float t[3] = {(qv[1] * v[2] - qv[2] * v[1]) + v[0] * qs, (qv[2] * v[0] - qv[0] * v[2]) + v[1] * qs,
                        (qv[0] * v[1] - qv[1] * v[0]) + v[2] * qs};
    // (qv x tmp) * 2 + v
    const float r[3] = {(qv[1] * t[2] - qv[2] * t[1]) * 2.f + v[0], (qv[2] * t[0] - qv[0] * t[2]) * 2.f + v[1],
                        (qv[0] * t[1] - qv[1] * t[0]) * 2.f + v[2]};
    c->wind[0] = r[0];
    c->wind[1] = r[1];
    c->wind[2] = r[2];
}

int32_t run_rain_vel(recs_gpu_ctx* c, System& sys, bool wind) {
    const float k_gravity = (-c->cfg.rain_gravity) * c->cfg.rain_dt;
    const float cw = c->cfg.rain_wind_accel * c->cfg.rain_dt;
    for (uint32_t arch : sys.archs) {
        Column* v;
        int32_t st;
        if ((st = need_column(c, arch, sys.comp[0], 12, &v)) != RECS_OK) return st;
        // independent rows, no exchange: every rank streams its own contiguous slice
        uint64_t rb = 0, re = v->count, slice;
        if (c->cfg.world_size > 1) shard_rows(c->cfg, v->count, &slice, &rb, &re);
        if (re == rb) continue;
        Timed t(c, T_OTHER);
        if (wind)
            RECS_CUDA(c, launch_rain_wind((float*)v->dev + 3 * rb, re - rb, cw * c->wind[0], cw * c->wind[1],
                                          cw * c->wind[2], c->cfg.rain_terminal, c->stream));
        else
            RECS_CUDA(c, launch_rain_gravity((float*)v->dev + 3 * rb, re - rb, k_gravity, c->cfg.rain_terminal, c->stream));
        c->kernel_launches++;
        v->version++;
    }
    return RECS_OK;
}

// collision / ground_collision: one flag per outer row, then a compacted list of flagged rows.
int32_t run_removal_system(recs_gpu_ctx* c, System& sys) {
    if (c->cfg.world_size > 1) return fail(c, RECS_ERR_UNSUPPORTED, "removal-deciding systems are single-GPU");
    const uint64_t POS = sys.comp[0];
    const float r2 = c->cfg.nbody_collision_radius * c->cfg.nbody_collision_radius;  // COLLISION_RADIUS * COLLISION_RADIUS
    for (auto& kv : sys.removals) kv.second.valid = false;
    for (uint32_t a : sys.archs) {
        Column* p;
        int32_t st;
        if ((st = need_column(c, a, POS, 12, &p)) != RECS_OK) return st;
        if (p->count >= (1ull << 31)) return fail(c, RECS_ERR_UNSUPPORTED, "more than 2^31 rows");
        const uint32_t n = (uint32_t)p->count;
        System::Removals& r = sys.removals[a];
        if (r.cap_rows < std::max<uint32_t>(n, 1)) {
            if (c->capturing) return fail(c, RECS_ERR_CUDA, "workspace growth during graph capture");
            if (r.flags) RECS_CUDA(c, cudaFree(r.flags));
            if (r.rows) RECS_CUDA(c, cudaFree(r.rows));
            if (!r.counter) RECS_CUDA(c, cudaMalloc((void**)&r.counter, sizeof(uint32_t)));
            size_t cap = 256;
            while (cap < n) cap <<= 1;
            RECS_CUDA(c, cudaMalloc((void**)&r.flags, cap * sizeof(uint32_t)));
            RECS_CUDA(c, cudaMalloc((void**)&r.rows, cap * sizeof(uint32_t)));
            r.cap_rows = cap;
        }
        r.n_rows = n;
        RECS_CUDA(c, cudaMemsetAsync(r.counter, 0, sizeof(uint32_t), c->stream));
        if (n == 0) {
            r.valid = true;
            continue;
        }
        Timed t(c, T_OTHER);
        if (sys.kind == RECS_SYS_NBODY_COLLISION) {
            RECS_CUDA(c, cudaMemsetAsync(r.flags, 0, (size_t)n * sizeof(uint32_t), c->stream));
            for (uint32_t qa : sys.qarchs) {
                Column* q;
                if ((st = need_column(c, qa, POS, 12, &q)) != RECS_OK) return st;
                RECS_CUDA(c, launch_collision_flags((const float*)q->dev, (uint32_t)q->count, (const float*)p->dev, n,
                                                    qa == a ? 1 : 0, r2, r.flags, c->stream));
                c->kernel_launches += q->count ? 1 : 0;
            }
        } else {
            RECS_CUDA(c, launch_below_flags((const float*)p->dev, n, /*y*/ 1, c->cfg.rain_ground_height, r.flags, c->stream));
            c->kernel_launches++;
        }
        RECS_CUDA(c, launch_compact_flags(r.flags, n, r.rows, r.counter, (uint32_t)r.cap_rows, c->stream));
        c->kernel_launches++;
        r.valid = true;
    }
    return RECS_OK;
}

// Generic system: one launch over the concatenation of the matched archetypes.
int32_t run_kernel_system(recs_gpu_ctx* c, System& sys) {
    KernelSys& k = *sys.ks;
    const size_t n_params = k.params.size();
    if (k.table_epoch != c->epoch) {
        // (re)build the table: work-item prefix, entity counts, column base pointers per parameter
        if (c->capturing) return fail(c, RECS_ERR_CUDA, "archetype table rebuild during graph capture");
        std::vector<uint32_t> archs;
        std::vector<std::vector<Column*>> cols;
        std::vector<std::pair<uint64_t, uint64_t>> ranges;
        for (uint32_t a : sys.archs) {
            std::vector<Column*> row(n_params, nullptr);
            int32_t st;
            Column* first = nullptr;
            for (size_t p = 0; p < n_params; ++p) {
                if (k.params[p].kind == RECS_KPARAM_COMMANDS) continue;
                if ((st = need_column(c, a, k.comp[p], k.params[p].elem_size, &row[p])) != RECS_OK) return st;
                if (!first) first = row[p];
                if ((st = same_count(c, first, row[p])) != RECS_OK) return st;
            }
            if (!first) return fail(c, RECS_ERR_UNSUPPORTED, "a generic system needs at least one Read / Write / Entity parameter");
            for (size_t p = 0; p < n_params; ++p)
                if (!row[p]) row[p] = first;  // Commands: counted like the first column
            // sharded contexts: every rank runs the body over its own contiguous slice of each archetype
            uint64_t rb = 0, re = row[0]->count, slice;
            if (c->cfg.world_size > 1) shard_rows(c->cfg, row[0]->count, &slice, &rb, &re);
            if (re == rb) continue;  // empty archetypes (or slices) take no work items
            archs.push_back(a);
            cols.push_back(row);
            ranges.push_back(std::make_pair(rb, re));
        }
        const size_t n_arch = archs.size();
        std::vector<unsigned long long>& t = k.table_host;
        t.assign(2 * n_arch + 1 + n_params * n_arch + 1, 0ull);  // + the create buffer
        unsigned long long items = 0;
        for (size_t i = 0; i < n_arch; ++i) {
            const unsigned long long count = ranges[i].second - ranges[i].first;
            t[i] = items;
            items += k.plan.staged ? (count + k.plan.chunk - 1) / k.plan.chunk : count;
            t[n_arch + 1 + i] = count;
            for (size_t p = 0; p < n_params; ++p) {
                if (k.params[p].kind == RECS_KPARAM_COMMANDS) {
                    // removal flags of this archetype (grown here, with the table, never during a captured tick)
                    System::Removals& r = sys.removals[archs[i]];
                    const uint32_t n = (uint32_t)cols[i][p]->count;
                    if (r.cap_rows < std::max<uint32_t>(n, 1)) {
                        if (r.flags) RECS_CUDA(c, cudaFree(r.flags));
                        if (r.rows) RECS_CUDA(c, cudaFree(r.rows));
                        if (!r.counter) RECS_CUDA(c, cudaMalloc((void**)&r.counter, sizeof(uint32_t)));
                        size_t cap = 256;
                        while (cap < n) cap <<= 1;
                        RECS_CUDA(c, cudaMalloc((void**)&r.flags, cap * sizeof(uint32_t)));
                        RECS_CUDA(c, cudaMalloc((void**)&r.rows, cap * sizeof(uint32_t)));
                        r.cap_rows = cap;
                    }
                    r.n_rows = n;
                    t[2 * n_arch + 1 + p * n_arch + i] = (unsigned long long)(uintptr_t)(r.flags + ranges[i].first);
                    continue;
                }
                t[2 * n_arch + 1 + p * n_arch + i] =
                    (unsigned long long)(uintptr_t)((unsigned char*)cols[i][p]->dev + ranges[i].first * k.params[p].elem_size);
            }
        }
        t[n_arch] = items;
        if (k.create_bytes) {
            const size_t bytes = 16 + (size_t)k.create_cap * k.create_stride();
            int32_t cst = ensure(c, &k.create_buf, &k.create_buf_cap, bytes);
            if (cst != RECS_OK) return cst;
            const uint32_t header[4] = {0u, k.create_cap, k.create_bytes, 0u};
            if ((cst = staged_h2d(c, k.create_buf, header, sizeof header)) != RECS_OK) return cst;
            t[2 * n_arch + 1 + n_params * n_arch] = (unsigned long long)(uintptr_t)k.create_buf;
        }
        k.footprint_bytes = 0;
        for (size_t i = 0; i < n_arch; ++i)
            for (size_t p = 0; p < n_params; ++p)
                if (k.params[p].kind != RECS_KPARAM_COMMANDS)
                    k.footprint_bytes += (ranges[i].second - ranges[i].first) * k.params[p].elem_size;
        int32_t st = ensure(c, &k.table, &k.table_cap, std::max<size_t>(t.size() * sizeof(unsigned long long), 64));
        if (st != RECS_OK) return st;
        if ((st = staged_h2d(c, k.table, t.data(), t.size() * sizeof(unsigned long long))) != RECS_OK) return st;
        k.n_items = items;
        k.n_arch = (uint32_t)n_arch;
        k.table_archs = archs;
        if (k.is_pair) {
            // the nested query: its own archetype list, every row of every matched archetype, in order
            const size_t nq = k.qparams.size();
            std::vector<std::vector<Column*>> qcols;
            for (uint32_t a : sys.qarchs) {
                std::vector<Column*> row(nq, nullptr);
                for (size_t p = 0; p < nq; ++p) {
                    if ((st = need_column(c, a, k.qcomp[p], k.qparams[p].elem_size, &row[p])) != RECS_OK) return st;
                    if ((st = same_count(c, row[0], row[p])) != RECS_OK) return st;
                }
                if (row[0]->count) qcols.push_back(row);
            }
            const size_t nqa = qcols.size();
            std::vector<unsigned long long>& q = k.qtable_host;
            q.assign(2 * nqa + 1 + nq * nqa, 0ull);
            unsigned long long rows = 0;
            for (size_t i = 0; i < nqa; ++i) {
                q[i] = rows;
                rows += qcols[i][0]->count;
                q[nqa + 1 + i] = qcols[i][0]->count;
                for (size_t p = 0; p < nq; ++p) q[2 * nqa + 1 + p * nqa + i] = (unsigned long long)(uintptr_t)qcols[i][p]->dev;
            }
            q[nqa] = rows;
            if ((st = ensure(c, &k.qtable, &k.qtable_cap, std::max<size_t>(q.size() * sizeof(unsigned long long), 64))) != RECS_OK) return st;
            if ((st = staged_h2d(c, k.qtable, q.data(), q.size() * sizeof(unsigned long long))) != RECS_OK) return st;
            k.n_qarch = (uint32_t)nqa;
        }
        k.table_epoch = c->epoch;
    }
    if (k.has_commands) {
        for (auto& kv : sys.removals) kv.second.valid = false;
        for (uint32_t a : sys.archs) {  // archetypes without rows still answer recs_gpu_system_removals (with nothing)
            System::Removals& r = sys.removals[a];
            if (!r.counter) {
                if (c->capturing) return fail(c, RECS_ERR_CUDA, "workspace growth during graph capture");
                RECS_CUDA(c, cudaMalloc((void**)&r.counter, sizeof(uint32_t)));
            }
            RECS_CUDA(c, cudaMemsetAsync(r.counter, 0, sizeof(uint32_t), c->stream));
            r.valid = true;
        }
        for (uint32_t a : k.table_archs) {
            System::Removals& r = sys.removals[a];
            RECS_CUDA(c, cudaMemsetAsync(r.flags, 0, (size_t)r.n_rows * sizeof(uint32_t), c->stream));
        }
    }
    if (k.create_bytes) {
        RECS_CUDA(c, cudaMemsetAsync(k.create_buf, 0, sizeof(uint32_t), c->stream));  // count of this run
        k.create_valid = true;
    }
    if (k.n_items == 0) return RECS_OK;
    // grid: persistent CTAs for the pipelined form (three per SM while shared memory allows), one resident wave of
    // grid-stride CTAs for the direct form
    uint64_t ctas, cap;
    if (k.plan.staged) {
        // Persistent CTAs per SM: three while the touched columns fit L2, ONE once the launch is HBM-bound -- fewer
        // concurrent chunk streams keep DRAM pages open (simple_iter 10 M entities: 6.25 TB/s with one CTA per SM,
        // 6.09 with two, 5.96 with three; profiles/r01_generic_systems.txt).
        static const int forced = getenv("RECS_GPU_KERNEL_CTAS") ? atoi(getenv("RECS_GPU_KERNEL_CTAS")) : 0;  // tuning knob
        uint64_t per_sm = std::max<uint64_t>(1, std::min<uint64_t>(3, (227u * 1024u) / (k.plan.smem_bytes + 2048u)));
        if (k.footprint_bytes > kStreamL2Footprint) per_sm = 1;
        if (forced > 0) per_sm = std::min<uint64_t>(per_sm, (uint64_t)forced);
        ctas = k.n_items;
        cap = (uint64_t)c->sm_count * per_sm;
    } else {
        const uint64_t per_cta = k.is_pair ? k.plan.chunk : k.plan.threads;  // pair kernels: several outer entities per thread
        ctas = (k.n_items + per_cta - 1) / per_cta;
        cap = (uint64_t)c->sm_count * (2048 / k.plan.threads);
    }
    const unsigned grid = (unsigned)std::min<uint64_t>(ctas, cap);
    const unsigned long long* table = k.table;
    unsigned n_arch = k.n_arch;
    unsigned long long n_items = k.n_items;
    const unsigned long long* qtable = k.qtable;
    unsigned n_qarch = k.n_qarch;
    void* args[6] = {(void*)&table, (void*)&n_arch, (void*)&n_items, nullptr, nullptr, nullptr};
    int n_args = 3;
    if (k.is_pair) {
        args[n_args++] = (void*)&qtable;
        args[n_args++] = (void*)&n_qarch;
    }
    if (k.has_uniform) args[n_args++] = (void*)k.uniform.data();
    {
        Timed t(c, T_OTHER);
        RECS_CUDA(c, cudaLaunchKernel((const void*)k.kernel, dim3(grid), dim3(k.plan.threads), args, k.plan.smem_bytes,
                                      c->stream));
    }
    c->kernel_launches++;
    if (k.has_commands)
        for (uint32_t a : k.table_archs) {
            System::Removals& r = sys.removals[a];
            RECS_CUDA(c, launch_compact_flags(r.flags, r.n_rows, r.rows, r.counter, (uint32_t)r.cap_rows, c->stream));
            c->kernel_launches++;
        }
    for (uint32_t a : sys.archs)
        for (size_t p = 0; p < n_params; ++p)
            if (k.params[p].kind == RECS_KPARAM_WRITE) {
                Column* col = find_column(c, a, k.comp[p]);
                if (col) col->version++;
            }
    return RECS_OK;
}

// NVTX range per system run: the counterpart of the reference's `#[instrument]` tracing spans on its systems
// (crates/n-body/src/lib.rs:87,95,104); free when no profiler is attached.
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

int32_t run_system_locked(recs_gpu_ctx* c, uint32_t id) {
    if (id >= c->systems.size()) return fail(c, RECS_ERR_INVALID_ARGUMENT, "unknown system id");
    System& s = c->systems[id];
    NvtxRange range(s.name.c_str());
    switch (s.kind) {
        case RECS_SYS_NBODY_GRAVITY:
            return run_gravity(c, s);
        case RECS_SYS_NBODY_MOVEMENT:
            return run_axpy(c, s, s.comp[0], s.comp[1], c->cfg.nbody_dt, T_MOVEMENT, true);
        case RECS_SYS_NBODY_ACCELERATION:
            return run_axpy(c, s, s.comp[0], s.comp[1], c->cfg.nbody_dt, T_ACCEL, true);
        case RECS_SYS_RAIN_GRAVITY:
            return run_rain_vel(c, s, false);
        case RECS_SYS_RAIN_WIND:
            return run_rain_vel(c, s, true);
        case RECS_SYS_RAIN_MOVEMENT:
            return run_axpy(c, s, s.comp[0], s.comp[1], c->cfg.rain_dt, T_OTHER, true);
        case RECS_SYS_RAIN_WIND_DIRECTION:
            if (c->capturing) return fail(c, RECS_ERR_UNSUPPORTED, "wind_direction inside a captured tick");
            rotate_wind(c);
            return RECS_OK;
        case RECS_SYS_QUERY_ADD:
            return run_axpy(c, s, s.comp[0], s.comp[1], 1.0f, T_OTHER, true);
        case RECS_SYS_NBODY_COLLISION:
        case RECS_SYS_RAIN_GROUND_COLLISION:
            return run_removal_system(c, s);
        case RECS_SYS_KERNEL:
            return run_kernel_system(c, s);
    }
    return fail(c, RECS_ERR_INVALID_ARGUMENT, "unknown system kind");
}

// wind -> gravity -> movement adjacent in the tick order and over the same archetypes: one pass.
bool try_fused_rain(recs_gpu_ctx* c, size_t i, int32_t* st) {
    if (i + 2 >= c->order.size()) return false;
    System& w = c->systems[c->order[i]];
    System& g = c->systems[c->order[i + 1]];
    System& m = c->systems[c->order[i + 2]];
    if (w.kind != RECS_SYS_RAIN_WIND || g.kind != RECS_SYS_RAIN_GRAVITY || m.kind != RECS_SYS_RAIN_MOVEMENT)
        return false;
    if (w.archs != g.archs || w.archs != m.archs) return false;
    if (w.comp[0] != g.comp[0] || w.comp[0] != m.comp[1]) return false;
    const float k_gravity = (-c->cfg.rain_gravity) * c->cfg.rain_dt;
    const float cw = c->cfg.rain_wind_accel * c->cfg.rain_dt;
    *st = RECS_OK;
    for (uint32_t arch : w.archs) {
        Column *v, *p;
        if ((*st = need_column(c, arch, w.comp[0], 12, &v)) != RECS_OK) return true;
        if ((*st = need_column(c, arch, m.comp[0], 12, &p)) != RECS_OK) return true;
        if ((*st = same_count(c, v, p)) != RECS_OK) return true;
        uint64_t rb = 0, re = v->count, slice;
        if (c->cfg.world_size > 1) shard_rows(c->cfg, v->count, &slice, &rb, &re);
        if (re == rb) continue;
        Timed t(c, T_OTHER);
        cudaError_t e = launch_rain_fused((float*)v->dev + 3 * rb, (float*)p->dev + 3 * rb, re - rb, cw * c->wind[0], cw * c->wind[1],
                                          cw * c->wind[2], k_gravity, c->cfg.rain_terminal, c->cfg.rain_dt, c->stream);
        if (e != cudaSuccess) {
            *st = fail(c, RECS_ERR_CUDA, cudaGetErrorString(e));
            return true;
        }
        c->kernel_launches++;
        v->version++;
        p->version++;
    }
    return true;
}

// gravity -> movement -> acceleration adjacent in the tick order over the same archetypes and
// components: the two integrations run in the tail of the fragment-reduce kernel.
bool try_fused_nbody(recs_gpu_ctx* c, size_t i, int32_t* st) {
    if (i + 2 >= c->order.size()) return false;
    System& g = c->systems[c->order[i]];
    System& m = c->systems[c->order[i + 1]];
    System& a = c->systems[c->order[i + 2]];
    if (g.kind != RECS_SYS_NBODY_GRAVITY || m.kind != RECS_SYS_NBODY_MOVEMENT || a.kind != RECS_SYS_NBODY_ACCELERATION)
        return false;
    if (g.archs != m.archs || g.archs != a.archs) return false;
    // movement(Write<Position>, Read<Velocity>), acceleration(Write<Velocity>, Read<Acceleration>)
    if (m.comp[0] != g.comp[0] || a.comp[0] != m.comp[1] || a.comp[1] != g.comp[1]) return false;
    *st = run_gravity(c, g, &m, &a);
    return true;
}

int32_t run_tick_once(recs_gpu_ctx* c) {
    NvtxRange range("tick");  // tracy_client::frame_mark() per tick in the reference (schedule/mod.rs:226-227)
    for (size_t i = 0; i < c->order.size(); ++i) {
        int32_t st = RECS_OK;
        if (try_fused_nbody(c, i, &st)) {
            if (st != RECS_OK) return st;
            i += 2;
            continue;
        }
        if (try_fused_rain(c, i, &st)) {
            if (st != RECS_OK) return st;
            i += 2;
            continue;
        }
        st = run_system_locked(c, c->order[i]);
        if (st != RECS_OK) return st;
    }
    return RECS_OK;
}

bool order_is_capturable(recs_gpu_ctx* c) {
    for (uint32_t id : c->order)
        if (c->systems[id].kind == RECS_SYS_RAIN_WIND_DIRECTION || c->systems[id].kind == RECS_SYS_RAIN_WIND)
            return false;  // wind direction is a host-side per-tick value
    return true;
}

void drop_graph(recs_gpu_ctx* c) {
    if (c->graph) cudaGraphExecDestroy(c->graph);
    c->graph = nullptr;
    c->graph_epoch = ~0ull;
    c->graph_elided_pack = false;
}

// True while no source column of the gravity mirror was written since it was packed / refreshed.
bool posm_is_current(recs_gpu_ctx* c) {
    for (const auto& src : c->posm_sources) {
        const Column* col = find_column(c, src.first.first, src.first.second);
        if (!col || col->version != src.second) return false;
    }
    return true;
}

int32_t check_device_error(recs_gpu_ctx* c) {
    if (c->dev_error_host && *reinterpret_cast<volatile uint32_t*>(c->dev_error_host)) {
        *c->dev_error_host = 0;
        return fail(c, RECS_ERR_NCCL, c->cfg.world_size > 1
                                          ? "peer exchange timed out: a rank did not push its slice of the position mirror (ranks must run the same ticks)"
                                          : "a device-side wait timed out (grid barrier of the resident n-body ticks)");
    }
    return RECS_OK;
}

void fill_system(System& s, recs_gpu_system_kind kind, const uint64_t* ids) {
    auto acc = [&](uint64_t comp, bool w) {
        recs_gpu_access a;
        a.component_id = comp;
        a.is_write = w ? 1 : 0;
        s.access.push_back(a);
    };
    s.kind = kind;
    switch (kind) {
        case RECS_SYS_NBODY_GRAVITY:  // (Read<Position>, Write<Acceleration>, Query<(Read<Position>, Read<Mass>)>)
            s.name = "gravity";
            s.n_comp = 3;
            acc(ids[0], false);
            acc(ids[1], true);
            acc(ids[0], false);
            acc(ids[2], false);
            break;
        case RECS_SYS_NBODY_MOVEMENT:  // (Write<Position>, Read<Velocity>)
        case RECS_SYS_RAIN_MOVEMENT:
            s.name = "movement";
            s.n_comp = 2;
            acc(ids[0], true);
            acc(ids[1], false);
            break;
        case RECS_SYS_NBODY_ACCELERATION:  // (Write<Velocity>, Read<Acceleration>)
            s.name = "acceleration";
            s.n_comp = 2;
            acc(ids[0], true);
            acc(ids[1], false);
            break;
        case RECS_SYS_RAIN_GRAVITY:  // (Write<Velocity>)
            s.name = "gravity";
            s.n_comp = 1;
            acc(ids[0], true);
            break;
        case RECS_SYS_RAIN_WIND:
            s.name = "wind";
            s.n_comp = 1;
            acc(ids[0], true);
            break;
        case RECS_SYS_RAIN_WIND_DIRECTION:
            s.name = "wind_direction";
            s.n_comp = 0;
            break;
        case RECS_SYS_QUERY_ADD:
            s.name = "query_add";
            s.n_comp = 2;
            acc(ids[0], true);
            acc(ids[1], false);
            break;
        case RECS_SYS_NBODY_COLLISION:  // (Entity, Read<Position>, Query<(Entity, Read<Position>)>, Commands)
            s.name = "collision";
            s.n_comp = 1;
            acc(ids[0], false);
            acc(ids[0], false);
            break;
        case RECS_SYS_RAIN_GROUND_COLLISION:  // (Entity, Read<Position>, Commands)
            s.name = "ground_collision";
            s.n_comp = 1;
            acc(ids[0], false);
            break;
        case RECS_SYS_KERNEL:  // described by recs_gpu_system_create_kernel itself
            break;
    }
    for (uint32_t i = 0; i < s.n_comp; ++i) s.comp[i] = ids[i];
}

uint32_t kind_components(int kind) {
    switch (kind) {
        case RECS_SYS_NBODY_GRAVITY:
            return 3;
        case RECS_SYS_NBODY_MOVEMENT:
        case RECS_SYS_NBODY_ACCELERATION:
        case RECS_SYS_RAIN_MOVEMENT:
        case RECS_SYS_QUERY_ADD:
            return 2;
        case RECS_SYS_RAIN_GRAVITY:
        case RECS_SYS_RAIN_WIND:
        case RECS_SYS_NBODY_COLLISION:
        case RECS_SYS_RAIN_GROUND_COLLISION:
            return 1;
        case RECS_SYS_RAIN_WIND_DIRECTION:
            return 0;
    }
    return ~0u;
}

}  // namespace

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" {

int32_t recs_gpu_abi_version(void) { return RECS_GPU_ABI_VERSION; }

void recs_gpu_default_config(recs_gpu_config* cfg) {
    if (!cfg) return;
    memset(cfg, 0, sizeof *cfg);
    cfg->device = 0;
    cfg->rank = 0;
    cfg->world_size = 1;
    cfg->use_cuda_graph = 1;
    cfg->nbody_dt = 1.0f / 20000.0f;     // crates/n-body/src/lib.rs:14
    cfg->nbody_g = 6.67e-11f;            // :123
    cfg->nbody_epsilon = 1.1920929e-7f;  // f32::EPSILON, :118
    cfg->rain_dt = 1.0f / 20.0f;         // crates/rain-simulation/src/lib.rs:15
    cfg->rain_gravity = 9.82f;           // :59
    cfg->rain_terminal = 9.0f;           // :60
    cfg->rain_wind_accel = 10.0f;        // :72
    cfg->rain_wind_deg_per_s = 10.0f;    // :76
    cfg->rain_ground_height = 0.0f;      // :150
    cfg->nbody_collision_radius = 0.1f;  // crates/n-body/examples/n_body_demo.rs:16
}

int32_t recs_gpu_create(const recs_gpu_config* cfg, recs_gpu_ctx** out_ctx) {
    if (!cfg || !out_ctx) return RECS_ERR_INVALID_ARGUMENT;
    *out_ctx = nullptr;
    if (cfg->world_size < 1 || cfg->rank < 0 || cfg->rank >= cfg->world_size) return RECS_ERR_INVALID_ARGUMENT;
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev <= 0) return RECS_ERR_NO_DEVICE;
    if (cfg->device < 0 || cfg->device >= n_dev) return RECS_ERR_INVALID_ARGUMENT;
    if (cudaSetDevice(cfg->device) != cudaSuccess) return RECS_ERR_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess) return RECS_ERR_CUDA;
    if (prop.major != 10) {
        fprintf(stderr, "recs_gpu: device %d is sm_%d%d; this library is built for sm_100a only\n", cfg->device,
                prop.major, prop.minor);
        return RECS_ERR_NO_DEVICE;
    }
    recs_gpu_ctx* c = new recs_gpu_ctx();
    c->cfg = *cfg;
    c->sm_count = prop.multiProcessorCount;
    bool ok = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&c->comm_stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaEventCreateWithFlags(&c->ev_pack, cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&c->ev_gather, cudaEventDisableTiming) == cudaSuccess &&
              gravity_prepare() == cudaSuccess;
    if (ok) {
        const char* v = getenv("RECS_GPU_GRAVITY_VARIANT");  // tuning knob; default chosen by measurement
        if (v && v[0] && atoi(v) >= 0 && atoi(v) < kGravVariantCount) {
            c->grav_variant = atoi(v);
            c->grav_variant_forced = true;
        }
        for (int k = 0; k < kGravVariantCount && ok; ++k) {
            c->grav_ctas_per_sm[k] = gravity_max_ctas_per_sm(k);
            ok = c->grav_ctas_per_sm[k] > 0;
        }
        // test knobs: neither may change a bit of the result (tests/test_gpu_nbody.py)
        if (const char* g = getenv("RECS_GPU_GRAVITY_GRID")) c->grav_grid_forced = atoi(g);
        c->grav_no_heavy = getenv("RECS_GPU_GRAVITY_NO_LAST_SPLIT") != nullptr;
        if (const char* r = getenv("RECS_GPU_RESIDENT_MAX_ROWS")) c->resident_max_rows = (uint64_t)atoll(r);
        if (const char* t = getenv("RECS_GPU_GRAPH_TICKS")) c->graph_ticks_wanted = (uint32_t)std::max(1, std::min(64, atoi(t)));
        if (c->resident_max_rows > (uint64_t)kResidentMaxTiles * kGravTile) c->resident_max_rows = (uint64_t)kResidentMaxTiles * kGravTile;
        if (const char* e = getenv("RECS_GPU_GRAVITY_EMULATE_RANK")) {
            int R = 0, r = 0;
            if (sscanf(e, "%d,%d", &R, &r) == 2 && R > 1 && R <= kGravMaxPeers && r >= 0 && r < R) c->emulate_ranks = R, c->emulate_rank = r;
        }
        const char* x = getenv("RECS_GPU_EXCHANGE");  // "nccl": keep the all-gather exchange even when peers are mapped
        c->exchange_nccl_forced = x && strcmp(x, "nccl") == 0;
        ok = ok && cudaHostAlloc((void**)&c->dev_error_host, 64, cudaHostAllocMapped) == cudaSuccess &&
             cudaHostGetDevicePointer((void**)&c->dev_error, c->dev_error_host, 0) == cudaSuccess;
        if (ok) *c->dev_error_host = 0;
    }
    if (!ok) {
        fprintf(stderr, "recs_gpu: context creation failed: %s\n", cudaGetErrorString(cudaGetLastError()));
        delete c;
        return RECS_ERR_CUDA;
    }
    *out_ctx = c;
    return RECS_OK;
}

int32_t recs_gpu_destroy(recs_gpu_ctx* c) {
    if (!c) return RECS_ERR_INVALID_ARGUMENT;
    {
        std::lock_guard<std::mutex> lock(c->mu);
        cudaSetDevice(c->cfg.device);
        cudaStreamSynchronize(c->stream);
        cudaStreamSynchronize(c->comm_stream);
        if (c->comm) c->nccl.CommDestroy(c->comm);
        drop_graph(c);
        for (auto& kv : c->columns)
            if (kv.second.dev) cudaFree(kv.second.dev);
        for (auto& sys : c->systems)
            for (auto& kv : sys.removals) {
                if (kv.second.flags) cudaFree(kv.second.flags);
                if (kv.second.rows) cudaFree(kv.second.rows);
                if (kv.second.counter) cudaFree(kv.second.counter);
            }
        if (c->posm) cudaFree(c->posm);
        if (c->posm_alt) cudaFree(c->posm_alt);
        if (c->resident_barrier) cudaFree(c->resident_barrier);
        if (c->frags) cudaFree(c->frags);
        for (int r = 0; r < kGravMaxPeers; ++r)
            if (c->xchg.peer[r] && c->xchg.peer[r] != c->xchg.base) cudaIpcCloseMemHandle(c->xchg.peer[r]);
        if (c->xchg.base) cudaFree(c->xchg.base);
        if (c->dev_error_host) cudaFreeHost(c->dev_error_host);
        if (c->l2_flush) cudaFree(c->l2_flush);
        if (c->move_idx) cudaFree(c->move_idx);
        if (c->move_tmp) cudaFree(c->move_tmp);
        if (c->stage) cudaFreeHost(c->stage);
        for (int h = 0; h < 2; ++h)
            if (c->stage_ev[h]) cudaEventDestroy(c->stage_ev[h]);
        for (auto& s : c->spans) {
            cudaEventDestroy(s.a);
            cudaEventDestroy(s.b);
        }
        for (auto e : c->event_pool) cudaEventDestroy(e);
        if (c->timer_a) cudaEventDestroy(c->timer_a);
        if (c->timer_b) cudaEventDestroy(c->timer_b);
        cudaEventDestroy(c->ev_pack);
        cudaEventDestroy(c->ev_gather);
        cudaStreamDestroy(c->stream);
        cudaStreamDestroy(c->comm_stream);
    }
    delete c;
    return RECS_OK;
}

const char* recs_gpu_last_error(recs_gpu_ctx* c) {
    if (!c) return "null context";
    std::lock_guard<std::mutex> lock(c->mu);
    return c->error.c_str();
}

int32_t recs_gpu_clear_error(recs_gpu_ctx* c) {
    if (!c) return RECS_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lock(c->mu);
    c->status = RECS_OK;
    c->error.clear();
    cudaGetLastError();
    return RECS_OK;
}

int32_t recs_gpu_synchronize(recs_gpu_ctx* c) {
    RECS_ENTER(c);
    RECS_CUDA(c, cudaStreamSynchronize(c->stream));
    RECS_CUDA(c, cudaStreamSynchronize(c->comm_stream));
    return check_device_error(c);
}

int32_t recs_gpu_bind_column(recs_gpu_ctx* c, uint32_t arch, uint64_t comp, void* host_ptr, uint32_t elem,
                             uint64_t count) {
    RECS_ENTER(c);
    if (elem == 0 || (count > 0 && !host_ptr)) return fail(c, RECS_ERR_INVALID_ARGUMENT, "bind_column: null host pointer or zero element size");
    Column& col = c->columns[ColKey(arch, comp)];
    const size_t bytes = std::max<size_t>((size_t)elem * count, 16);
    const bool layout_changed = col.cap < bytes || !col.dev || col.elem != elem || col.count != count || col.stale;
    if (col.cap < bytes || !col.dev) {
        // wait for in-flight work before releasing the old buffer
        RECS_CUDA(c, cudaStreamSynchronize(c->stream));
        if (col.dev) RECS_CUDA(c, cudaFree(col.dev));
        col.dev = nullptr;
        col.cap = 0;
        // round capacity up so growing archetypes do not reallocate every tick
        size_t cap = 256;
        while (cap < bytes) cap <<= 1;
        if (cap > bytes + (64u << 20)) cap = (bytes + (64u << 20) - 1) / (64u << 20) * (64u << 20);
        RECS_CUDA(c, cudaMalloc(&col.dev, cap));
        col.cap = cap;
    }
    col.host = host_ptr;
    col.elem = elem;
    col.count = count;
    col.stale = false;
    col.version++;
    if (layout_changed) c->epoch++;  // the layout epoch: device pointers, row counts, archetype lists (tables, graph)
    return RECS_OK;
}

int32_t recs_gpu_upload(recs_gpu_ctx* c, uint32_t arch, uint64_t comp) {
    RECS_ENTER(c);
    Column* col = find_column(c, arch, comp);
    if (!col || !col->dev) return fail(c, RECS_ERR_MISSING_PARAMETER, "upload: column is not bound");
    if (col->stale) return fail(c, RECS_ERR_STALE_BINDING, "upload: column binding is stale");
    if (col->count)
        RECS_CUDA(c, cudaMemcpyAsync(col->dev, col->host, (size_t)col->elem * col->count, cudaMemcpyHostToDevice,
                                     c->stream));
    c->h2d_bytes += (uint64_t)col->elem * col->count;
    col->version++;  // (a captured tick that elided the mirror re-pack is dropped by the version check in recs_gpu_tick)
    return RECS_OK;
}

// Rows [row_begin, row_end) only: what a rank of an index-sharded run owns (recs_gpu_shard_rows).
int32_t recs_gpu_upload_rows(recs_gpu_ctx* c, uint32_t arch, uint64_t comp, uint64_t row_begin, uint64_t row_end) {
    RECS_ENTER(c);
    Column* col = find_column(c, arch, comp);
    if (!col || !col->dev) return fail(c, RECS_ERR_MISSING_PARAMETER, "upload_rows: column is not bound");
    if (col->stale) return fail(c, RECS_ERR_STALE_BINDING, "upload_rows: column binding is stale");
    if (row_begin > row_end || row_end > col->count) return fail(c, RECS_ERR_INVALID_ARGUMENT, "upload_rows: row range outside the column");
    const size_t off = (size_t)col->elem * row_begin, bytes = (size_t)col->elem * (row_end - row_begin);
    if (bytes)
        RECS_CUDA(c, cudaMemcpyAsync((unsigned char*)col->dev + off, (const unsigned char*)col->host + off, bytes,
                                     cudaMemcpyHostToDevice, c->stream));
    c->h2d_bytes += bytes;
    col->version++;
    return RECS_OK;
}

int32_t recs_gpu_download_rows(recs_gpu_ctx* c, uint32_t arch, uint64_t comp, uint64_t row_begin, uint64_t row_end) {
    RECS_ENTER(c);
    Column* col = find_column(c, arch, comp);
    if (!col || !col->dev) return fail(c, RECS_ERR_MISSING_PARAMETER, "download_rows: column is not bound");
    if (col->stale) return fail(c, RECS_ERR_STALE_BINDING, "download_rows: column binding is stale");
    if (row_begin > row_end || row_end > col->count) return fail(c, RECS_ERR_INVALID_ARGUMENT, "download_rows: row range outside the column");
    const size_t off = (size_t)col->elem * row_begin, bytes = (size_t)col->elem * (row_end - row_begin);
    if (bytes)
        RECS_CUDA(c, cudaMemcpyAsync((unsigned char*)col->host + off, (const unsigned char*)col->dev + off, bytes,
                                     cudaMemcpyDeviceToHost, c->stream));
    c->d2h_bytes += bytes;
    RECS_CUDA(c, cudaStreamSynchronize(c->stream));
    return check_device_error(c);
}

int32_t recs_gpu_download(recs_gpu_ctx* c, uint32_t arch, uint64_t comp) {
    RECS_ENTER(c);
    Column* col = find_column(c, arch, comp);
    if (!col || !col->dev) return fail(c, RECS_ERR_MISSING_PARAMETER, "download: column is not bound");
    if (col->stale) return fail(c, RECS_ERR_STALE_BINDING, "download: column binding is stale");
    if (col->count)
        RECS_CUDA(c, cudaMemcpyAsync(col->host, col->dev, (size_t)col->elem * col->count, cudaMemcpyDeviceToHost,
                                     c->stream));
    c->d2h_bytes += (uint64_t)col->elem * col->count;
    RECS_CUDA(c, cudaStreamSynchronize(c->stream));
    return check_device_error(c);
}

// ---- structural changes without a round trip ------------------------------------------------------------------------
int32_t recs_gpu_archetype_move_rows(recs_gpu_ctx* c, uint32_t arch, const uint64_t* comps, uint32_t n_comps,
                                     const uint32_t* dst, const uint32_t* src, uint32_t n_moves, uint64_t new_count) {
    RECS_ENTER(c);
    if ((n_moves && (!dst || !src)) || (n_comps && !comps))
        return fail(c, RECS_ERR_INVALID_ARGUMENT, "archetype_move_rows: null list");
    std::vector<Column*> cols;
    uint32_t max_elem = 0;
    for (uint32_t i = 0; i < n_comps; ++i) {
        Column* col = find_column(c, arch, comps[i]);
        if (!col || !col->dev) return fail(c, RECS_ERR_MISSING_PARAMETER, "archetype_move_rows: column is not bound");
        if (col->stale) return fail(c, RECS_ERR_STALE_BINDING, "archetype_move_rows: column binding is stale");
        if (new_count > col->count)
            return fail(c, RECS_ERR_SIZE_MISMATCH, "archetype_move_rows: new row count exceeds the bound column (append first)");
        cols.push_back(col);
        max_elem = std::max(max_elem, col->elem);
    }
    if (cols.empty()) return RECS_OK;
    if (n_moves) {
        int32_t st = ensure(c, &c->move_idx, &c->move_idx_cap, (size_t)n_moves * 2 * sizeof(uint32_t));
        if (st != RECS_OK) return st;
        if ((st = ensure(c, &c->move_tmp, &c->move_tmp_cap, (size_t)n_moves * max_elem)) != RECS_OK) return st;
        if ((st = staged_h2d(c, c->move_idx, dst, (size_t)n_moves * 4)) != RECS_OK) return st;
        if ((st = staged_h2d(c, c->move_idx + n_moves, src, (size_t)n_moves * 4)) != RECS_OK) return st;
        c->h2d_bytes += (uint64_t)n_moves * 8;
    }
    for (Column* col : cols) {
        if (n_moves) {
            RECS_CUDA(c, launch_rows_move(col->dev, col->elem, c->move_idx, c->move_idx + n_moves, n_moves, c->move_tmp, c->stream));
            c->kernel_launches += 2;
        }
        col->count = new_count;
        col->version++;
    }
    c->epoch++;
    return RECS_OK;
}

int32_t recs_gpu_column_append(recs_gpu_ctx* c, uint32_t arch, uint64_t comp, void* host_ptr, uint64_t new_count) {
    RECS_ENTER(c);
    Column* col = find_column(c, arch, comp);
    if (!col || !col->dev) return fail(c, RECS_ERR_MISSING_PARAMETER, "column_append: column is not bound");
    if (col->stale) return fail(c, RECS_ERR_STALE_BINDING, "column_append: column binding is stale");
    if (new_count < col->count || (new_count && !host_ptr))
        return fail(c, RECS_ERR_INVALID_ARGUMENT, "column_append: rows can only be added (and need a host pointer)");
    const size_t old_bytes = (size_t)col->elem * col->count, new_bytes = (size_t)col->elem * new_count;
    if (new_bytes > col->cap) {
        size_t cap = 256;
        while (cap < new_bytes) cap <<= 1;
        if (cap > new_bytes + (64u << 20)) cap = (new_bytes + (64u << 20) - 1) / (64u << 20) * (64u << 20);
        void* grown = nullptr;
        RECS_CUDA(c, cudaMalloc(&grown, cap));
        if (old_bytes) RECS_CUDA(c, cudaMemcpyAsync(grown, col->dev, old_bytes, cudaMemcpyDeviceToDevice, c->stream));
        RECS_CUDA(c, cudaStreamSynchronize(c->stream));
        RECS_CUDA(c, cudaFree(col->dev));
        col->dev = grown;
        col->cap = cap;
    }
    if (new_bytes > old_bytes) {
        // (the host rows may be pageable and short-lived: staged, or waited for when large)
        int32_t st = staged_h2d(c, (unsigned char*)col->dev + old_bytes, (const unsigned char*)host_ptr + old_bytes,
                                new_bytes - old_bytes);
        if (st != RECS_OK) return st;
        c->h2d_bytes += new_bytes - old_bytes;
    }
    col->host = host_ptr;
    col->count = new_count;
    col->version++;
    c->epoch++;
    return RECS_OK;
}

int32_t recs_gpu_column_rebind_host(recs_gpu_ctx* c, uint32_t arch, uint64_t comp, void* host_ptr) {
    RECS_ENTER(c);
    Column* col = find_column(c, arch, comp);
    if (!col || !col->dev) return fail(c, RECS_ERR_MISSING_PARAMETER, "column_rebind_host: column is not bound");
    col->host = host_ptr;
    return RECS_OK;
}

int32_t recs_gpu_get_transfer_bytes(recs_gpu_ctx* c, uint64_t* out_h2d, uint64_t* out_d2h) {
    RECS_ENTER(c);
    if (out_h2d) *out_h2d = c->h2d_bytes;
    if (out_d2h) *out_d2h = c->d2h_bytes;
    return RECS_OK;
}

int32_t recs_gpu_invalidate(recs_gpu_ctx* c, uint32_t arch) {
    RECS_ENTER(c);
    for (auto& kv : c->columns)
        if (kv.first.first == arch) kv.second.stale = true;
    c->epoch++;
    return RECS_OK;
}

int32_t recs_gpu_column_device_ptr(recs_gpu_ctx* c, uint32_t arch, uint64_t comp, void** out_ptr, uint64_t* out_count) {
    RECS_ENTER(c);
    Column* col = find_column(c, arch, comp);
    if (!col || !col->dev || !out_ptr) return fail(c, RECS_ERR_MISSING_PARAMETER, "column_device_ptr: column is not bound");
    *out_ptr = col->dev;
    if (out_count) *out_count = col->count;
    return RECS_OK;
}

int32_t recs_gpu_system_create(recs_gpu_ctx* c, recs_gpu_system_kind kind, const uint64_t* ids, uint32_t n_ids,
                               uint32_t* out_id) {
    RECS_ENTER(c);
    const uint32_t need = kind_components((int)kind);
    if (need == ~0u) return fail(c, RECS_ERR_INVALID_ARGUMENT, "system_create: unknown system kind");
    if (n_ids != need || (need && !ids) || !out_id)
        return fail(c, RECS_ERR_INVALID_ARGUMENT, "system_create: wrong number of component ids for this kind");
    System s;
    uint64_t none[3] = {0, 0, 0};
    fill_system(s, kind, need ? ids : none);
    c->systems.push_back(s);
    *out_id = (uint32_t)c->systems.size() - 1;
    c->epoch++;
    return RECS_OK;
}

int32_t recs_gpu_system_create_kernel(recs_gpu_ctx* c, const char* name, const char* source, const char* body,
                                      const recs_gpu_kernel_param* params, uint32_t n_params, const char* uniform_type,
                                      uint32_t uniform_bytes, uint32_t* out_id) {
    RECS_ENTER(c);
    if (!name || !source || !body || !out_id || (n_params && !params) || n_params == 0 ||
        n_params > RECS_GPU_MAX_ACCESSES)
        return fail(c, RECS_ERR_INVALID_ARGUMENT, "system_create_kernel: bad argument");
    if ((uniform_type != nullptr && uniform_type[0]) != (uniform_bytes != 0) || uniform_bytes > 256)
        return fail(c, RECS_ERR_INVALID_ARGUMENT, "system_create_kernel: uniform type / size (<= 256 bytes) mismatch");
    auto ks = std::make_shared<KernelSys>();
    System s;
    s.kind = RECS_SYS_KERNEL;
    s.name = name;
    s.n_comp = 0;
    for (uint32_t i = 0; i < n_params; ++i) {
        const recs_gpu_kernel_param& p = params[i];
        if (p.kind > RECS_KPARAM_COMMANDS) return fail(c, RECS_ERR_INVALID_ARGUMENT, "system_create_kernel: bad parameter kind");
        if (p.kind == RECS_KPARAM_COMMANDS) {  // no column: one removal flag per row, owned by the context
            if (ks->has_commands) return fail(c, RECS_ERR_INVALID_ARGUMENT, "system_create_kernel: more than one Commands parameter");
            if (c->cfg.world_size > 1) return fail(c, RECS_ERR_UNSUPPORTED, "removal-deciding systems are single-GPU");
            ks->has_commands = true;
            ks->params.push_back(commands_param_spec(p));
            ks->create_bytes = ks->params.back().elem_size;
            ks->comp.push_back(0);
            continue;
        }
        if (!p.type_name || !p.type_name[0] || p.elem_size == 0)
            return fail(c, RECS_ERR_INVALID_ARGUMENT, "system_create_kernel: bad parameter description");
        if (p.kind == RECS_KPARAM_ENTITY && p.elem_size != 8)
            return fail(c, RECS_ERR_INVALID_ARGUMENT, "system_create_kernel: Entity parameters are 8 bytes (generation << 32 | id)");
        ks->params.push_back({p.type_name, p.elem_size, p.kind});
        const uint64_t comp = p.kind == RECS_KPARAM_ENTITY ? RECS_GPU_ENTITY_COMPONENT : p.component_id;
        ks->comp.push_back(comp);
        if (p.kind != RECS_KPARAM_ENTITY) {  // `Entity` is no component access (crates/ecs/src/systems/mod.rs:893-895)
            recs_gpu_access a;
            a.component_id = comp;
            a.is_write = p.kind == RECS_KPARAM_WRITE ? 1 : 0;
            s.access.push_back(a);
        }
    }
    static const bool force_direct = getenv("RECS_GPU_KERNEL_DIRECT") != nullptr;  // A/B knob
    ks->plan = plan_kernel_system(source, body, ks->params, uniform_type ? uniform_type : "", force_direct ? 1 : 0);
    std::vector<char> cubin;
    std::string log;
    if (!compile_kernel_system(ks->plan.source, &cubin, &log))
        return fail(c, RECS_ERR_COMPILE, std::string("system '") + name + "' did not compile:\n" + log);
    RECS_CUDA(c, cudaLibraryLoadData(&ks->lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0));
    RECS_CUDA(c, cudaLibraryGetKernel(&ks->kernel, ks->lib, "recs_generic_system"));
    if (ks->plan.smem_bytes > 32u * 1024u)  // dynamic + the static barrier words must stay under the 48 KiB default
        RECS_CUDA(c, cudaFuncSetAttribute((const void*)ks->kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)ks->plan.smem_bytes));
    ks->has_uniform = uniform_bytes != 0;
    ks->uniform.assign(std::max<uint32_t>(uniform_bytes, 1), 0);
    s.ks = ks;
    c->systems.push_back(s);
    *out_id = (uint32_t)c->systems.size() - 1;
    c->epoch++;
    return RECS_OK;
}

int32_t recs_gpu_system_create_pair_kernel(recs_gpu_ctx* c, const char* name, const char* source, const char* acc_type,
                                           const char* pair_fn, const char* finish_fn, const recs_gpu_kernel_param* params,
                                           uint32_t n_params, const char* uniform_type, uint32_t uniform_bytes,
                                           uint32_t* out_id) {
    RECS_ENTER(c);
    if (!name || !source || !acc_type || !pair_fn || !finish_fn || !out_id || !params || n_params == 0 ||
        n_params > 2 * RECS_GPU_MAX_ACCESSES)
        return fail(c, RECS_ERR_INVALID_ARGUMENT, "system_create_pair_kernel: bad argument");
    if ((uniform_type != nullptr && uniform_type[0]) != (uniform_bytes != 0) || uniform_bytes > 256)
        return fail(c, RECS_ERR_INVALID_ARGUMENT, "system_create_pair_kernel: uniform type / size (<= 256 bytes) mismatch");
    // Sharded contexts: every rank folds the WHOLE nested query for the outer rows it owns (run_kernel_system shards the
    // outer archetypes), so the query's columns must be complete on every rank -- recs_gpu_column_allgather after a
    // sharded system wrote one of them.
    auto ks = std::make_shared<KernelSys>();
    ks->is_pair = true;
    System s;
    s.kind = RECS_SYS_KERNEL;
    s.name = name;
    s.n_comp = 0;
    std::vector<recs_gpu_access> query_access;
    for (uint32_t i = 0; i < n_params; ++i) {
        const recs_gpu_kernel_param& p = params[i];
        if (p.kind > RECS_KPARAM_QUERY_ENTITY) return fail(c, RECS_ERR_INVALID_ARGUMENT, "system_create_pair_kernel: bad parameter kind");
        if (p.kind == RECS_KPARAM_COMMANDS) {
            if (ks->has_commands) return fail(c, RECS_ERR_INVALID_ARGUMENT, "system_create_pair_kernel: more than one Commands parameter");
            ks->has_commands = true;
            ks->params.push_back(commands_param_spec(p));
            ks->create_bytes = ks->params.back().elem_size;
            ks->comp.push_back(0);
            continue;
        }
        if (!p.type_name || !p.type_name[0] || p.elem_size == 0)
            return fail(c, RECS_ERR_INVALID_ARGUMENT, "system_create_pair_kernel: bad parameter description");
        const bool entity = p.kind == RECS_KPARAM_ENTITY || p.kind == RECS_KPARAM_QUERY_ENTITY;
        if (entity && p.elem_size != 8)
            return fail(c, RECS_ERR_INVALID_ARGUMENT, "system_create_pair_kernel: Entity parameters are 8 bytes (generation << 32 | id)");
        const uint64_t comp = entity ? RECS_GPU_ENTITY_COMPONENT : p.component_id;
        const bool nested = p.kind == RECS_KPARAM_QUERY_READ || p.kind == RECS_KPARAM_QUERY_ENTITY;
        if (nested) {
            ks->qparams.push_back({p.type_name, p.elem_size, p.kind});
            ks->qcomp.push_back(comp);
        } else {
            ks->params.push_back({p.type_name, p.elem_size, p.kind});
            ks->comp.push_back(comp);
        }
        if (!entity) {  // accesses in parameter order, the nested query's after the outer ones (systems/mod.rs:973-978)
            recs_gpu_access a;
            a.component_id = comp;
            a.is_write = p.kind == RECS_KPARAM_WRITE ? 1 : 0;
            (nested ? query_access : s.access).push_back(a);
        }
    }
    if (ks->qparams.empty()) return fail(c, RECS_ERR_INVALID_ARGUMENT, "system_create_pair_kernel: no nested query parameter");
    bool outer_column = false;
    for (const auto& p : ks->params) outer_column = outer_column || p.kind != RECS_KPARAM_COMMANDS;
    if (!outer_column) return fail(c, RECS_ERR_UNSUPPORTED, "a generic system needs at least one outer Read / Write / Entity parameter");
    s.access.insert(s.access.end(), query_access.begin(), query_access.end());
    if (s.access.size() > RECS_GPU_MAX_ACCESSES) return fail(c, RECS_ERR_INVALID_ARGUMENT, "system_create_pair_kernel: too many component accesses");
    ks->plan = plan_pair_kernel_system(source, acc_type, pair_fn, finish_fn, ks->params, ks->qparams, uniform_type ? uniform_type : "");
    std::vector<char> cubin;
    std::string log;
    if (!compile_kernel_system(ks->plan.source, &cubin, &log))
        return fail(c, RECS_ERR_COMPILE, std::string("system '") + name + "' did not compile:\n" + log);
    RECS_CUDA(c, cudaLibraryLoadData(&ks->lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0));
    RECS_CUDA(c, cudaLibraryGetKernel(&ks->kernel, ks->lib, "recs_generic_system"));
    ks->has_uniform = uniform_bytes != 0;
    ks->uniform.assign(std::max<uint32_t>(uniform_bytes, 1), 0);
    s.ks = ks;
    c->systems.push_back(s);
    *out_id = (uint32_t)c->systems.size() - 1;
    c->epoch++;
    return RECS_OK;
}

int32_t recs_gpu_system_set_uniform(recs_gpu_ctx* c, uint32_t id, const void* data, uint32_t bytes) {
    RECS_ENTER(c);
    if (id >= c->systems.size() || !c->systems[id].ks || !data)
        return fail(c, RECS_ERR_INVALID_ARGUMENT, "system_set_uniform: not a generic system");
    KernelSys& k = *c->systems[id].ks;
    if (!k.has_uniform || bytes != k.uniform.size())
        return fail(c, RECS_ERR_SIZE_MISMATCH, "system_set_uniform: size differs from the registered uniform type");
    if (memcmp(k.uniform.data(), data, bytes) != 0) {
        memcpy(k.uniform.data(), data, bytes);
        drop_graph(c);  // a captured tick holds the old value
    }
    return RECS_OK;
}

int32_t recs_gpu_system_describe(recs_gpu_ctx* c, uint32_t id, recs_gpu_system_desc* out) {
    RECS_ENTER(c);
    if (id >= c->systems.size() || !out) return fail(c, RECS_ERR_INVALID_ARGUMENT, "system_describe: unknown system id");
    const System& s = c->systems[id];
    out->name = s.name.c_str();
    out->n_access = (uint32_t)s.access.size();
    for (size_t i = 0; i < s.access.size(); ++i) out->access[i] = s.access[i];
    return RECS_OK;
}

int32_t recs_gpu_system_set_archetypes(recs_gpu_ctx* c, uint32_t id, const uint32_t* archs, uint32_t n) {
    RECS_ENTER(c);
    if (id >= c->systems.size() || (n && !archs)) return fail(c, RECS_ERR_INVALID_ARGUMENT, "system_set_archetypes: bad argument");
    std::vector<uint32_t> next(archs, archs + n);
    if (next != c->systems[id].archs) {  // unchanged membership keeps the captured tick graph valid
        c->systems[id].archs.swap(next);
        c->epoch++;
    }
    return RECS_OK;
}

int32_t recs_gpu_system_set_query_archetypes(recs_gpu_ctx* c, uint32_t id, const uint32_t* archs, uint32_t n) {
    RECS_ENTER(c);
    if (id >= c->systems.size() || (n && !archs)) return fail(c, RECS_ERR_INVALID_ARGUMENT, "system_set_query_archetypes: bad argument");
    std::vector<uint32_t> next(archs, archs + n);
    if (next != c->systems[id].qarchs) {
        c->systems[id].qarchs.swap(next);
        c->epoch++;
    }
    return RECS_OK;
}

int32_t recs_gpu_system_removals(recs_gpu_ctx* c, uint32_t id, uint32_t arch, uint32_t* out_rows, uint32_t cap,
                                 uint32_t* out_count) {
    RECS_ENTER(c);
    if (id >= c->systems.size() || !out_count) return fail(c, RECS_ERR_INVALID_ARGUMENT, "system_removals: bad argument");
    System& s = c->systems[id];
    auto it = s.removals.find(arch);
    if (it == s.removals.end() || !it->second.valid)
        return fail(c, RECS_ERR_INVALID_ARGUMENT, "system_removals: the system has not run over this archetype");
    System::Removals& r = it->second;
    uint32_t n = 0;
    RECS_CUDA(c, cudaMemcpyAsync(&n, r.counter, sizeof n, cudaMemcpyDeviceToHost, c->stream));
    RECS_CUDA(c, cudaStreamSynchronize(c->stream));
    *out_count = n;
    if (out_rows && n) {
        std::vector<uint32_t> rows(n);
        RECS_CUDA(c, cudaMemcpy(rows.data(), r.rows, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
        std::sort(rows.begin(), rows.end());  // playback order: ascending component index
        for (uint32_t i = 0; i < n && i < cap; ++i) out_rows[i] = rows[i];
    }
    return RECS_OK;
}

int32_t recs_gpu_system_set_create_capacity(recs_gpu_ctx* c, uint32_t id, uint32_t max_records) {
    RECS_ENTER(c);
    if (id >= c->systems.size() || !c->systems[id].ks || !c->systems[id].ks->create_bytes || max_records == 0)
        return fail(c, RECS_ERR_INVALID_ARGUMENT, "system_set_create_capacity: not a generic system with a create record");
    KernelSys& k = *c->systems[id].ks;
    if (k.create_cap != max_records) {
        k.create_cap = max_records;
        k.table_epoch = ~0ull;  // the buffer is (re)allocated with the table
        k.create_valid = false;
        drop_graph(c);
    }
    return RECS_OK;
}

int32_t recs_gpu_system_creates(recs_gpu_ctx* c, uint32_t id, void* out_records, uint32_t cap_records, uint32_t* out_count,
                                uint32_t* out_record_bytes) {
    RECS_ENTER(c);
    if (id >= c->systems.size() || !c->systems[id].ks || !out_count)
        return fail(c, RECS_ERR_INVALID_ARGUMENT, "system_creates: not a generic system");
    KernelSys& k = *c->systems[id].ks;
    if (out_record_bytes) *out_record_bytes = k.create_bytes;
    *out_count = 0;
    if (!k.create_bytes || !k.create_valid || !k.create_buf) return RECS_OK;  // nothing recorded (or never run)
    uint32_t header[4] = {0, 0, 0, 0};
    RECS_CUDA(c, cudaMemcpyAsync(header, k.create_buf, sizeof header, cudaMemcpyDeviceToHost, c->stream));
    RECS_CUDA(c, cudaStreamSynchronize(c->stream));
    if (header[3]) return fail(c, RECS_ERR_SIZE_MISMATCH, "Commands::create was called with a record whose size differs from the registered one");
    const uint32_t n = header[0];
    if (n > k.create_cap) {
        char buf[200];
        snprintf(buf, sizeof buf, "Commands::create: %u records in one run, the buffer holds %u (recs_gpu_system_set_create_capacity)", n, k.create_cap);
        return fail(c, RECS_ERR_SIZE_MISMATCH, buf);
    }
    *out_count = n;
    if (!out_records || n == 0) return RECS_OK;
    const uint32_t stride = k.create_stride();
    std::vector<unsigned char> raw((size_t)n * stride);
    RECS_CUDA(c, cudaMemcpy(raw.data(), k.create_buf + 16, raw.size(), cudaMemcpyDeviceToHost));
    c->d2h_bytes += raw.size();
    // playback order = the order a sequential run would have recorded them in: entity order, then call order
    struct Key {
        unsigned long long entity;
        uint32_t call, slot;
        bool operator<(const Key& o) const { return entity != o.entity ? entity < o.entity : call < o.call; }
    };
    std::vector<Key> order(n);
    for (uint32_t i = 0; i < n; ++i) {
        memcpy(&order[i].entity, raw.data() + (size_t)i * stride, 8);
        memcpy(&order[i].call, raw.data() + (size_t)i * stride + 8, 4);
        order[i].slot = i;
    }
    if (!std::is_sorted(order.begin(), order.end())) std::sort(order.begin(), order.end());
    for (uint32_t i = 0; i < n && i < cap_records; ++i)
        memcpy((unsigned char*)out_records + (size_t)i * k.create_bytes, raw.data() + (size_t)order[i].slot * stride + 16, k.create_bytes);
    return RECS_OK;
}

int32_t recs_gpu_run_system(recs_gpu_ctx* c, uint32_t id, uint32_t flags) {
    RECS_ENTER(c);
    int32_t st = run_system_locked(c, id);
    if (st != RECS_OK) return st;
    if (flags & RECS_GPU_SYNC) RECS_CUDA(c, cudaStreamSynchronize(c->stream));
    return RECS_OK;
}

int32_t recs_gpu_set_tick_order(recs_gpu_ctx* c, const uint32_t* ids, uint32_t n) {
    RECS_ENTER(c);
    for (uint32_t i = 0; i < n; ++i)
        if (!ids || ids[i] >= c->systems.size()) return fail(c, RECS_ERR_INVALID_ARGUMENT, "set_tick_order: unknown system id");
    std::vector<uint32_t> next(ids, ids + n);
    if (next != c->order) {
        c->order.swap(next);
        drop_graph(c);  // the captured chain is another one; archetype tables of generic systems stay valid
    }
    return RECS_OK;
}

// The tick is exactly gravity -> movement -> acceleration over one small body archetype (the reference's benchmark at its
// own size): `n_ticks` further ticks in one cooperative launch.  The mirror must be current (a tick of the ordinary chain
// just ran).  Bit-identical to running the chain tick by tick.
bool resident_ticks_apply(recs_gpu_ctx* c) {
    if (c->order.size() != 3 || c->cfg.world_size != 1 || c->emulate_ranks > 1 || c->grav_variant_forced || c->timing) return false;
    System& g = c->systems[c->order[0]];
    System& m = c->systems[c->order[1]];
    System& a = c->systems[c->order[2]];
    if (g.kind != RECS_SYS_NBODY_GRAVITY || m.kind != RECS_SYS_NBODY_MOVEMENT || a.kind != RECS_SYS_NBODY_ACCELERATION) return false;
    if (g.archs.size() != 1 || g.archs != g.qarchs || g.archs != m.archs || g.archs != a.archs) return false;
    if (m.comp[0] != g.comp[0] || a.comp[0] != m.comp[1] || a.comp[1] != g.comp[1]) return false;
    const Column* p = find_column(c, g.archs[0], g.comp[0]);
    return p && p->count >= 1 && p->count <= c->resident_max_rows && p->count <= kGravSmallRows;
}

// The mirror holds exactly this archetype's bodies, unscaled, and no source column was written since it was refreshed.
bool resident_mirror_ready(recs_gpu_ctx* c) {
    System& g = c->systems[c->order[0]];
    const uint32_t arch = g.archs[0];
    const Column* p = find_column(c, arch, g.comp[0]);
    if (!p || !c->posm || c->posm_scaled) return false;
    const uint64_t n_pad = (p->count + kGravTile - 1) / kGravTile * kGravTile;
    if (c->posm_bodies != n_pad || c->posm_sources.size() != 2) return false;
    if (c->posm_sources[0].first != ColKey(arch, g.comp[0]) || c->posm_sources[1].first != ColKey(arch, g.comp[2])) return false;
    return posm_is_current(c);
}

int32_t run_resident_ticks(recs_gpu_ctx* c, uint32_t n_ticks, bool* launched) {
    *launched = true;
    System& g = c->systems[c->order[0]];
    const uint32_t arch = g.archs[0];
    Column *pos, *vel, *acc, *mass;
    int32_t st;
    if ((st = need_column(c, arch, g.comp[0], 12, &pos)) != RECS_OK) return st;
    if ((st = need_column(c, arch, c->systems[c->order[1]].comp[1], 12, &vel)) != RECS_OK) return st;
    if ((st = need_column(c, arch, g.comp[1], 12, &acc)) != RECS_OK) return st;
    if ((st = need_column(c, arch, g.comp[2], 4, &mass)) != RECS_OK) return st;
    if ((st = same_count(c, pos, vel)) != RECS_OK || (st = same_count(c, pos, acc)) != RECS_OK) return st;
    const uint64_t n = pos->count;
    const uint64_t n_pad = (n + kGravTile - 1) / kGravTile * kGravTile;
    if (c->posm_bodies != n_pad || c->posm_scaled || !posm_is_current(c))
        return fail(c, RECS_ERR_UNSUPPORTED, "resident ticks need the current unscaled mirror of the archetype");
    const float* alt_before = c->posm_alt;
    if ((st = ensure(c, &c->posm_alt, &c->posm_alt_cap, n_pad * 16)) != RECS_OK) return st;
    if (c->posm_alt != alt_before) c->posm_alt_serial = ~0ull;
    if (!c->resident_barrier) {
        RECS_CUDA(c, cudaMalloc((void**)&c->resident_barrier, 256));
        RECS_CUDA(c, cudaMemsetAsync(c->resident_barrier, 0, 256, c->stream));
        c->resident_barrier_count = 0;
    }
    if (c->posm_alt_serial != c->posm_pack_serial) {
        // G*m and the null bodies of the second buffer: the kernel refreshes positions only, so once both buffers carry the
        // same pack they stay interchangeable until the next one
        RECS_CUDA(c, cudaMemcpyAsync(c->posm_alt, c->posm, n_pad * 16, cudaMemcpyDeviceToDevice, c->stream));
        c->posm_alt_serial = c->posm_pack_serial;
    }
    ResidentTicksArgs ra;
    memset(&ra, 0, sizeof ra);
    ra.posm[0] = c->posm;
    ra.posm[1] = c->posm_alt;
    ra.pos = (float*)pos->dev;
    ra.vel = (float*)vel->dev;
    ra.acc = (float*)acc->dev;
    ra.n_rows = (uint32_t)n;
    ra.n_tiles = (uint32_t)(n_pad / kGravTile);
    ra.n_ticks = n_ticks;
    ra.dt = c->cfg.nbody_dt;
    ra.eps = c->cfg.nbody_epsilon;
    ra.barrier = c->resident_barrier;
    ra.barrier_base = c->resident_barrier_count;
    ra.error = c->dev_error;
    {
        NvtxRange range("resident ticks");
        const cudaError_t e = launch_resident_ticks(ra, c->stream);
        if (e == cudaErrorCooperativeLaunchTooLarge) {
            // the grid cannot be co-resident on this device (fewer SMs than CTAs): the ordinary chain runs the ticks
            cudaGetLastError();
            *launched = false;
            return RECS_OK;
        }
        RECS_CUDA(c, e);
    }
    c->kernel_launches++;
    c->gravity_launches++;
    c->resident_launches++;
    c->resident_barrier_count += (n_ticks - 1) * ((ra.n_rows + 31) / 32);  // one arrival per CTA and tick but the last
    if (n_ticks & 1) {  // the last tick refreshed the second buffer: it is the mirror now
        std::swap(c->posm, c->posm_alt);
        std::swap(c->posm_cap, c->posm_alt_cap);
    }
    pos->version += n_ticks;
    vel->version += n_ticks;
    acc->version += n_ticks;
    for (auto& src : c->posm_sources) src.second = find_column(c, src.first.first, src.first.second)->version;
    return RECS_OK;
}

int32_t recs_gpu_tick(recs_gpu_ctx* c, uint32_t n_ticks) {
    RECS_ENTER(c);
    if (n_ticks == 0 || c->order.empty()) return RECS_OK;
    if (!c->capturing && resident_ticks_apply(c)) {
        // with the mirror of exactly this archetype current, every tick of the call is in ONE launch; else the first tick goes
        // through the ordinary chain (binds, sizes, packs the mirror) and the rest follows in one launch
        if (!resident_mirror_ready(c)) {
            int32_t st = run_tick_once(c);
            if (st != RECS_OK) return st;
            if (--n_ticks == 0) return RECS_OK;
        }
        if (resident_mirror_ready(c)) {
            // (at most 2^20 ticks per launch: the barrier counter is compared through a signed 32-bit difference)
            bool launched = true;
            while (n_ticks && launched) {
                const uint32_t chunk = std::min<uint32_t>(n_ticks, 1u << 20);
                int32_t st = run_resident_ticks(c, chunk, &launched);
                if (st != RECS_OK) return st;
                if (launched) n_ticks -= chunk;
            }
            if (n_ticks == 0) return RECS_OK;
            c->resident_max_rows = 0;  // not on this device: the chain below, now and from here on
        }
    }
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (c->timing) {
        e0 = get_event(c);
        e1 = get_event(c);
        cudaEventRecord(e0, c->stream);
    }
    const bool want_graph = c->cfg.use_cuda_graph && c->cfg.world_size == 1 && !c->timing && order_is_capturable(c);
    uint32_t done = 0;
    if (want_graph && n_ticks > 1) {
        // a captured tick that found the gravity mirror current holds no pack kernel: it is only valid while nothing
        // outside the chain (recs_gpu_run_system, an upload) has written the mirror's source columns since
        if (c->graph && c->graph_elided_pack && !posm_is_current(c)) drop_graph(c);
        const uint32_t per_graph = std::max(1u, std::min(c->graph_ticks_wanted, (n_ticks - 1) / 2));
        if (!c->graph || c->graph_epoch != c->epoch || c->graph_ticks != per_graph) {
            drop_graph(c);
            // one eager tick sizes every workspace, then the same chain is captured
            int32_t st = run_tick_once(c);
            if (st != RECS_OK) return st;
            done = 1;
            const uint64_t epoch = c->epoch;
            cudaGraph_t g = nullptr;
            RECS_CUDA(c, cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
            c->capturing = true;
            for (uint32_t k = 0; k < per_graph && st == RECS_OK; ++k) st = run_tick_once(c);
            c->capturing = false;
            c->graph_ticks = per_graph;
            cudaError_t ce = cudaStreamEndCapture(c->stream, &g);
            if (st != RECS_OK) {
                if (g) cudaGraphDestroy(g);
                return st;
            }
            RECS_CUDA(c, ce);
            ce = cudaGraphInstantiate(&c->graph, g, 0);
            cudaGraphDestroy(g);
            RECS_CUDA(c, ce);
            c->graph_epoch = epoch;
        }
        for (; done + c->graph_ticks <= n_ticks; done += c->graph_ticks) RECS_CUDA(c, cudaGraphLaunch(c->graph, c->stream));
    }
    for (; done < n_ticks; ++done) {
        int32_t st = run_tick_once(c);
        if (st != RECS_OK) return st;
    }
    if (c->timing) {
        cudaEventRecord(e1, c->stream);
        TimedSpan s;
        s.a = e0;
        s.b = e1;
        s.slot = T_CHAIN;
        c->spans.push_back(s);
    }
    return RECS_OK;
}

int32_t recs_gpu_set_wind_direction(recs_gpu_ctx* c, const float xyz[3]) {
    RECS_ENTER(c);
    if (!xyz) return fail(c, RECS_ERR_INVALID_ARGUMENT, "set_wind_direction: null");
    memcpy(c->wind, xyz, sizeof c->wind);
    return RECS_OK;
}
int32_t recs_gpu_get_wind_direction(recs_gpu_ctx* c, float xyz[3]) {
    RECS_ENTER(c);
    if (!xyz) return fail(c, RECS_ERR_INVALID_ARGUMENT, "get_wind_direction: null");
    memcpy(xyz, c->wind, sizeof c->wind);
    return RECS_OK;
}

int32_t recs_gpu_comm_get_unique_id(recs_gpu_ctx* c, uint8_t id_bytes[RECS_GPU_NCCL_ID_BYTES]) {
    RECS_ENTER(c);
    const char* why = "";
    if (!c->nccl.load(&why)) return fail(c, RECS_ERR_NCCL, why);
    NcclApi::unique_id id;
    int r = c->nccl.GetUniqueId(&id);
    if (r != 0) return fail(c, RECS_ERR_NCCL, std::string("ncclGetUniqueId: ") + c->nccl.GetErrorString(r));
    memcpy(id_bytes, id.internal, RECS_GPU_NCCL_ID_BYTES);
    return RECS_OK;
}

int32_t recs_gpu_comm_init(recs_gpu_ctx* c, const uint8_t id_bytes[RECS_GPU_NCCL_ID_BYTES]) {
    RECS_ENTER(c);
    const char* why = "";
    if (!c->nccl.load(&why)) return fail(c, RECS_ERR_NCCL, why);
    if (c->comm) return fail(c, RECS_ERR_NCCL, "communicator already initialised");
    NcclApi::unique_id id;
    memcpy(id.internal, id_bytes, RECS_GPU_NCCL_ID_BYTES);
    int r = c->nccl.CommInitRank(&c->comm, c->cfg.world_size, id, c->cfg.rank);
    if (r != 0) return fail(c, RECS_ERR_NCCL, std::string("ncclCommInitRank: ") + c->nccl.GetErrorString(r));
    return RECS_OK;
}

// ---- peer-pushed position exchange (sharded runs) -------------------------------------------------------------------
int32_t recs_gpu_exchange_export(recs_gpu_ctx* c, uint64_t max_bodies, uint8_t handle_bytes[RECS_GPU_IPC_HANDLE_BYTES]) {
    RECS_ENTER(c);
    if (!handle_bytes || max_bodies == 0) return fail(c, RECS_ERR_INVALID_ARGUMENT, "exchange_export: bad argument");
    if (c->cfg.world_size > kGravMaxPeers) return fail(c, RECS_ERR_UNSUPPORTED, "exchange_export: more than 16 ranks");
    static_assert(sizeof(cudaIpcMemHandle_t) == RECS_GPU_IPC_HANDLE_BYTES, "IPC handle size");
    RECS_CUDA(c, cudaStreamSynchronize(c->stream));
    for (int r = 0; r < kGravMaxPeers; ++r) {
        if (c->xchg.peer[r] && c->xchg.peer[r] != c->xchg.base) RECS_CUDA(c, cudaIpcCloseMemHandle(c->xchg.peer[r]));
        c->xchg.peer[r] = nullptr;
    }
    if (c->xchg.base) RECS_CUDA(c, cudaFree(c->xchg.base));
    c->xchg = recs_gpu_ctx::Exchange();
    uint64_t slice, rb, re;
    shard_rows(c->cfg, max_bodies, &slice, &rb, &re);
    c->xchg.cap_bodies = slice * (uint64_t)c->cfg.world_size;
    RECS_CUDA(c, cudaMalloc((void**)&c->xchg.base, c->xchg.bytes()));
    // all zero: null bodies everywhere, generation 0 on every flag
    RECS_CUDA(c, cudaMemset(c->xchg.base, 0, c->xchg.bytes()));
    cudaIpcMemHandle_t h;
    RECS_CUDA(c, cudaIpcGetMemHandle(&h, c->xchg.base));
    memcpy(handle_bytes, &h, sizeof h);
    c->posm_sources.clear();  // the mirror moves into the exchange buffer: pack again
    c->posm_bodies = 0;
    return RECS_OK;
}

int32_t recs_gpu_exchange_import(recs_gpu_ctx* c, const uint8_t* handles, uint32_t n_handles) {
    RECS_ENTER(c);
    if (!handles || n_handles != (uint32_t)c->cfg.world_size) return fail(c, RECS_ERR_INVALID_ARGUMENT, "exchange_import: one handle per rank, in rank order");
    if (!c->xchg.base) return fail(c, RECS_ERR_INVALID_ARGUMENT, "exchange_import: call recs_gpu_exchange_export first");
    for (int r = 0; r < c->cfg.world_size; ++r) {
        if (r == c->cfg.rank) {
            c->xchg.peer[r] = c->xchg.base;
            continue;
        }
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + (size_t)r * RECS_GPU_IPC_HANDLE_BYTES, sizeof h);
        void* p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return fail(c, RECS_ERR_CUDA, std::string("cudaIpcOpenMemHandle (rank ") + std::to_string(r) + "): " + cudaGetErrorString(e));
        }
        c->xchg.peer[r] = (unsigned char*)p;
    }
    c->xchg.imported = true;
    return RECS_OK;
}

int32_t recs_gpu_exchange_mode(recs_gpu_ctx* c) {
    if (!c) return RECS_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lock(c->mu);
    if (c->cfg.world_size <= 1) return 0;
    return c->xchg.imported && !c->exchange_nccl_forced ? 2 : 1;
}

// Every rank's own rows of a column (its recs_gpu_shard_rows slice), delivered to all ranks: what makes a column
// written by a row-sharded system complete again (the nested Query of a generic system reads all of it).
int32_t recs_gpu_column_allgather(recs_gpu_ctx* c, uint32_t arch, uint64_t comp) {
    RECS_ENTER(c);
    if (c->cfg.world_size <= 1) return RECS_OK;
    if (!c->comm) return fail(c, RECS_ERR_NCCL, "column_allgather: recs_gpu_comm_init was not called");
    Column* col = find_column(c, arch, comp);
    if (!col || !col->dev) return fail(c, RECS_ERR_MISSING_PARAMETER, "column_allgather: column is not bound");
    if (col->stale) return fail(c, RECS_ERR_STALE_BINDING, "column_allgather: column binding is stale");
    // one in-place broadcast per owner, grouped: slices are contiguous row ranges, the last ones may be short or empty
    int r = c->nccl.GroupStart();
    for (int root = 0; root < c->cfg.world_size && r == 0; ++root) {
        uint64_t slice, rb, re;
        shard_rows_for(root, c->cfg.world_size, col->count, &slice, &rb, &re);
        if (re == rb) continue;
        unsigned char* p = (unsigned char*)col->dev + rb * col->elem;
        r = c->nccl.Broadcast(p, p, (size_t)(re - rb) * col->elem, /*ncclUint8*/ 1, root, c->comm, c->stream);
    }
    const int e = c->nccl.GroupEnd();
    if (r == 0) r = e;
    if (r != 0) return fail(c, RECS_ERR_NCCL, std::string("ncclBroadcast: ") + c->nccl.GetErrorString(r));
    col->version++;
    return RECS_OK;
}

int32_t recs_gpu_shard_rows(recs_gpu_ctx* c, uint64_t n_rows, uint64_t* row_begin, uint64_t* row_end) {
    if (!c || !row_begin || !row_end) return RECS_ERR_INVALID_ARGUMENT;
    uint64_t slice;
    shard_rows(c->cfg, n_rows, &slice, row_begin, row_end);
    return RECS_OK;
}

int32_t recs_gpu_shard_plan(uint64_t n_rows, int32_t rank, int32_t world_size, uint64_t* row_begin, uint64_t* row_end,
                            uint64_t* slice_rows) {
    if (world_size < 1 || rank < 0 || rank >= world_size || !row_begin || !row_end) return RECS_ERR_INVALID_ARGUMENT;
    recs_gpu_config cfg;
    recs_gpu_default_config(&cfg);
    cfg.rank = rank;
    cfg.world_size = world_size;
    uint64_t slice;
    shard_rows(cfg, n_rows, &slice, row_begin, row_end);
    if (slice_rows) *slice_rows = slice;
    return RECS_OK;
}

int32_t recs_gpu_enable_timing(recs_gpu_ctx* c, int32_t enabled) {
    RECS_ENTER(c);
    c->timing = enabled != 0;
    return RECS_OK;
}

int32_t recs_gpu_get_timing(recs_gpu_ctx* c, recs_gpu_timing* out) {
    RECS_ENTER(c);
    if (!out) return fail(c, RECS_ERR_INVALID_ARGUMENT, "get_timing: null");
    RECS_CUDA(c, cudaStreamSynchronize(c->stream));
    RECS_CUDA(c, cudaStreamSynchronize(c->comm_stream));
    float acc[T_COUNT] = {0};
    for (auto& s : c->spans) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, s.a, s.b) == cudaSuccess) {
            acc[s.slot] += ms;
        }
    }
    memset(out, 0, sizeof *out);
    out->last_tick_chain_ms = acc[T_CHAIN];
    out->gravity_ms = acc[T_GRAVITY];
    out->gravity_reduce_ms = acc[T_REDUCE];
    out->movement_ms = acc[T_MOVEMENT];
    out->acceleration_ms = acc[T_ACCEL];
    out->pack_ms = acc[T_PACK];
    out->other_ms = acc[T_OTHER];
    out->gather_ms = acc[T_GATHER];
    out->kernel_launches = c->kernel_launches;
    out->gravity_launches = c->gravity_launches;
    return RECS_OK;
}

int32_t recs_gpu_reset_timing(recs_gpu_ctx* c) {
    RECS_ENTER(c);
    RECS_CUDA(c, cudaStreamSynchronize(c->stream));
    for (auto& s : c->spans) {
        c->event_pool.push_back(s.a);
        c->event_pool.push_back(s.b);
    }
    c->spans.clear();
    c->kernel_launches = 0;
    c->gravity_launches = 0;
    return RECS_OK;
}

int32_t recs_gpu_timer_start(recs_gpu_ctx* c) {
    RECS_ENTER(c);
    if (!c->timer_a) {
        RECS_CUDA(c, cudaEventCreate(&c->timer_a));
        RECS_CUDA(c, cudaEventCreate(&c->timer_b));
    }
    RECS_CUDA(c, cudaEventRecord(c->timer_a, c->stream));
    return RECS_OK;
}

int32_t recs_gpu_timer_stop(recs_gpu_ctx* c, float* out_ms) {
    RECS_ENTER(c);
    if (!c->timer_a || !out_ms) return fail(c, RECS_ERR_INVALID_ARGUMENT, "timer_stop without timer_start");
    RECS_CUDA(c, cudaEventRecord(c->timer_b, c->stream));
    RECS_CUDA(c, cudaEventSynchronize(c->timer_b));
    RECS_CUDA(c, cudaStreamSynchronize(c->comm_stream));
    RECS_CUDA(c, cudaEventElapsedTime(out_ms, c->timer_a, c->timer_b));
    return check_device_error(c);
}

int32_t recs_gpu_measure_fp32_peak(recs_gpu_ctx* c, int32_t packed, float* out_tflops, float* out_sm_mhz) {
    RECS_ENTER(c);
    if (!out_tflops || !out_sm_mhz) return fail(c, RECS_ERR_INVALID_ARGUMENT, "measure_fp32_peak: null");
    const uint32_t grid = (uint32_t)c->sm_count * 4;  // one resident wave: 4 CTAs x 256 threads per SM
    const uint32_t iters = 4096;
    float* sink = nullptr;
    unsigned long long* cycles = nullptr;
    RECS_CUDA(c, cudaMalloc((void**)&sink, 16));
    RECS_CUDA(c, cudaMalloc((void**)&cycles, grid * sizeof(unsigned long long)));
    cudaEvent_t a = get_event(c), b = get_event(c);
    float best_ms = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(a, c->stream);
        cudaError_t e = launch_fp32_peak(packed, iters, grid, sink, cycles, c->stream);
        cudaEventRecord(b, c->stream);
        if (e != cudaSuccess || cudaStreamSynchronize(c->stream) != cudaSuccess) {
            cudaFree(sink);
            cudaFree(cycles);
            return fail(c, RECS_ERR_CUDA, "fp32 peak kernel failed");
        }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, a, b);
        if (rep > 0 && ms < best_ms) best_ms = ms;
        c->kernel_launches++;
    }
    std::vector<unsigned long long> h(grid);
    cudaMemcpy(h.data(), cycles, grid * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    unsigned long long cyc_max = 0;
    for (auto v : h) cyc_max = std::max(cyc_max, v);
    // FMAs: grid * 256 threads * iters * 4 * (12 packed x 2 | 24 scalar)
    const double fmas = (double)grid * 256.0 * iters * 4.0 * 24.0;
    *out_tflops = (float)(2.0 * fmas / (best_ms * 1e-3) / 1e12);
    *out_sm_mhz = (float)((double)cyc_max / (best_ms * 1e-3) / 1e6);
    c->event_pool.push_back(a);
    c->event_pool.push_back(b);
    cudaFree(sink);
    cudaFree(cycles);
    return RECS_OK;
}

int32_t recs_gpu_flush_l2(recs_gpu_ctx* c) {
    RECS_ENTER(c);
    // Write a buffer larger than the 126 MB L2, then stream a second one through it with loads: everything a
    // previous pass left is evicted, and the lines that remain are CLEAN -- after the write alone the next (timed)
    // kernel would pay for writing back up to 126 MB of the flush's own dirty lines.
    const size_t bytes = 256u << 20;
    int32_t st = ensure(c, &c->l2_flush, &c->l2_flush_bytes, 2 * bytes + 16);
    if (st != RECS_OK) return st;
    if (!c->l2_flush_zeroed) {
        RECS_CUDA(c, cudaMemsetAsync(c->l2_flush, 0, 2 * bytes + 16, c->stream));
        c->l2_flush_zeroed = true;
    }
    float* second = c->l2_flush + bytes / 4;
    RECS_CUDA(c, launch_fill(c->l2_flush, bytes / 4, 0.f, c->stream));
    RECS_CUDA(c, launch_read_sweep(second, bytes / 4, second + bytes / 4, c->stream));
    return RECS_OK;
}

int32_t recs_gpu_host_alloc(recs_gpu_ctx* c, size_t bytes, void** out_ptr) {
    RECS_ENTER(c);
    if (!out_ptr) return fail(c, RECS_ERR_INVALID_ARGUMENT, "host_alloc: null");
    RECS_CUDA(c, cudaHostAlloc(out_ptr, bytes ? bytes : 16, cudaHostAllocDefault));
    return RECS_OK;
}

int32_t recs_gpu_host_free(recs_gpu_ctx* c, void* ptr) {
    RECS_ENTER(c);
    if (ptr) RECS_CUDA(c, cudaFreeHost(ptr));
    return RECS_OK;
}

}  // extern "C"
