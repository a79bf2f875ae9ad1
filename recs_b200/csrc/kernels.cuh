// Internal C++ interface between the context (context.cu) and the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

namespace recs {

// ---- gravity (crates/n-body/src/lib.rs:104-135) ------------------------------------------------
constexpr int kGravMaxConsumerWarps = 31;                      // consumer warps per CTA are variant dependent (4, 8 or 12)
constexpr int kGravMaxRowsPerThread = 8;                       // register blocking over i (variant dependent)
constexpr int kGravTile = 256;                                 // j bodies per smem tile
constexpr int kGravStages = 4;
constexpr int kGravTileBytes = kGravTile * 16;                 // 4 KiB
constexpr int kGravSmemBytes = kGravStages * kGravTileBytes;   // 16 KiB dynamic smem

// Kernel variants (register blocking x resident CTAs per SM), selected per context.
struct GravityVariant {
    int rows_per_thread;  // bodies i per consumer thread
    int min_ctas;         // __launch_bounds__ residency target
    int consumer_warps;   // 4, 8 or 12 (+ 1 TMA producer warp)
    bool scaled = false;  // posm is stored range-shifted (kGravPosScale), see gravity.cu
    int consumer_threads() const { return 32 * consumer_warps; }
    int threads() const { return consumer_threads() + 32; }
    int rows() const { return consumer_threads() * rows_per_thread; }  // rows per i-block
};
#ifdef RECS_GPU_ABLATIONS
constexpr int kGravVariantCount = 11;  // 9 and 10 are timing-only ablations with WRONG results (tools/ builds only)
#else
constexpr int kGravVariantCount = 9;
#endif
constexpr int kGravVariantDefault = 0;
constexpr int kGravVariantSmall = 2;    // up to kGravSmallRows outer rows
constexpr int kGravVariantMedium = 3;   // up to kGravMediumRows outer rows
constexpr uint64_t kGravSmallRows = 5120;    // (profiles/r02_shape_thresholds.txt)
constexpr uint64_t kGravMediumRows = 22528;
GravityVariant gravity_variant(int index);

// Range shift of the overflow-detector variants: positions x 2^-32 (exact power-of-two scaling, G*m unscaled);
// tile sums come out x 2^64 and are scaled back before they join the running sums.
constexpr float kGravPosScale = 2.3283064365386963e-10f;   // 2^-32
constexpr float kGravPosUnscale = 4294967296.f;            // 2^32
constexpr float kGravOutScale = 5.4210108624275222e-20f;   // 2^-64 = pos_scale^2

// ---- the summation tree of one row (what makes a result independent of grid size, launch split and rank count) --
//   tile sum   : the 256 bodies of one j tile, in registers, starting from zero
//   chunk sum  : `chunk_tiles` consecutive tile sums (tiles [c*K, (c+1)*K) of the mirror), added in ascending
//                tile order starting from zero
//   row total  : the chunk sums, added in ascending chunk order starting from zero
// chunk_tiles is a function of the body count alone (gravity_chunk_tiles), so which CTA / launch / GPU produced a
// tile sum does not change a bit of the result.  The work of a launch is stream-K dealt as before: the units
// (i-block, tile) of a PHASE -- a rotated run of whole chunks -- are cut into equal contiguous spans, one per
// persistent CTA.  A CTA that enters a chunk at its first tile accumulates the chunk's running prefix and stores it
// once ("prefix fragment" of (i-block, chunk)); a CTA whose span begins in the middle of a chunk stores the tile sums of
// the rest of that chunk one by one ("tile fragments", at most K-1 per CTA and phase), and the reduce kernel continues
// the prefix with them in tile order.
constexpr uint32_t kGravTargetChunks = 32;
inline uint32_t gravity_chunk_tiles(uint64_t n_bodies) {
    const uint64_t tiles = (n_bodies + kGravTile - 1) / kGravTile;
    const uint64_t k = (tiles + kGravTargetChunks - 1) / kGravTargetChunks;
    return (uint32_t)(k ? k : 1);
}

// One phase: i-blocks [iblock_begin, iblock_begin + n_iblocks) against tiles [tile_offset, tile_offset + n_vtiles) (mod
// n_tiles_total).  A single-GPU launch has one tile range (all tiles); a rank of a sharded run visits its own slice
// first, then the other ranks' slices in rotated order.  A partly filled LAST i-block gets phases of its own: its warps
// without rows skip the arithmetic, so its units are cheaper than the others -- dealt evenly over all CTAs by themselves
// they cannot unbalance the stream-K split, whatever they cost.
struct GravityPhase {
    uint32_t n_vtiles;        // tiles visited per i-block: whole chunks
    uint32_t tile_offset;     // first tile visited: a chunk boundary
    uint32_t iblock_begin;    // first i-block of the phase
    uint32_t n_iblocks;       // i-blocks of the phase
    uint32_t frag_base;       // prefix fragment of (i-block b, chunk cv of the phase): frag_base + (b - iblock_begin) * (n_vtiles / K) + cv
    uint32_t tile_frag_base;  // tile fragment of CTA g, unit u: tile_frag_base + g * (K - 1) + (u - span begin of g)
};

constexpr int kGravMaxPhases = 4;
struct GravityWork {
    uint32_t grid;            // persistent CTAs (stream-K workers), the same for every phase
    uint32_t n_iblocks;       // ceil(row_count / rows per i-block)
    uint32_t n_tiles_total;   // tiles of the mirror: a multiple of chunk_tiles
    uint32_t chunk_tiles;     // K
    uint32_t n_phases;
    uint32_t n_own_phases;    // phases over the rank's own tiles come first (the NCCL exchange launches them separately)
    // Chunk sums of the rows of a separately dealt last i-block, worked out by nbody_gravity_chunksum_kernel before the
    // reduce (slot chunksum_base + c of the fragment workspace); ~0u: no such i-block.
    uint32_t heavy_iblock;
    uint32_t chunksum_base;
    GravityPhase phase[kGravMaxPhases];
};

// Number of fragment slots (of rows-per-i-block float4 each) a work description needs; assigns the bases.
inline uint32_t gravity_assign_frag_slots(GravityWork& w) {
    uint32_t next = 0;
    for (uint32_t p = 0; p < w.n_phases; ++p) {
        w.phase[p].frag_base = next;
        next += w.phase[p].n_iblocks * (w.phase[p].n_vtiles / w.chunk_tiles);
        w.phase[p].tile_frag_base = next;
        next += w.phase[p].n_vtiles ? w.grid * (w.chunk_tiles - 1) : 0;
    }
    if (w.heavy_iblock != 0xffffffffu) {
        w.chunksum_base = next;
        next += w.n_tiles_total / w.chunk_tiles;
    }
    return next;
}

struct GravityArgs {
    const float4* posm;   // pair-interleaved {x0 x1 y0 y1 | z0 z1 gm0 gm1}, n_tiles_total tiles
    const float* ipos;    // outer Position column, 3 floats per row
    float4* frags;        // fragment workspace, rows-per-i-block float4 per slot
    uint32_t row_begin;
    uint32_t row_count;
    float eps;
    uint32_t self_offset; // mirror index of outer row 0 when the outer archetype is part of the nested query, else ~0u:
                          // the tile holding a warp's own bodies goes straight to the exact path
    GravityWork work;
    uint32_t phase_begin, phase_end;  // phases this launch runs
    // sharded run with a peer-pushed mirror (null flags: every tile is ready when the kernel starts)
    const uint32_t* flags;    // flags[r] = newest mirror generation rank r has pushed into this GPU
    uint32_t wait_gen;        // generation this launch reads
    uint32_t tiles_per_rank;
    uint32_t own_rank;
    uint32_t* error;          // set to 1 when a wait timed out (host-mapped)
    int variant;
};

// Where refreshed mirror rows go besides the local mirror: the same slot of every peer's NEXT mirror buffer, followed
// by a generation flag (release, system scope) once every CTA of the pushing kernel has stored its rows.
constexpr int kGravMaxPeers = 16;
struct PeerPush {
    uint32_t n;                          // 0: single GPU / NCCL exchange, write `posm` of the tail / pack alone
    uint32_t gen;                        // generation being pushed
    float* posm[kGravMaxPeers];          // every rank's buffer for `gen` (own rank included)
    uint32_t* flag[kGravMaxPeers];       // &flags[own rank] on every rank
    const uint32_t* local_flags;         // own flags: all must have reached gen - 1 before the push overwrites
    uint32_t* done_counter;              // local: CTAs of the pushing kernel that have stored their rows
    uint32_t* error;
    uint32_t wait;                       // this launch is the first of its generation: wait for the peers before storing
    uint32_t signal;                     // this launch is the last of its generation: raise the flag when all CTAs are done
};
inline PeerPush no_peer_push() {
    PeerPush p;
    memset(&p, 0, sizeof p);
    return p;
}

// Optional fused tail of the n-body tick (schedule order gravity -> movement -> acceleration over the
// same rows): after the fragment sum the same thread integrates its row and refreshes posm.
struct GravityTail {
    float* pos;           // Position column (null: plain reduce)
    float* vel;           // Velocity column
    const float* mass;    // Mass column (for the posm refresh; may be null together with posm)
    float* posm;          // pair-interleaved mirror to refresh, or null
    uint32_t posm_offset; // posm body index of row 0 of this column
    float dt;
    float g;
    float pscale;         // range shift of the posm mirror (1 for the unscaled variants)
};

struct GravityReduceArgs {
    const float4* frags;
    float* acc;           // Acceleration column, 3 floats per row
    uint32_t row_begin;
    uint32_t row_count;
    uint32_t iblock_rows; // rows per i-block of the variant that produced the fragments
    GravityWork work;
    GravityTail tail;
    PeerPush push;
};

int gravity_max_ctas_per_sm(int variant);  // occupancy of the interaction kernel on the current device
cudaError_t gravity_prepare();  // one-time function attributes
cudaError_t launch_gravity(const GravityArgs& a, cudaStream_t s);
cudaError_t launch_gravity_reduce(const GravityReduceArgs& a, cudaStream_t s);
// Chunk sums of the rows of the separately dealt last i-block (a.work.heavy_iblock): one warp per row, lanes = chunks.
cudaError_t launch_gravity_chunksum(const GravityReduceArgs& a, cudaStream_t s);

// Many fused n-body ticks of a small single-archetype world in one cooperative launch (gravity.cu).
constexpr uint32_t kResidentMaxTiles = 16;  // mirror of up to 4 096 bodies in shared memory, 32 x tiles threads per CTA
struct ResidentTicksArgs {
    float* posm[2];    // two pair-interleaved mirror buffers holding the same G*m; tick k reads [k & 1] and refreshes the other
    float* pos;        // Position, Velocity, Acceleration columns of the archetype (row 0 = mirror body 0)
    float* vel;
    float* acc;
    uint32_t n_rows;
    uint32_t n_tiles;  // mirror size / 256
    uint32_t n_ticks;
    float dt;
    float eps;
    uint32_t* barrier; // arrival counter of the grid barrier; holds barrier_base at launch
    uint32_t barrier_base;
    uint32_t* error;   // raised if the grid barrier times out
};
cudaError_t launch_resident_ticks(const ResidentTicksArgs& a, cudaStream_t s);

// Position(12 B) + Mass(4 B) rows [row_begin, row_begin + n) -> pair-interleaved posm bodies [dst_offset, dst_offset+n),
// gm = G * m rounded once like the reference's `GRAVITATIONAL_CONSTANT * body_mass`; bodies [dst_offset + n,
// dst_offset + n + n_pad) become null bodies (gm = 0) so partially filled tiles contribute nothing.  With
// push.n > 0 the rows go to every peer's mirror instead of `posm`, followed by the generation flag.
cudaError_t launch_pack_posm(const float* pos, const float* mass, float g, float* posm,
                             uint64_t row_begin, uint64_t n, uint64_t n_pad, uint64_t dst_offset, float pscale,
                             const PeerPush& push, cudaStream_t s);

// Above this many bytes touched per launch a streaming kernel is HBM-bound (the L2 is 126 MB); below it, its
// columns stay L2-resident from one tick to the next.  Measured kernel choices hang on it (streaming.cu, context.cu).
constexpr uint64_t kStreamL2Footprint = 96ull << 20;

// ---- streaming systems ---------------------------------------------------------------------------
// y[k] = y[k] + x[k] * a for k in [0, n): separate IEEE mul and add (no FMA contraction), the
// arithmetic of `position.point += velocity * FIXED_TIME_STEP` (crates/n-body/src/lib.rs:92),
// `*velocity += acceleration * FIXED_TIME_STEP` (:101) and, with a == 1, simple_iter's
// `position.0 += velocity.0` (crates/ecs_bench_suite/src/recs/simple_iter.rs:48).
cudaError_t launch_axpy_rn(float* y, const float* x, float a, uint64_t n, cudaStream_t s);
// rain (crates/rain-simulation/src/lib.rs:62-96); vel/pos are 3-float rows.
cudaError_t launch_rain_gravity(float* vel, uint64_t n, float gravity_dt, float terminal,
                                cudaStream_t s);
cudaError_t launch_rain_wind(float* vel, uint64_t n, float wx, float wy, float wz, float terminal,
                             cudaStream_t s);
// wind -> gravity -> movement in one pass (the bench registration order's DAG).
cudaError_t launch_rain_fused(float* vel, float* pos, uint64_t n, float wx, float wy, float wz,
                              float gravity_dt, float terminal, float dt, cudaStream_t s);

// ---- removal-deciding systems (select.cu) --------------------------------------------------------
// flags[row] |= 1 if another row of `qpos` is closer than sqrt(radius2) (crates/n-body/examples/n_body_demo.rs:54-71)
cudaError_t launch_collision_flags(const float* qpos, uint32_t n_q, const float* ipos, uint32_t n_rows,
                                   int same_archetype, float radius2, uint32_t* flags, cudaStream_t s);
// flags[row] = pos[row].field <= threshold (crates/rain-simulation/src/lib.rs:152-156)
cudaError_t launch_below_flags(const float* pos, uint32_t n_rows, uint32_t field, float threshold, uint32_t* flags,
                               cudaStream_t s);
cudaError_t launch_compact_flags(const uint32_t* flags, uint32_t n_rows, uint32_t* out, uint32_t* counter, uint32_t cap,
                                 cudaStream_t s);

// row[dst[k]] = row[src[k]] (sources read before any destination is written); tmp holds n_moves rows
cudaError_t launch_rows_move(void* col, uint32_t elem, const uint32_t* dst, const uint32_t* src, uint32_t n_moves, void* tmp,
                             cudaStream_t s);

// ---- measurement ---------------------------------------------------------------------------------
cudaError_t launch_fp32_peak(int packed, uint32_t iters, uint32_t grid, float* sink,
                             unsigned long long* cycles, cudaStream_t s);
cudaError_t launch_fill(float* p, uint64_t n, float v, cudaStream_t s);
cudaError_t launch_read_sweep(const float* p, uint64_t n, float* sink, cudaStream_t s);

}  // namespace recs
