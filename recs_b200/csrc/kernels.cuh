// Internal C++ interface between the context (context.cu) and the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace recs {

// ---- gravity (crates/n-body/src/lib.rs:104-135) ------------------------------------------------
constexpr int kGravMaxConsumerWarps = 31;                      // consumer warps per CTA are variant dependent (4, 8 or 12)
constexpr int kGravMaxRowsPerThread = 8;                       // register blocking over i (variant dependent)
constexpr int kGravTile = 256;                                 // j bodies per smem tile
constexpr int kGravStages = 4;
constexpr int kGravTileBytes = kGravTile * 16;                 // 4 KiB
constexpr int kGravSmemBytes = kGravStages * kGravTileBytes;   // 16 KiB dynamic smem

// One launch of the interaction kernel: rows [row_begin, row_begin+row_count) of the outer
// Position column against j tiles [tile_offset, tile_offset + n_vtiles) (mod n_tiles_total).
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
constexpr int kGravVariantCount = 10;
constexpr int kGravVariantDefault = 0;
constexpr int kGravVariantSmall = 2;    // up to kGravSmallRows outer rows
constexpr int kGravVariantMedium = 3;   // up to kGravMediumRows outer rows
constexpr uint64_t kGravSmallRows = 16384;
constexpr uint64_t kGravMediumRows = 49152;
GravityVariant gravity_variant(int index);

// Range shift of the overflow-detector variants: positions x 2^-32 (exact power-of-two scaling, G*m unscaled);
// tile sums come out x 2^64 and are scaled back before they join the running sums.
constexpr float kGravPosScale = 2.3283064365386963e-10f;   // 2^-32
constexpr float kGravPosUnscale = 4294967296.f;            // 2^32
constexpr float kGravOutScale = 5.4210108624275222e-20f;   // 2^-64 = pos_scale^2

struct GravityLaunch {
    uint32_t grid;        // persistent CTAs (stream-K workers)
    uint32_t n_iblocks;   // ceil(row_count / rows per i-block)
    uint32_t n_vtiles;    // tiles visited by this launch
    uint32_t tile_offset; // first tile visited
    uint32_t frag_base;   // first fragment slot of this launch
};

struct GravityArgs {
    const float4* posm;   // pair-interleaved {x0 x1 y0 y1 | z0 z1 gm0 gm1}, n_tiles_total tiles
    const float* ipos;    // outer Position column, 3 floats per row
    float4* frags;        // fragment workspace, rows-per-i-block float4 per slot
    uint32_t row_begin;
    uint32_t row_count;
    uint32_t n_tiles_total;
    float eps;
    GravityLaunch launch;
    int variant;
};

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
    uint32_t n_launches;  // 1 (single GPU) or 2 (local tiles, then remote tiles)
    uint32_t iblock_rows; // rows per i-block of the variant that produced the fragments
    GravityLaunch launch[2];
    GravityTail tail;
};

// Number of fragment slots a launch may touch.
inline uint32_t gravity_frag_slots(const GravityLaunch& l) { return l.grid + l.n_iblocks; }
int gravity_max_ctas_per_sm(int variant);  // occupancy of the interaction kernel on the current device
cudaError_t gravity_prepare();  // one-time function attributes
cudaError_t launch_gravity(const GravityArgs& a, cudaStream_t s);
cudaError_t launch_gravity_reduce(const GravityReduceArgs& a, cudaStream_t s);

// Position(12 B) + Mass(4 B) rows [0, n) -> pair-interleaved posm bodies [dst_offset, dst_offset+n),
// gm = G * m rounded once like the reference's `GRAVITATIONAL_CONSTANT * body_mass`.
cudaError_t launch_pack_posm(const float* pos, const float* mass, float g, float* posm,
                             uint64_t row_begin, uint64_t n, uint64_t dst_offset, float pscale, cudaStream_t s);
// Null bodies (gm = 0) for [begin, end) so partially filled tiles contribute nothing.
cudaError_t launch_pad_posm(float* posm, uint64_t begin, uint64_t end, cudaStream_t s);

// Above this many bytes touched per launch a streaming kernel is HBM-bound (the L2 is 126 MB); below it, its
// columns stay L2-resident from one tick to the next.  Measured kernel choices hang on it (streaming.cu, context.cu).
constexpr uint64_t kStreamL2Footprint = 96ull << 20;

// ---- streaming systems ---------------------------------------------------------------------------
// y[k] = y[k] + x[k] * a for k in [0, n): separate IEEE mul and add (no FMA contraction), the
// arithmetic of `position.point += velocity * FIXED_TIME_STEP` (crates/n-body/src/lib.rs:92),
// `*velocity += acceleration * FIXED_TIME_STEP` (:101) and, with a == 1, simple_iter's
// `position.0 += velocity.0` (crates/ecs_bench_suite/src/recs/simple_iter.rs:48).
cudaError_t launch_axpy_rn(float* y, const float* x, float a, uint64_t n, cudaStream_t s);
// movement fused with the posm refresh of the moved rows (n-body tick): pos += vel*dt and
// posm[dst_offset + row] = {pos, G*m}.
cudaError_t launch_movement_pack(float* pos, const float* vel, const float* mass, float dt, float g,
                                 float* posm, uint64_t row_begin, uint64_t n, uint64_t dst_offset,
                                 cudaStream_t s);

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
