// Systems that only DECIDE which entities to remove (the device half of a `Commands::remove`):
//   * collision (crates/n-body/examples/n_body_demo.rs:54-71): an entity removes itself if any
//     OTHER entity of the nested Query<(Entity, Read<Position>)> is closer than COLLISION_RADIUS;
//   * ground_collision (crates/rain-simulation/src/lib.rs:152-156): remove if position.y <= 0.
// The kernels write one flag per outer row; `compact_flags_kernel` turns flags into a row list that
// the host plays back against the World (command-buffer playback, crates/ecs/src/lib.rs:321-323).
#include <algorithm>

#include "kernels.cuh"
#include "ptx.cuh"

namespace recs {
namespace {

constexpr int kSelThreads = 256;

// ---- collision (crates/n-body/examples/n_body_demo.rs:54-71) --------------------------------------------------------
//   for (other_entity, other_position) in bodies_query { if other_entity == entity { continue }
//       if position.point.distance2(other_position.point) < COLLISION_RADIUS * COLLISION_RADIUS { commands.remove(entity); break } }
// One launch per archetype of the nested query (flags are OR-ed).  Shape: the all-pairs tile loop of gravity.cu --
//   * j tiles of 256 rows are bulk-copied (1-D TMA, UBLKCP) STRAIGHT from the 12-byte Position column into a 3-stage
//     shared-memory ring by one producer thread (256 rows x 12 B = 3072 B: every tile starts 16-byte aligned), no
//     repacking, no __syncthreads; consumers read them with broadcast LDS;
//   * 4 outer rows per thread, two per f32x2 lane pair: FADD2 / FMUL2 for the differences and squares; the two sums of
//     d2 = (dx*dx + dy*dy) + dz*dz are scalar adds so that every product is rounded on its own like cgmath's
//     `distance2` (ptxas would contract packed mul + add, see ptx.cuh) -- the `< radius^2` decision is bit-exact;
//   * the reference's early exit (`break` after the first hit) at warp granularity: a warp whose 128 rows have all hit
//     skips the arithmetic of the remaining tiles (it still releases the ring's stages);
//   * `other_entity == entity` costs nothing outside the (at most five) tiles that overlap the CTA's own rows.
constexpr int kColTile = 256;                       // rows per j tile
constexpr int kColTileBytes = kColTile * 12;        // 3072
constexpr int kColStages = 3;
constexpr int kColWarps = 8;                        // consumer warps (+ 1 producer warp)
constexpr int kColRowsPerThread = 4;
constexpr int kColRows = kColWarps * 32 * kColRowsPerThread;  // 1024 outer rows per CTA

template <bool CHECK_SELF>
__device__ __forceinline__ void collision_tile(const float* __restrict__ t, uint32_t n_here, uint32_t tile_row0,
                                               const ptx::f32x2 (&nx)[2], const ptx::f32x2 (&ny)[2], const ptx::f32x2 (&nz)[2],
                                               const uint32_t (&row)[4], float (&dmin)[4]) {
#pragma unroll 4
    for (uint32_t k = 0; k < n_here; ++k) {
        const float xj = t[3 * k], yj = t[3 * k + 1], zj = t[3 * k + 2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            // other_position - position, for rows (2q, 2q+1) of this thread
            const ptx::f32x2 dx = ptx::add2(nx[q], ptx::pack(xj, xj));
            const ptx::f32x2 dy = ptx::add2(ny[q], ptx::pack(yj, yj));
            const ptx::f32x2 dz = ptx::add2(nz[q], ptx::pack(zj, zj));
            ptx::f32x2 d2 = ptx::add2_unfused(ptx::mul2(dx, dx), ptx::mul2(dy, dy));
            d2 = ptx::add2_unfused(d2, ptx::mul2(dz, dz));
            float a, b;
            ptx::unpack(d2, a, b);
            if (CHECK_SELF) {  // `if other_entity == entity { continue }`: the same row of the same archetype
                if (tile_row0 + k == row[2 * q]) a = 3.0e38f;
                if (tile_row0 + k == row[2 * q + 1]) b = 3.0e38f;
            }
            dmin[2 * q] = fminf(dmin[2 * q], a);
            dmin[2 * q + 1] = fminf(dmin[2 * q + 1], b);
        }
    }
}

__global__ void __launch_bounds__(32 * kColWarps + 32)
    collision_flags_kernel(const float* __restrict__ qpos, uint32_t n_q, const float* __restrict__ ipos, uint32_t n_rows,
                           int same_archetype, float radius2, uint32_t* __restrict__ flags) {
    __shared__ __align__(128) unsigned char tiles[kColStages * kColTileBytes];
    __shared__ __align__(8) uint64_t full_bar[kColStages], empty_bar[kColStages];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // blockIdx.y: this CTA's share of the query's tiles (flags are OR-ed, so the nested query may be cut freely;
    // small outer archetypes still fill the machine)
    const uint32_t all_tiles = (n_q + kColTile - 1) / kColTile;
    const uint32_t per_split = (all_tiles + gridDim.y - 1) / gridDim.y;
    const uint32_t tile_lo = min(all_tiles, blockIdx.y * per_split);
    const uint32_t n_tiles = min(all_tiles, tile_lo + per_split) - tile_lo;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kColStages; ++s) {
            ptx::mbar_init(&full_bar[s], 1);
            ptx::mbar_init(&empty_bar[s], kColWarps);
        }
        ptx::mbar_fence_init();
    }
    __syncthreads();
    if (warp == kColWarps) {
        if (lane == 0)
            for (uint32_t k = 0; k < n_tiles; ++k) {
                const uint32_t stage = k % kColStages, phase = (k / kColStages) & 1u;
                ptx::mbar_wait(&empty_bar[stage], phase ^ 1u);
                // whole 16-byte units: the last tile may read a few bytes past its last row (inside the column's
                // allocation, which is rounded up to 256 bytes); the loop bound keeps those bytes out of the result
                const uint32_t rows_here = min((uint32_t)kColTile, n_q - (tile_lo + k) * kColTile);
                const uint32_t bytes = (rows_here * 12u + 15u) & ~15u;
                ptx::mbar_arrive_expect_tx(&full_bar[stage], bytes);
                ptx::bulk_g2s(tiles + stage * kColTileBytes,
                              reinterpret_cast<const unsigned char*>(qpos) + (size_t)(tile_lo + k) * kColTileBytes, bytes, &full_bar[stage]);
            }
        return;
    }
    const uint32_t cta_row0 = blockIdx.x * kColRows;
    uint32_t row[4];
    ptx::f32x2 nx[2], ny[2], nz[2];
    float px[4], py[4], pz[4], dmin[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        row[r] = cta_row0 + warp * 128 + r * 32 + lane;
        const bool valid = row[r] < n_rows;
        px[r] = valid ? ipos[3ull * row[r]] : 0.f;
        py[r] = valid ? ipos[3ull * row[r] + 1] : 0.f;
        pz[r] = valid ? ipos[3ull * row[r] + 2] : 0.f;
        dmin[r] = valid ? 3.0e38f : -1.f;  // rows past the end count as decided: they never hold the warp back
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        nx[q] = ptx::pack(-px[2 * q], -px[2 * q + 1]);
        ny[q] = ptx::pack(-py[2 * q], -py[2 * q + 1]);
        nz[q] = ptx::pack(-pz[2 * q], -pz[2 * q + 1]);
    }
    bool warp_done = false;
    for (uint32_t k = 0; k < n_tiles; ++k) {
        const uint32_t stage = k % kColStages, phase = (k / kColStages) & 1u;
        ptx::mbar_wait(&full_bar[stage], phase);
        if (!warp_done) {
            const float* t = reinterpret_cast<const float*>(tiles + stage * kColTileBytes);
            const uint32_t tile_row0 = (tile_lo + k) * kColTile;
            const uint32_t n_here = min((uint32_t)kColTile, n_q - tile_row0);
            // only tiles that overlap this warp's own 128 rows can hold `other_entity == entity`
            const uint32_t w0 = cta_row0 + warp * 128;
            if (same_archetype && tile_row0 < w0 + 128 && w0 < tile_row0 + n_here)
                collision_tile<true>(t, n_here, tile_row0, nx, ny, nz, row, dmin);
            else
                collision_tile<false>(t, n_here, tile_row0, nx, ny, nz, row, dmin);
            // the reference stops at the first hit: once every row of the warp has one, the rest is skipped
            const bool all_hit = dmin[0] < radius2 && dmin[1] < radius2 && dmin[2] < radius2 && dmin[3] < radius2;
            warp_done = __all_sync(0xffffffffu, all_hit);
        }
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&empty_bar[stage]);
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
        if (row[r] < n_rows && dmin[r] < radius2 && dmin[r] >= 0.f) flags[row[r]] = 1u;
}

__global__ void __launch_bounds__(kSelThreads)
    below_flags_kernel(const float* __restrict__ pos, uint32_t n_rows, uint32_t field, float threshold,
                       uint32_t* __restrict__ flags) {
    const uint32_t row = blockIdx.x * kSelThreads + threadIdx.x;
    if (row < n_rows) flags[row] = pos[3ull * row + field] <= threshold ? 1u : 0u;
}

// flags -> unordered list of flagged rows (warp-aggregated append); the host sorts it ascending,
// which fixes the playback order.
__global__ void __launch_bounds__(kSelThreads)
    compact_flags_kernel(const uint32_t* __restrict__ flags, uint32_t n_rows, uint32_t* __restrict__ out,
                         uint32_t* __restrict__ counter, uint32_t cap) {
    const uint32_t row = blockIdx.x * kSelThreads + threadIdx.x;
    const bool f = row < n_rows && flags[row] != 0;
    const unsigned ballot = __ballot_sync(0xffffffffu, f);
    if (!ballot) return;
    const int lane = threadIdx.x & 31;
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(counter, __popc(ballot));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (f) {
        const uint32_t slot = base + __popc(ballot & ((1u << lane) - 1));
        if (slot < cap) out[slot] = row;
    }
}

}  // namespace

cudaError_t launch_collision_flags(const float* qpos, uint32_t n_q, const float* ipos, uint32_t n_rows,
                                   int same_archetype, float radius2, uint32_t* flags, cudaStream_t s) {
    if (n_rows == 0 || n_q == 0) return cudaSuccess;
    const uint32_t row_blocks = (n_rows + kColRows - 1) / kColRows;
    const uint32_t tiles = (n_q + kColTile - 1) / kColTile;
    // about two waves of CTAs on 148 SMs (5 CTAs of 288 threads fit an SM), at least 8 tiles per CTA
    uint32_t splits = std::max(1u, std::min((2u * 148u * 5u + row_blocks - 1) / row_blocks, (tiles + 7) / 8));
    collision_flags_kernel<<<dim3(row_blocks, splits), 32 * kColWarps + 32, 0, s>>>(qpos, n_q, ipos, n_rows, same_archetype,
                                                                                    radius2, flags);
    return cudaGetLastError();
}

cudaError_t launch_below_flags(const float* pos, uint32_t n_rows, uint32_t field, float threshold, uint32_t* flags,
                               cudaStream_t s) {
    if (n_rows == 0) return cudaSuccess;
    below_flags_kernel<<<(n_rows + kSelThreads - 1) / kSelThreads, kSelThreads, 0, s>>>(pos, n_rows, field, threshold,
                                                                                         flags);
    return cudaGetLastError();
}

cudaError_t launch_compact_flags(const uint32_t* flags, uint32_t n_rows, uint32_t* out, uint32_t* counter, uint32_t cap,
                                 cudaStream_t s) {
    if (n_rows == 0) return cudaSuccess;
    compact_flags_kernel<<<(n_rows + kSelThreads - 1) / kSelThreads, kSelThreads, 0, s>>>(flags, n_rows, out, counter, cap);
    return cudaGetLastError();
}

// ---- structural changes on resident columns ------------------------------------------------------------------
// Replay of a batch of swap-removes (crates/ecs/src/archetypes.rs:233-243, 315-328) on a device column:
// row[dst[k]] = row[src[k]] for all k, every source taken from the column as it was BEFORE the batch.  A row can
// be the source of one move and the destination of another, hence gather into a scratch buffer, then scatter.
namespace {
template <typename W>
__global__ void __launch_bounds__(256) rows_gather_kernel(const W* __restrict__ col, const uint32_t* __restrict__ src,
                                                          W* __restrict__ tmp, uint32_t words_per_row, uint64_t total) {
    const uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (i >= total) return;
    const uint64_t k = i / words_per_row, w = i % words_per_row;
    tmp[i] = col[static_cast<uint64_t>(src[k]) * words_per_row + w];
}
template <typename W>
__global__ void __launch_bounds__(256) rows_scatter_kernel(W* __restrict__ col, const uint32_t* __restrict__ dst,
                                                           const W* __restrict__ tmp, uint32_t words_per_row, uint64_t total) {
    const uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (i >= total) return;
    const uint64_t k = i / words_per_row, w = i % words_per_row;
    col[static_cast<uint64_t>(dst[k]) * words_per_row + w] = tmp[i];
}
}  // namespace

cudaError_t launch_rows_move(void* col, uint32_t elem, const uint32_t* dst, const uint32_t* src, uint32_t n_moves, void* tmp,
                             cudaStream_t s) {
    if (n_moves == 0) return cudaSuccess;
    if (elem % 4 == 0) {
        const uint32_t wpr = elem / 4;
        const uint64_t total = static_cast<uint64_t>(n_moves) * wpr;
        const uint32_t grid = static_cast<uint32_t>((total + 255) / 256);
        rows_gather_kernel<uint32_t><<<grid, 256, 0, s>>>(static_cast<const uint32_t*>(col), src, static_cast<uint32_t*>(tmp), wpr, total);
        rows_scatter_kernel<uint32_t><<<grid, 256, 0, s>>>(static_cast<uint32_t*>(col), dst, static_cast<const uint32_t*>(tmp), wpr, total);
    } else {
        const uint64_t total = static_cast<uint64_t>(n_moves) * elem;
        const uint32_t grid = static_cast<uint32_t>((total + 255) / 256);
        rows_gather_kernel<uint8_t><<<grid, 256, 0, s>>>(static_cast<const uint8_t*>(col), src, static_cast<uint8_t*>(tmp), elem, total);
        rows_scatter_kernel<uint8_t><<<grid, 256, 0, s>>>(static_cast<uint8_t*>(col), dst, static_cast<const uint8_t*>(tmp), elem, total);
    }
    return cudaGetLastError();
}

}  // namespace recs
