// Systems that only DECIDE which entities to remove (the device half of a `Commands::remove`):
//   * collision (crates/n-body/examples/n_body_demo.rs:54-71): an entity removes itself if any
//     OTHER entity of the nested Query<(Entity, Read<Position>)> is closer than COLLISION_RADIUS;
//   * ground_collision (crates/rain-simulation/src/lib.rs:152-156): remove if position.y <= 0.
// The kernels write one flag per outer row; `compact_flags_kernel` turns flags into a row list that
// the host plays back against the World (command-buffer playback, crates/ecs/src/lib.rs:321-323).
#include "kernels.cuh"

namespace recs {
namespace {

constexpr int kSelThreads = 256;

// One thread per outer row against the rows of ONE archetype of the nested query (the host launches
// it once per matched archetype; flags are OR-ed).  j rows are staged in shared memory in tiles of
// 256, straight from the 12-byte Position column.  d2 is formed exactly like cgmath's
// `distance2` = (dx*dx + dy*dy) + dz*dz with separate roundings, so `< radius^2` is bit-exact.
__global__ void __launch_bounds__(kSelThreads)
    collision_flags_kernel(const float* __restrict__ qpos, uint32_t n_q, const float* __restrict__ ipos, uint32_t n_rows,
                           int same_archetype, float radius2, uint32_t* __restrict__ flags) {
    __shared__ float sx[kSelThreads], sy[kSelThreads], sz[kSelThreads];
    const uint32_t row = blockIdx.x * kSelThreads + threadIdx.x;
    const bool valid = row < n_rows;
    float x = 0.f, y = 0.f, z = 0.f;
    if (valid) {
        x = ipos[3ull * row];
        y = ipos[3ull * row + 1];
        z = ipos[3ull * row + 2];
    }
    bool hit = false;
    for (uint32_t base = 0; base < n_q; base += kSelThreads) {
        __syncthreads();
        const uint32_t j = base + threadIdx.x;
        if (j < n_q) {
            sx[threadIdx.x] = qpos[3ull * j];
            sy[threadIdx.x] = qpos[3ull * j + 1];
            sz[threadIdx.x] = qpos[3ull * j + 2];
        }
        __syncthreads();
        const uint32_t n_here = min(static_cast<uint32_t>(kSelThreads), n_q - base);
        // `if other_entity == entity { continue }`: the same row of the same archetype
        const uint32_t self = same_archetype && row >= base && row < base + n_here ? row - base : 0xffffffffu;
        float dmin = 3.0e38f;
#pragma unroll 8
        for (uint32_t k = 0; k < n_here; ++k) {
            const float dx = __fsub_rn(sx[k], x), dy = __fsub_rn(sy[k], y), dz = __fsub_rn(sz[k], z);
            float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
            if (k == self) d2 = 3.0e38f;
            dmin = fminf(dmin, d2);
        }
        hit = hit || dmin < radius2;
    }
    if (valid && hit) flags[row] = 1u;
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
    collision_flags_kernel<<<(n_rows + kSelThreads - 1) / kSelThreads, kSelThreads, 0, s>>>(qpos, n_q, ipos, n_rows,
                                                                                             same_archetype, radius2, flags);
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
