// HBM-bound streaming systems for sm_100a: the par-iter query path of recs
// (crates/ecs/src/systems/iteration.rs:151-245) for plain `(Write<A>, Read<B>)` f32-vector
// systems.  One grid replaces the reference's per-segment tasks; columns are read and written as
// flat float arrays with 128-bit accesses, and every thread keeps several independent 16-byte
// loads in flight so that ~35 KB per SM are outstanding (what 6.5 TB/s x HBM latency needs).
//
// Arithmetic is kept bit-identical to the Rust code: rustc never contracts `a + b * c` into an
// FMA, so every mul/add below is an explicit round-to-nearest intrinsic.
#include "kernels.cuh"
#include "ptx.cuh"
#include <stdlib.h>

namespace recs {
namespace {

constexpr int kStreamThreads = 256;
constexpr int kStreamUnroll = 4;  // float4 per thread per array in flight

__device__ __forceinline__ float4 ld_stream(const float4* p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream(float4* p, float4 v) { __stcs(p, v); }

// y = y + x * a over n4 float4 (both 16-byte aligned).
__global__ void __launch_bounds__(kStreamThreads) axpy_rn_vec_kernel(float4* __restrict__ y,
                                                                     const float4* __restrict__ x,
                                                                     float a, uint64_t n4) {
    const uint64_t chunk = static_cast<uint64_t>(kStreamThreads) * kStreamUnroll;
    for (uint64_t base = blockIdx.x * chunk; base < n4; base += gridDim.x * chunk) {
        float4 yv[kStreamUnroll], xv[kStreamUnroll];
#pragma unroll
        for (int u = 0; u < kStreamUnroll; ++u) {
            const uint64_t i = base + u * kStreamThreads + threadIdx.x;
            if (i < n4) {
                xv[u] = ld_stream(x + i);
                yv[u] = ld_stream(y + i);
            }
        }
#pragma unroll
        for (int u = 0; u < kStreamUnroll; ++u) {
            const uint64_t i = base + u * kStreamThreads + threadIdx.x;
            if (i < n4) {
                float4 r;
                r.x = __fadd_rn(yv[u].x, __fmul_rn(xv[u].x, a));
                r.y = __fadd_rn(yv[u].y, __fmul_rn(xv[u].y, a));
                r.z = __fadd_rn(yv[u].z, __fmul_rn(xv[u].z, a));
                r.w = __fadd_rn(yv[u].w, __fmul_rn(xv[u].w, a));
                st_stream(y + i, r);
            }
        }
    }
}

__global__ void __launch_bounds__(kStreamThreads) axpy_rn_scalar_kernel(float* __restrict__ y,
                                                                        const float* __restrict__ x,
                                                                        float a, uint64_t n) {
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < n; i += stride)
        y[i] = __fadd_rn(y[i], __fmul_rn(x[i], a));
}

// movement + posm refresh, one thread per body.
__global__ void __launch_bounds__(kStreamThreads)
    movement_pack_kernel(float* __restrict__ pos, const float* __restrict__ vel,
                         const float* __restrict__ mass, float dt, float g, float* __restrict__ posm,
                         uint64_t row_begin, uint64_t n, uint64_t dst_offset) {
    const uint64_t k = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (k >= n) return;
    const uint64_t row = row_begin + k;
    const float x = __fadd_rn(pos[3 * row + 0], __fmul_rn(vel[3 * row + 0], dt));
    const float y = __fadd_rn(pos[3 * row + 1], __fmul_rn(vel[3 * row + 1], dt));
    const float z = __fadd_rn(pos[3 * row + 2], __fmul_rn(vel[3 * row + 2], dt));
    pos[3 * row + 0] = x;
    pos[3 * row + 1] = y;
    pos[3 * row + 2] = z;
    const uint64_t J = dst_offset + k;
    float* d = posm + (J >> 1) * 8 + (J & 1);
    d[0] = x;
    d[2] = y;
    d[4] = z;
    d[6] = __fmul_rn(g, mass[row]);
}

// ---- rain --------------------------------------------------------------------------------------
// Four entities (12 floats = three float4) per thread per column.
struct Vec3x4 {
    float4 a, b, c;  // x0 y0 z0 x1 | y1 z1 x2 y2 | z2 x3 y3 z3
};
__device__ __forceinline__ void get3(const Vec3x4& v, int e, float& x, float& y, float& z) {
    const float f[12] = {v.a.x, v.a.y, v.a.z, v.a.w, v.b.x, v.b.y, v.b.z, v.b.w, v.c.x, v.c.y, v.c.z, v.c.w};
    x = f[3 * e];
    y = f[3 * e + 1];
    z = f[3 * e + 2];
}

// gravity: `if abs(v.y) >= TERMINAL { return } v += -G * dt * unit_y`
//   (crates/rain-simulation/src/lib.rs:65-70).  -9.82*0.05 is folded by rustc at compile time in
//   f32 ((-G) * dt), then multiplied into unit_y = (0,1,0): x += k*0, y += k*1, z += k*0.
__device__ __forceinline__ void rain_gravity_1(float& x, float& y, float& z, float k, float terminal) {
    if (fabsf(y) >= terminal) return;
    x = __fadd_rn(x, __fmul_rn(k, 0.f));
    y = __fadd_rn(y, __fmul_rn(k, 1.f));
    z = __fadd_rn(z, __fmul_rn(k, 0.f));
}
// wind: `if |v| >= TERMINAL { return } v += (WIND_ACCELERATION * dt) * wind_direction`
//   (crates/rain-simulation/src/lib.rs:86-92); magnitude = sqrt((x*x + y*y) + z*z).
//   (wx, wy, wz) arrive pre-multiplied by WIND_ACCELERATION * dt on the host, in f32, which is
//   the same `f32 * Vector3` product the reference forms before the `+=`.
__device__ __forceinline__ void rain_wind_1(float& x, float& y, float& z, float wx, float wy, float wz,
                                            float terminal) {
    const float m2 = __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
    if (__fsqrt_rn(m2) >= terminal) return;
    x = __fadd_rn(x, wx);
    y = __fadd_rn(y, wy);
    z = __fadd_rn(z, wz);
}

template <int MODE>  // 0 = gravity only, 1 = wind only, 2 = wind -> gravity -> movement
__global__ void __launch_bounds__(kStreamThreads)
    rain_kernel(float* __restrict__ vel, float* __restrict__ pos, uint64_t n, float wx, float wy,
                float wz, float k_gravity, float terminal, float dt) {
    const uint64_t groups = n / 4;  // full groups of 4 entities
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t gidx = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; gidx < groups;
         gidx += stride) {
        float4* v4 = reinterpret_cast<float4*>(vel) + 3 * gidx;
        Vec3x4 v;
        v.a = __ldcs(v4 + 0);
        v.b = __ldcs(v4 + 1);
        v.c = __ldcs(v4 + 2);
        Vec3x4 p;
        float4* p4 = nullptr;
        if (MODE == 2) {
            p4 = reinterpret_cast<float4*>(pos) + 3 * gidx;
            p.a = __ldcs(p4 + 0);
            p.b = __ldcs(p4 + 1);
            p.c = __ldcs(p4 + 2);
        }
        float o[12], q[12];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float x, y, z;
            get3(v, e, x, y, z);
            if (MODE == 1 || MODE == 2) rain_wind_1(x, y, z, wx, wy, wz, terminal);
            if (MODE == 0 || MODE == 2) rain_gravity_1(x, y, z, k_gravity, terminal);
            o[3 * e] = x;
            o[3 * e + 1] = y;
            o[3 * e + 2] = z;
            if (MODE == 2) {
                float px, py, pz;
                get3(p, e, px, py, pz);
                q[3 * e] = __fadd_rn(px, __fmul_rn(x, dt));
                q[3 * e + 1] = __fadd_rn(py, __fmul_rn(y, dt));
                q[3 * e + 2] = __fadd_rn(pz, __fmul_rn(z, dt));
            }
        }
        __stcs(v4 + 0, make_float4(o[0], o[1], o[2], o[3]));
        __stcs(v4 + 1, make_float4(o[4], o[5], o[6], o[7]));
        __stcs(v4 + 2, make_float4(o[8], o[9], o[10], o[11]));
        if (MODE == 2) {
            __stcs(p4 + 0, make_float4(q[0], q[1], q[2], q[3]));
            __stcs(p4 + 1, make_float4(q[4], q[5], q[6], q[7]));
            __stcs(p4 + 2, make_float4(q[8], q[9], q[10], q[11]));
        }
    }
    // tail entities (n % 4), handled by the first threads of block 0
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const uint64_t e = groups * 4 + threadIdx.x;
        float x = vel[3 * e], y = vel[3 * e + 1], z = vel[3 * e + 2];
        if (MODE == 1 || MODE == 2) rain_wind_1(x, y, z, wx, wy, wz, terminal);
        if (MODE == 0 || MODE == 2) rain_gravity_1(x, y, z, k_gravity, terminal);
        vel[3 * e] = x;
        vel[3 * e + 1] = y;
        vel[3 * e + 2] = z;
        if (MODE == 2) {
            pos[3 * e] = __fadd_rn(pos[3 * e], __fmul_rn(x, dt));
            pos[3 * e + 1] = __fadd_rn(pos[3 * e + 1], __fmul_rn(y, dt));
            pos[3 * e + 2] = __fadd_rn(pos[3 * e + 2], __fmul_rn(z, dt));
        }
    }
}


// Same systems, full 1024-entity chunks: the 12-byte rows are moved with perfectly coalesced
// 128-bit accesses (thread t touches float4 t, t+256, t+512 of the chunk) and regrouped into whole
// entities through shared memory (each thread then owns entities 4t..4t+3 = float4 3t..3t+2; a
// 48-byte stride is conflict-free for LDS.128 quarter-warps).
constexpr int kRainChunk = 1024;                      // entities per block iteration
constexpr int kRainChunkVec = kRainChunk * 3 / 4;     // 768 float4 per column per chunk
template <int MODE>
__global__ void __launch_bounds__(kStreamThreads)
    rain_staged_kernel(float4* __restrict__ vel, float4* __restrict__ pos, uint64_t n_chunks, float wx, float wy,
                       float wz, float k_gravity, float terminal, float dt) {
    __shared__ float4 sv[kRainChunkVec];
    __shared__ float4 sp[MODE == 2 ? kRainChunkVec : 1];
    const int t = threadIdx.x;
    for (uint64_t chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
        float4* gv = vel + chunk * kRainChunkVec;
        float4* gp = MODE == 2 ? pos + chunk * kRainChunkVec : nullptr;
        float4 v0 = __ldcs(gv + t), v1 = __ldcs(gv + t + 256), v2 = __ldcs(gv + t + 512);
        float4 p0, p1, p2;
        if (MODE == 2) {
            p0 = __ldcs(gp + t);
            p1 = __ldcs(gp + t + 256);
            p2 = __ldcs(gp + t + 512);
        }
        sv[t] = v0;
        sv[t + 256] = v1;
        sv[t + 512] = v2;
        if (MODE == 2) {
            sp[t] = p0;
            sp[t + 256] = p1;
            sp[t + 512] = p2;
        }
        __syncthreads();
        Vec3x4 v, p;
        v.a = sv[3 * t];
        v.b = sv[3 * t + 1];
        v.c = sv[3 * t + 2];
        if (MODE == 2) {
            p.a = sp[3 * t];
            p.b = sp[3 * t + 1];
            p.c = sp[3 * t + 2];
        }
        float o[12], q[12];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float x, y, z;
            get3(v, e, x, y, z);
            if (MODE == 1 || MODE == 2) rain_wind_1(x, y, z, wx, wy, wz, terminal);
            if (MODE == 0 || MODE == 2) rain_gravity_1(x, y, z, k_gravity, terminal);
            o[3 * e] = x;
            o[3 * e + 1] = y;
            o[3 * e + 2] = z;
            if (MODE == 2) {
                float px, py, pz;
                get3(p, e, px, py, pz);
                q[3 * e] = __fadd_rn(px, __fmul_rn(x, dt));
                q[3 * e + 1] = __fadd_rn(py, __fmul_rn(y, dt));
                q[3 * e + 2] = __fadd_rn(pz, __fmul_rn(z, dt));
            }
        }
        sv[3 * t] = make_float4(o[0], o[1], o[2], o[3]);
        sv[3 * t + 1] = make_float4(o[4], o[5], o[6], o[7]);
        sv[3 * t + 2] = make_float4(o[8], o[9], o[10], o[11]);
        if (MODE == 2) {
            sp[3 * t] = make_float4(q[0], q[1], q[2], q[3]);
            sp[3 * t + 1] = make_float4(q[4], q[5], q[6], q[7]);
            sp[3 * t + 2] = make_float4(q[8], q[9], q[10], q[11]);
        }
        __syncthreads();
        __stcs(gv + t, sv[t]);
        __stcs(gv + t + 256, sv[t + 256]);
        __stcs(gv + t + 512, sv[t + 512]);
        if (MODE == 2) {
            __stcs(gp + t, sp[t]);
            __stcs(gp + t + 256, sp[t + 256]);
            __stcs(gp + t + 512, sp[t + 512]);
        }
        __syncthreads();
    }
}


// TMA-pipelined form of the same systems (the default for whole 1024-entity chunks): a persistent CTA per SM slot,
// a 3-stage ring of {velocity chunk, position chunk} in shared memory, one producer thread issuing 1-D bulk loads
// (cp.async.bulk -> UBLKCP) against full barriers, eight consumer warps updating the chunk in place in shared
// memory, and one store thread writing it back with bulk stores (shared -> global) before it recycles the stage.
// Loads of chunks k+1 and k+2 and the store of chunk k-1 are in flight while chunk k is computed, with no
// __syncthreads and no register staging; 72 KiB of shared memory per CTA, two CTAs per SM.
constexpr int kRainStages = 3;
constexpr int kRainCtasPerSm = 3;  // L2-resident sizes: 1: 32.8 us, 2: 31.7 us, 3: 30.7 us per fused tick of 4 Mi drops
constexpr int kRainChunkBytes = kRainChunkVec * 16;  // 12 KiB per column per chunk
template <int MODE>
__global__ void __launch_bounds__(kStreamThreads + 64)
    rain_tma_kernel(float4* __restrict__ vel, float4* __restrict__ pos, uint64_t n_chunks, float wx, float wy, float wz,
                    float k_gravity, float terminal, float dt) {
    constexpr int kCols = MODE == 2 ? 2 : 1;
    extern __shared__ __align__(128) unsigned char rain_smem[];
    float4* ring = reinterpret_cast<float4*>(rain_smem);  // [stage][column][kRainChunkVec]
    __shared__ __align__(8) uint64_t full_bar[kRainStages];
    __shared__ __align__(8) uint64_t done_bar[kRainStages];
    __shared__ __align__(8) uint64_t empty_bar[kRainStages];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < kRainStages; ++s) {
            ptx::mbar_init(&full_bar[s], 1);
            ptx::mbar_init(&done_bar[s], kStreamThreads / 32);
            ptx::mbar_init(&empty_bar[s], 1);
        }
        ptx::mbar_fence_init();
    }
    __syncthreads();
    const uint64_t first = blockIdx.x;
    const uint64_t my_chunks = first < n_chunks ? (n_chunks - first + gridDim.x - 1) / gridDim.x : 0;

    if (warp == kStreamThreads / 32) {  // ---- load producer
        if (lane == 0) {
            for (uint64_t k = 0; k < my_chunks; ++k) {
                const uint32_t stage = k % kRainStages, phase = (k / kRainStages) & 1u;
                const uint64_t chunk = first + k * gridDim.x;
                ptx::mbar_wait(&empty_bar[stage], phase ^ 1u);
                float4* dst = ring + stage * kCols * kRainChunkVec;
                ptx::mbar_arrive_expect_tx(&full_bar[stage], kCols * kRainChunkBytes);
                ptx::bulk_g2s(dst, vel + chunk * kRainChunkVec, kRainChunkBytes, &full_bar[stage]);
                if (MODE == 2) ptx::bulk_g2s(dst + kRainChunkVec, pos + chunk * kRainChunkVec, kRainChunkBytes, &full_bar[stage]);
            }
        }
        return;
    }
    if (warp == kStreamThreads / 32 + 1) {  // ---- store thread
        if (lane == 0) {
            for (uint64_t k = 0; k < my_chunks; ++k) {
                const uint32_t stage = k % kRainStages, phase = (k / kRainStages) & 1u;
                const uint64_t chunk = first + k * gridDim.x;
                ptx::mbar_wait(&done_bar[stage], phase);
                const float4* src = ring + stage * kCols * kRainChunkVec;
                ptx::bulk_s2g(vel + chunk * kRainChunkVec, src, kRainChunkBytes);
                if (MODE == 2) ptx::bulk_s2g(pos + chunk * kRainChunkVec, src + kRainChunkVec, kRainChunkBytes);
                ptx::bulk_commit();
                ptx::bulk_wait_read0();
                ptx::mbar_arrive(&empty_bar[stage]);
            }
            ptx::bulk_wait0();
        }
        return;
    }
    // ---- consumers: thread t owns entities 4t..4t+3 of the chunk = float4 3t..3t+2 of each column
    const int t = threadIdx.x;
    for (uint64_t k = 0; k < my_chunks; ++k) {
        const uint32_t stage = k % kRainStages, phase = (k / kRainStages) & 1u;
        ptx::mbar_wait(&full_bar[stage], phase);
        float4* sv = ring + stage * kCols * kRainChunkVec;
        float4* sp = sv + kRainChunkVec;
        Vec3x4 v, p;
        v.a = sv[3 * t];
        v.b = sv[3 * t + 1];
        v.c = sv[3 * t + 2];
        if (MODE == 2) {
            p.a = sp[3 * t];
            p.b = sp[3 * t + 1];
            p.c = sp[3 * t + 2];
        }
        float o[12], q[12];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float x, y, z;
            get3(v, e, x, y, z);
            if (MODE == 1 || MODE == 2) rain_wind_1(x, y, z, wx, wy, wz, terminal);
            if (MODE == 0 || MODE == 2) rain_gravity_1(x, y, z, k_gravity, terminal);
            o[3 * e] = x;
            o[3 * e + 1] = y;
            o[3 * e + 2] = z;
            if (MODE == 2) {
                float px, py, pz;
                get3(p, e, px, py, pz);
                q[3 * e] = __fadd_rn(px, __fmul_rn(x, dt));
                q[3 * e + 1] = __fadd_rn(py, __fmul_rn(y, dt));
                q[3 * e + 2] = __fadd_rn(pz, __fmul_rn(z, dt));
            }
        }
        sv[3 * t] = make_float4(o[0], o[1], o[2], o[3]);
        sv[3 * t + 1] = make_float4(o[4], o[5], o[6], o[7]);
        sv[3 * t + 2] = make_float4(o[8], o[9], o[10], o[11]);
        if (MODE == 2) {
            sp[3 * t] = make_float4(q[0], q[1], q[2], q[3]);
            sp[3 * t + 1] = make_float4(q[4], q[5], q[6], q[7]);
            sp[3 * t + 2] = make_float4(q[8], q[9], q[10], q[11]);
        }
        ptx::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&done_bar[stage]);
    }
}

// Grid for a grid-stride streaming kernel: at most `cap` resident-ish CTAs, and every CTA gets the
// same number of chunks (a ragged last pass would leave most SMs idle for a whole chunk time).
inline uint32_t stream_grid(uint64_t work_items, uint32_t per_block) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const uint64_t chunks = (work_items + per_block - 1) / per_block;
    static const int cap_per_sm = getenv("RECS_GPU_STREAM_CTAS") ? atoi(getenv("RECS_GPU_STREAM_CTAS")) : 16;  // tuning knob
    const uint64_t cap = static_cast<uint64_t>(sms) * cap_per_sm;  // default: 8 resident CTAs of 256 threads x 2 waves
    if (chunks <= cap) return static_cast<uint32_t>(chunks ? chunks : 1);
    const uint64_t passes = (chunks + cap - 1) / cap;
    return static_cast<uint32_t>((chunks + passes - 1) / passes);
}

}  // namespace

cudaError_t launch_axpy_rn(float* y, const float* x, float a, uint64_t n, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    const uintptr_t ya = reinterpret_cast<uintptr_t>(y), xa = reinterpret_cast<uintptr_t>(x);
    if ((ya & 15) != (xa & 15) || (ya & 3)) {
        axpy_rn_scalar_kernel<<<stream_grid(n, kStreamThreads), kStreamThreads, 0, s>>>(y, x, a, n);
        return cudaGetLastError();
    }
    // scalar head up to 16-byte alignment, float4 body, scalar tail
    uint64_t head = ((16 - (ya & 15)) & 15) / 4;
    if (head > n) head = n;
    if (head) axpy_rn_scalar_kernel<<<1, 32, 0, s>>>(y, x, a, head);
    const uint64_t n4 = (n - head) / 4;
    if (n4) {
        // (a TMA-pipelined version of this kernel was measured and dropped: 60.4 us with one CTA per SM, 63.8 with
        // three, against 57.6 us for this LDG/STG kernel on 10 M entities; profiles/r01_rain_sizes.txt)
        const uint32_t grid = stream_grid(n4, kStreamThreads * kStreamUnroll);
        axpy_rn_vec_kernel<<<grid, kStreamThreads, 0, s>>>(reinterpret_cast<float4*>(y + head),
                                                          reinterpret_cast<const float4*>(x + head), a,
                                                          n4);
    }
    const uint64_t tail = n - head - 4 * n4;
    if (tail) axpy_rn_scalar_kernel<<<1, 32, 0, s>>>(y + (n - tail), x + (n - tail), a, tail);
    return cudaGetLastError();
}

cudaError_t launch_movement_pack(float* pos, const float* vel, const float* mass, float dt, float g,
                                 float* posm, uint64_t row_begin, uint64_t n, uint64_t dst_offset,
                                 cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    const uint64_t grid = (n + kStreamThreads - 1) / kStreamThreads;
    movement_pack_kernel<<<static_cast<uint32_t>(grid), kStreamThreads, 0, s>>>(
        pos, vel, mass, dt, g, posm, row_begin, n, dst_offset);
    return cudaGetLastError();
}

namespace {
template <int MODE>
cudaError_t launch_rain_impl(float* vel, float* pos, uint64_t n, float wx, float wy, float wz, float k_gravity,
                             float terminal, float dt, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    const bool aligned = (reinterpret_cast<uintptr_t>(vel) & 15) == 0 && (MODE != 2 || (reinterpret_cast<uintptr_t>(pos) & 15) == 0);
    const uint64_t n_chunks = aligned ? n / kRainChunk : 0;
    if (n_chunks) {
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        // Measured on 4 / 16 / 32 Mi drops (profiles/r01_rain_sizes.txt): the TMA pipeline with three persistent CTAs
        // per SM while the columns fit L2 (30 vs 35 us at 4 Mi drops), with ONE CTA per SM once the tick is HBM-bound
        // (6.5-6.6 TB/s back to back at 16-32 Mi drops against 5.9 with three CTAs and 6.15 for the LDG/STS staged
        // kernel): fewer concurrent chunk streams keep DRAM pages open.  RECS_GPU_RAIN_STAGED forces the staged kernel.
        static const bool staged = getenv("RECS_GPU_RAIN_STAGED") != nullptr;
        const uint64_t footprint = n * 12ull * (MODE == 2 ? 2 : 1);
        if (staged) {
            const uint64_t cap = static_cast<uint64_t>(sms) * 8;  // resident CTAs: 8 x 256 threads per SM
            const uint64_t passes = (n_chunks + cap - 1) / cap;   // same number of chunks for every CTA
            const uint32_t grid = static_cast<uint32_t>((n_chunks + passes - 1) / passes);
            rain_staged_kernel<MODE><<<grid, kStreamThreads, 0, s>>>(reinterpret_cast<float4*>(vel),
                                                                     reinterpret_cast<float4*>(pos), n_chunks, wx, wy, wz,
                                                                     k_gravity, terminal, dt);
        } else {
            constexpr int smem = kRainStages * (MODE == 2 ? 2 : 1) * kRainChunkBytes;
            // per device and cheap: set on every launch rather than caching it per process
            cudaFuncSetAttribute(rain_tma_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            static const int forced_ctas = getenv("RECS_GPU_RAIN_CTAS") ? atoi(getenv("RECS_GPU_RAIN_CTAS")) : 0;
            const int ctas_per_sm = forced_ctas > 0 ? forced_ctas : (footprint > kStreamL2Footprint ? 1 : kRainCtasPerSm);
            const uint64_t cap = static_cast<uint64_t>(sms) * ctas_per_sm;
            const uint32_t grid = static_cast<uint32_t>(n_chunks < cap ? n_chunks : cap);
            rain_tma_kernel<MODE><<<grid, kStreamThreads + 64, smem, s>>>(reinterpret_cast<float4*>(vel),
                                                                          reinterpret_cast<float4*>(pos), n_chunks, wx, wy,
                                                                          wz, k_gravity, terminal, dt);
        }
    }
    const uint64_t done = n_chunks * kRainChunk;
    if (done < n) {  // tail (< 1024 entities, or everything when unaligned)
        const uint64_t rest = n - done;
        rain_kernel<MODE><<<stream_grid(rest / 4 + 1, kStreamThreads), kStreamThreads, 0, s>>>(
            vel + 3 * done, pos ? pos + 3 * done : nullptr, rest, wx, wy, wz, k_gravity, terminal, dt);
    }
    return cudaGetLastError();
}
}  // namespace

cudaError_t launch_rain_gravity(float* vel, uint64_t n, float gravity_dt, float terminal, cudaStream_t s) {
    return launch_rain_impl<0>(vel, nullptr, n, 0.f, 0.f, 0.f, gravity_dt, terminal, 0.f, s);
}
cudaError_t launch_rain_wind(float* vel, uint64_t n, float wx, float wy, float wz, float terminal, cudaStream_t s) {
    return launch_rain_impl<1>(vel, nullptr, n, wx, wy, wz, 0.f, terminal, 0.f, s);
}
cudaError_t launch_rain_fused(float* vel, float* pos, uint64_t n, float wx, float wy, float wz, float gravity_dt,
                              float terminal, float dt, cudaStream_t s) {
    return launch_rain_impl<2>(vel, pos, n, wx, wy, wz, gravity_dt, terminal, dt, s);
}

// ---- measurement ---------------------------------------------------------------------------------
namespace {
template <int PACKED>
__global__ void __launch_bounds__(256) fp32_peak_kernel(uint32_t iters, float* sink,
                                                        unsigned long long* cycles) {
    // 16 independent accumulator chains per thread, all operands in registers.
    float seed = 1.0f + 1e-7f * threadIdx.x;
    const long long t0 = clock64();
    if (PACKED) {
        unsigned long long acc[12];
#pragma unroll
        for (int i = 0; i < 12; ++i)
            asm("mov.b64 %0, {%1, %2};" : "=l"(acc[i]) : "f"(seed + i), "f"(seed - i));
        unsigned long long m, c;
        asm("mov.b64 %0, {%1, %2};" : "=l"(m) : "f"(0.999999f), "f"(1.000001f));
        asm("mov.b64 %0, {%1, %2};" : "=l"(c) : "f"(1e-6f), "f"(-1e-6f));
        for (uint32_t k = 0; k < iters; ++k) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int i = 0; i < 12; ++i)
                    asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[i]) : "l"(m), "l"(c));
        }
        float r = 0.f;
#pragma unroll
        for (int i = 0; i < 12; ++i) {
            float lo, hi;
            asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[i]));
            r += lo + hi;
        }
        seed = r;
    } else {
        float acc[24];
#pragma unroll
        for (int i = 0; i < 24; ++i) acc[i] = seed + i;
        const float m = 0.999999f, c = 1e-6f;
        for (uint32_t k = 0; k < iters; ++k) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int i = 0; i < 24; ++i)
                    asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(acc[i]) : "f"(m), "f"(c));
        }
        float r = 0.f;
#pragma unroll
        for (int i = 0; i < 24; ++i) r += acc[i];
        seed = r;
    }
    const long long t1 = clock64();
    if (seed == 123.456f) sink[0] = seed;  // keep the chains alive
    if (threadIdx.x == 0) cycles[blockIdx.x] = static_cast<unsigned long long>(t1 - t0);
}

__global__ void fill_kernel(float* p, uint64_t n, float v) {
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < n; i += stride)
        p[i] = v;
}
// Streams n floats through L2 without leaving dirty lines behind (the sum is stored only if it is NaN,
// which a zero-filled buffer never produces).
__global__ void read_sweep_kernel(const float4* __restrict__ p, uint64_t n4, float* sink) {
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    float acc = 0.f;
    for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < n4; i += stride) {
        const float4 v = __ldcg(p + i);
        acc += (v.x + v.y) + (v.z + v.w);
    }
    if (acc != acc) *sink = acc;
}
}  // namespace

cudaError_t launch_read_sweep(const float* p, uint64_t n, float* sink, cudaStream_t s) {
    if (n < 4) return cudaSuccess;
    read_sweep_kernel<<<stream_grid(n / 4, 256 * 4), 256, 0, s>>>(reinterpret_cast<const float4*>(p), n / 4, sink);
    return cudaGetLastError();
}

cudaError_t launch_fp32_peak(int packed, uint32_t iters, uint32_t grid, float* sink,
                             unsigned long long* cycles, cudaStream_t s) {
    if (packed)
        fp32_peak_kernel<1><<<grid, 256, 0, s>>>(iters, sink, cycles);
    else
        fp32_peak_kernel<0><<<grid, 256, 0, s>>>(iters, sink, cycles);
    return cudaGetLastError();
}

cudaError_t launch_fill(float* p, uint64_t n, float v, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    fill_kernel<<<stream_grid(n, 256 * 4), 256, 0, s>>>(p, n, v);
    return cudaGetLastError();
}

}  // namespace recs
