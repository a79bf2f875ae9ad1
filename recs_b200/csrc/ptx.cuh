// Thin inline-PTX layer for sm_100a: packed FP32x2 arithmetic (FFMA2/FADD2/FMUL2 in SASS),
// mbarrier and 1-D TMA bulk copies (UBLKCP in SASS).  Nothing here is portable to other
// architectures on purpose.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace recs {
namespace ptx {

typedef unsigned long long f32x2;  // two f32 lanes in one 64-bit register pair: {lo, hi}

__device__ __forceinline__ f32x2 pack(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack(f32x2 v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
// a + b with the guarantee that the add stays an add.  ptxas (12.9) contracts `mul.rn.f32x2` + `add.rn.f32x2` into
// FFMA2 despite the explicit .rn and regardless of -fmad=false (it even rewrites fma(a, b, -0) and fma(a, 1, b) to get
// there), so where a product must be rounded on its own the sum is formed with two scalar add.rn.f32, which it
// leaves alone (checked in SASS: FMUL2 x3 + FADD x4 for the reference's (dx*dx + dy*dy) + dz*dz).
__device__ __forceinline__ f32x2 add2_unfused(f32x2 a, f32x2 b) {
    float a0, a1, b0, b1;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a0), "=f"(a1) : "l"(a));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(b0), "=f"(b1) : "l"(b));
    return pack(__fadd_rn(a0, b0), __fadd_rn(a1, b1));
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
// MUFU.RSQ without the denormal fix-up sequence rsqrtf() expands to (a flushed denormal d2 is
// below the reference's epsilon and is masked by the caller anyway).
__device__ __forceinline__ float rsqrt_approx(float x) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier ------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}

// ---- TMA 1-D bulk copy global -> shared, completion on an mbarrier ---------------------------
// size must be a multiple of 16 bytes, both addresses 16-byte aligned.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                         uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// ---- TMA 1-D bulk copy shared -> global (bulk async-group completion) ----------------------------
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst),
                 "r"(smem_u32(smem_src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// Returns when the shared-memory SOURCES of all committed bulk groups have been read (the stage may be refilled).
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// Returns when all committed bulk groups are complete (global writes performed).
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// Makes this thread's generic-proxy shared-memory writes visible to the async proxy (TMA) that will read them.
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}


// ---- cross-GPU flags (peer-mapped memory over NVLink) ------------------------------------------------
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// Orders earlier generic-proxy observations (the acquire of a flag) before later async-proxy (TMA) reads of global memory.
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

// Spins until *flag has reached generation `gen` (wrap-safe).  Bounded: after `timeout_ns` it raises *error and
// returns false, so a missing peer ends in a reported error instead of a hung GPU.
__device__ __forceinline__ bool wait_generation(const uint32_t* flag, uint32_t gen, uint32_t* error,
                                                unsigned long long timeout_ns = 30000000000ull) {
    if (static_cast<int32_t>(ld_acquire_sys(flag) - gen) >= 0) return true;
    const unsigned long long t0 = globaltimer_ns();
    uint32_t spins = 0;
    while (static_cast<int32_t>(ld_acquire_sys(flag) - gen) < 0) {
        __nanosleep(64);
        if ((++spins & 1023u) == 0 && globaltimer_ns() - t0 > timeout_ns) {
            if (error) *reinterpret_cast<volatile uint32_t*>(error) = 1u;
            return false;
        }
    }
    return true;
}

// ---- grid-wide barrier of a cooperative launch (one GPU) ---------------------------------------------------------
// One thread per CTA: arrive (release, gpu scope: this CTA's earlier stores, ordered before by the CTA barrier, become
// visible with it) and spin until `target` arrivals have been counted since the launch.  Tight spin -- the wait is a
// microsecond -- but bounded like wait_generation.
__device__ __forceinline__ bool grid_barrier(uint32_t* counter, uint32_t target, uint32_t* error,
                                             unsigned long long timeout_ns = 2000000000ull) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
    uint32_t v, spins = 0;
    unsigned long long t0 = 0;
    for (;;) {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
        if (static_cast<int32_t>(v - target) >= 0) return true;
        if ((++spins & 4095u) == 0) {
            if (t0 == 0) {
                t0 = globaltimer_ns();
            } else if (globaltimer_ns() - t0 > timeout_ns) {
                if (error) *reinterpret_cast<volatile uint32_t*>(error) = 1u;
                return false;
            }
        }
    }
}

}  // namespace ptx
}  // namespace recs
