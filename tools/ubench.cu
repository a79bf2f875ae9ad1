// Issue-slot / pipe microbenchmarks for sm_100a (not product code).
// Each kernel runs ITER iterations of a fully unrolled body on every thread; we report
// SM cycles per body per warp-scheduler (SMSP), i.e. elapsed_cycles / (ITER * bodies_per_iter * warps_per_smsp).
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;

#define FMA2(acc, m, c) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc) : "l"(m), "l"(c))
#define FMA2D(acc, b, c) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(b), "l"(c))
#define FMA1(acc, m, c) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(acc) : "f"(m), "f"(c))
#define ALU(x, y) asm volatile("lop3.b32 %0, %0, %1, %1, 0x96;" : "+r"(x) : "r"(y))
#define FSELOP(x, y, p) asm volatile("{.reg .pred q; setp.ne.u32 q, %2, 0; selp.f32 %0, %0, %1, q;}" : "+f"(x) : "f"(y), "r"(p))
#define MUFU(x) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(x))
#define FMNMX3(x, y, z) asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(x) : "f"(y), "f"(z))
#define MASKSEL(r, d, eps) asm volatile("{.reg .pred q; setp.gt.f32 q, %1, %2; selp.f32 %0, %0, 0f00000000, q;}" : "+f"(r) : "f"(d), "f"(eps))
#define FMNMX(x, y) asm volatile("max.f32 %0, %0, %1;" : "+f"(x) : "f"(y))

template <int MODE>
__global__ void __launch_bounds__(256, 4) k(int iters, float* sink, u64* cyc) {
    __shared__ float4 sm[512];
    for (int i = threadIdx.x; i < 512; i += blockDim.x) sm[i] = make_float4(1.f + i, 2.f, 3.f, 4.f);
    __syncthreads();
    float4 l0 = make_float4(0, 0, 0, 0), l1 = l0;
    u64 a[8];
    float f[8];
    unsigned x[8];
    for (int i = 0; i < 8; ++i) {
        float lo = 1.0f + threadIdx.x * 1e-6f + i, hi = 2.0f + i;
        asm("mov.b64 %0, {%1,%2};" : "=l"(a[i]) : "f"(lo), "f"(hi));
        f[i] = 1.5f + i + threadIdx.x;
        x[i] = threadIdx.x * 7 + i;
    }
    u64 m, c;
    asm("mov.b64 %0, {%1,%2};" : "=l"(m) : "f"(0.999999f), "f"(1.000001f));
    asm("mov.b64 %0, {%1,%2};" : "=l"(c) : "f"(1e-6f), "f"(-1e-6f));
    float fm = 0.99999f, fc = 1e-6f;
    unsigned y = blockIdx.x + 12345;
    const unsigned smbase = (unsigned)__cvta_generic_to_shared(sm);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (MODE == 0) {  // 8 independent FFMA2
#pragma unroll
                for (int i = 0; i < 8; ++i) FMA2(a[i], m, c);
            } else if (MODE == 1) {  // 8 x (FFMA2 + ALU)
#pragma unroll
                for (int i = 0; i < 8; ++i) { FMA2(a[i], m, c); ALU(x[i], y); }
            } else if (MODE == 2) {  // 8 x (FFMA2 + 2 ALU)
#pragma unroll
                for (int i = 0; i < 8; ++i) { FMA2(a[i], m, c); ALU(x[i], y); ALU(x[(i + 4) & 7], y); }
            } else if (MODE == 3) {  // 8 scalar FFMA
#pragma unroll
                for (int i = 0; i < 8; ++i) FMA1(f[i], fm, fc);
            } else if (MODE == 4) {  // 8 x (FFMA + ALU)
#pragma unroll
                for (int i = 0; i < 8; ++i) { FMA1(f[i], fm, fc); ALU(x[i], y); }
            } else if (MODE == 5) {  // 6 FFMA2 + 1 MUFU (x8 -> body counts 48 FFMA2 + 8 MUFU)
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    FMA2(a[i], m, c); FMA2(a[(i + 1) & 7], m, c); FMA2(a[(i + 2) & 7], m, c);
                    FMA2(a[(i + 3) & 7], m, c); FMA2(a[(i + 4) & 7], m, c); FMA2(a[(i + 5) & 7], m, c);
                    MUFU(f[i]);
                }
            } else if (MODE == 6) {  // dependent FFMA2 chain (latency): 8 serial on one accumulator
#pragma unroll
                for (int i = 0; i < 8; ++i) FMA2(a[0], m, c);
            } else if (MODE == 7) {  // dependent scalar FFMA chain
#pragma unroll
                for (int i = 0; i < 8; ++i) FMA1(f[0], fm, fc);
            } else if (MODE == 8) {  // FFMA2 with 3 distinct register operands, no reuse (acc += b*c, b/c vary)
#pragma unroll
                for (int i = 0; i < 8; ++i) FMA2D(a[i], a[(i + 3) & 7], a[(i + 5) & 7]);
            } else if (MODE == 9) {  // gravity-like mix per pair-interaction: 12 FFMA2 + 2 MUFU + 4 ALU
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    for (int j = 0; j < 12; ++j) FMA2(a[(i * 4 + j) & 7], m, c);
                    MUFU(f[i]); MUFU(f[i + 2]);
                    ALU(x[i], y); ALU(x[i + 2], y); ALU(x[i + 4], y); ALU(x[i + 6], y);
                }
            } else if (MODE == 10) {  // 12 FFMA2 + 4 ALU (no MUFU)
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    for (int j = 0; j < 12; ++j) FMA2(a[(i * 4 + j) & 7], m, c);
                    ALU(x[i], y); ALU(x[i + 2], y); ALU(x[i + 4], y); ALU(x[i + 6], y);
                }
            } else if (MODE == 11) {  // 12 FFMA2 + 2 MUFU
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    for (int j = 0; j < 12; ++j) FMA2(a[(i * 4 + j) & 7], m, c);
                    MUFU(f[i]); MUFU(f[i + 2]);
                }
            } else if (MODE == 13) {
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    for (int j = 0; j < 12; ++j) FMA2(a[(i * 4 + j) & 7], m, c);
                    MUFU(f[i]); MUFU(f[i + 2]);
                    FMNMX(f[i + 4], fm); FMNMX(f[i + 6], fc); FMNMX(f[i + 4], fc); FMNMX(f[i + 6], fm);
                }
            } else if (MODE == 14) {  // exact mask: 12 FFMA2 + 2 MUFU + 2 (FSETP+FSEL)
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    for (int j = 0; j < 12; ++j) FMA2(a[(i * 4 + j) & 7], m, c);
                    MUFU(f[i]); MUFU(f[i + 2]);
                    MASKSEL(f[i], f[i + 4], fc); MASKSEL(f[i + 2], f[i + 6], fc);
                }
            } else if (MODE == 15) {  // detector: 12 FFMA2 + 2 MUFU + 1 FMNMX3
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    for (int j = 0; j < 12; ++j) FMA2(a[(i * 4 + j) & 7], m, c);
                    MUFU(f[i]); MUFU(f[i + 2]);
                    FMNMX3(f[7], f[i + 4], f[i + 5]);
                }
            } else if (MODE == 16) {  // detector with two 2-input mins
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    for (int j = 0; j < 12; ++j) FMA2(a[(i * 4 + j) & 7], m, c);
                    MUFU(f[i]); MUFU(f[i + 2]);
                    FMNMX(f[7], f[i + 4]); FMNMX(f[6], f[i + 5]);
                }
            } else if (MODE == 17) {  // no mask at all: 12 FFMA2 + 2 MUFU
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    for (int j = 0; j < 12; ++j) FMA2(a[(i * 4 + j) & 7], m, c);
                    MUFU(f[i]); MUFU(f[i + 2]);
                }
            } else if (MODE == 18) {  // 12 FFMA2 + 2 MUFU + 1 FMNMX + 1 LDS.128 (uniform address)
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    for (int j = 0; j < 12; ++j) FMA2(a[(i * 4 + j) & 7], m, c);
                    MUFU(f[i]); MUFU(f[i + 2]);
                    FMNMX(f[7], f[i + 4]);
                    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(l0.x), "=f"(l0.y), "=f"(l0.z), "=f"(l0.w) : "r"(smbase + ((it * 8 + u * 2 + i) & 255) * 16));
                    f[4] += l0.x;
                }
            } else if (MODE == 19) {  // same with 2 LDS.128 per body
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    for (int j = 0; j < 12; ++j) FMA2(a[(i * 4 + j) & 7], m, c);
                    MUFU(f[i]); MUFU(f[i + 2]);
                    FMNMX(f[7], f[i + 4]);
                    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(l0.x), "=f"(l0.y), "=f"(l0.z), "=f"(l0.w) : "r"(smbase + ((it * 8 + u * 2 + i) & 255) * 16));
                    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(l1.x), "=f"(l1.y), "=f"(l1.z), "=f"(l1.w) : "r"(smbase + 4096 + ((it * 8 + u * 2 + i) & 255) * 16));
                    f[4] += l0.x; f[5] += l1.y;
                }
            } else if (MODE == 20) {  // 12 FFMA2 + 2 MUFU + 1 FMNMX, no LDS (reference for 18/19, same extra FADDs)
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    for (int j = 0; j < 12; ++j) FMA2(a[(i * 4 + j) & 7], m, c);
                    MUFU(f[i]); MUFU(f[i + 2]);
                    FMNMX(f[7], f[i + 4]);
                    f[4] += l0.x;
                }
            } else if (MODE == 12) {  // 24 scalar FFMA + 2 MUFU + 4 ALU (scalar formulation of 2 interactions)
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    for (int j = 0; j < 24; ++j) FMA1(f[(j & 3) + 4], fm, fc);
                    MUFU(f[i]); MUFU(f[i + 2]);
                    ALU(x[i], y); ALU(x[i + 2], y); ALU(x[i + 4], y); ALU(x[i + 6], y);
                }
            }
        }
    }
    long long t1 = clock64();
    float r = 0;
    for (int i = 0; i < 8; ++i) {
        float lo, hi;
        asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a[i]));
        r += lo + hi + f[i] + x[i] + l0.y + l1.z;
    }
    if (r == 12345.678f) sink[0] = r;
    if (threadIdx.x == 0) cyc[blockIdx.x] = (u64)(t1 - t0);
}


static double g_clk_hz = 0;

template <int MODE>
double time_kernel(int grid, int block, int iters, u64* max_cycles) {
    float* sink;
    u64* cyc;
    cudaMalloc(&sink, 4);
    cudaMalloc(&cyc, grid * 8);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    k<MODE><<<grid, block>>>(iters, sink, cyc);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(a);
        k<MODE><<<grid, block>>>(iters, sink, cyc);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    if (max_cycles) {
        u64* h = new u64[grid];
        cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost);
        u64 m = 0;
        for (int i = 0; i < grid; ++i) m = h[i] > m ? h[i] : m;
        *max_cycles = m;
        delete[] h;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(e));
    cudaFree(sink);
    cudaFree(cyc);
    return best * 1e-3;
}

template <int MODE>
void run(const char* name, int insts_per_u, int warps_per_smsp) {
    const int iters = 20000;
    const int ctas_per_sm = warps_per_smsp * 4 / 8;  // 256-thread CTAs, one wave
    const int grid = 148 * ctas_per_sm;
    const double sec = time_kernel<MODE>(grid, 256, iters, nullptr);
    const double cycles = sec * g_clk_hz;
    const double per_u = cycles / (iters * 4.0) / warps_per_smsp;  // SMSP cycles per body per warp
    printf("%-40s w/SMSP=%d  %7.2f cyc/body  (%2d insts -> %.2f cyc/inst)\n", name, warps_per_smsp, per_u,
           insts_per_u, per_u / insts_per_u);
}

int main() {
    {   // calibrate the SM clock: single-wave kernel, max per-CTA clock64 delta over event time
        u64 cyc = 0;
        const double sec = time_kernel<3>(148 * 4, 256, 20000, &cyc);
        g_clk_hz = cyc / sec;
        printf("SM clock ~ %.0f MHz (kernel %.3f ms)\n", g_clk_hz / 1e6, sec * 1e3);
    }
    for (int w : {4, 8}) {
        run<0>("8 FFMA2", 8, w);
        run<1>("8 x (FFMA2 + ALU)", 16, w);
        run<2>("8 x (FFMA2 + 2 ALU)", 24, w);
        run<3>("8 FFMA", 8, w);
        run<4>("8 x (FFMA + ALU)", 16, w);
        run<5>("8 x (6 FFMA2 + MUFU)", 56, w);
        run<8>("8 FFMA2 3-distinct-regs", 8, w);
        run<9>("2 x (12 FFMA2 + 2 MUFU + 4 ALU)", 36, w);
        run<10>("2 x (12 FFMA2 + 4 ALU)", 32, w);
        run<11>("2 x (12 FFMA2 + 2 MUFU)", 28, w);
        run<12>("2 x (24 FFMA + 2 MUFU + 4 ALU)", 60, w);
        run<13>("2 x (12 FFMA2 + 2 MUFU + 4 FSEL-like)", 36, w);
        run<14>("2 x (12 FFMA2 + 2 MUFU + 2 FSETP + 2 FSEL)", 36, w);
        run<15>("2 x (12 FFMA2 + 2 MUFU + 1 FMNMX3)", 30, w);
        run<16>("2 x (12 FFMA2 + 2 MUFU + 2 FMNMX)", 32, w);
        run<17>("2 x (12 FFMA2 + 2 MUFU)", 28, w);
        run<20>("2 x (12 FFMA2 + 2 MUFU + FMNMX + FADD)", 32, w);
        run<18>("2 x (.. + 1 LDS.128)", 34, w);
        run<19>("2 x (.. + 2 LDS.128)", 36, w);
    }
    {   // latency: one warp per SM, dependent chain
        const int iters = 20000;
        double s6 = time_kernel<6>(148, 32, iters, nullptr), s7 = time_kernel<7>(148, 32, iters, nullptr);
        printf("dependent FFMA2 latency ~ %.2f cycles, dependent FFMA latency ~ %.2f cycles\n",
               s6 * g_clk_hz / (iters * 32.0), s7 * g_clk_hz / (iters * 32.0));
    }
    return 0;
}
