// All-pairs gravity for sm_100a -- the device body of recs' `gravity` system
// (crates/n-body/src/lib.rs:104-135):
//
//   for each body i:  acc_i = sum_j  to_body * ((G*m_j / d2) / sqrt(d2)),   to_body = p_j - p_i,
//                     d2 = |to_body|^2, pairs with d2 <= f32::EPSILON contribute zero (this is
//                     what removes j == i).
//
// Design (B200-first, not a translation of the scalar Rust loop; DESIGN.md 3.1 has the measurements):
//   * FP32-FMA-pipe bound, no tensor cores (not a contraction).  The whole interaction is packed FP32
//     (FADD2/FMUL2/FFMA2): 12 FMA-pipe instructions per TWO interactions.  In the default "i-packed" form the
//     two f32x2 lanes are two bodies i of the thread and body j is a broadcast scalar operand; the older
//     "j-packed" form (lanes = two bodies j) is kept as an A/B variant.
//   * the reference's `d2 <= f32::EPSILON => zero` costs no instruction per pair: positions are stored
//     scaled by 2^-32 (exact), so r^-3 overflows to +inf for every pair the reference masks (and a few
//     more); one finiteness test per row and tile sends that row to an exact masked redo of the tile.
//   * j tiles live in HBM pair-interleaved ({x0 x1 y0 y1 | z0 z1 gm0 gm1}, 32 B per pair), so a tile is one
//     contiguous 4 KiB block: a single elected thread of a producer warp moves it with a 1-D TMA bulk copy
//     (cp.async.bulk -> UBLKCP) into a 4-stage shared-memory ring guarded by full/empty mbarriers.
//     Consumers read tiles with uniform (broadcast) LDS.128.
//   * register blocking: 4 bodies i per thread (2 row pairs), so one LDS.128 pair feeds 4 pair-interactions;
//     a warp owns 128 consecutive rows, i.e. its own bodies sit in one j tile, which goes straight to the exact path.
//     Which rows take a tile's exact result is decided row by row from data alone (own tile, or the row's fast sum
//     poisoned), never by which rows happen to share a thread.
//   * stream-K work split over a FIXED SUMMATION TREE (kernels.cuh): the (i-block, j-tile) units of a phase are cut
//     into equal contiguous spans, one per persistent CTA (grid = SMs x resident CTAs), so 65 536 bodies load all 148
//     SMs as evenly as 1 M bodies do.  Tile sums are added into chunk sums (K consecutive tiles) and chunk sums into
//     the row total, both in ascending order from zero: a CTA that enters a chunk at its first tile stores the chunk's
//     running prefix once, a CTA whose span begins inside a chunk stores the remaining tile sums one by one, and the
//     reduce kernel continues the prefix with them in tile order -- deterministic, no float atomics, and not a bit of
//     a row depends on the grid, on the launch split or on the number of GPUs.  The reduce kernel then integrates.
//   * phases: a tile range (a rank of a sharded run walks its own slice first, then the other ranks' slices, whose
//     tiles it loads only after acquiring that rank's generation flag -- the peers push their refreshed positions
//     into this GPU's memory over NVLink) times an i-block range (a partly filled last i-block is dealt by itself, its
//     warps without rows skip the arithmetic).  One persistent launch runs all phases.
#include <algorithm>

#include "kernels.cuh"
#include "ptx.cuh"

namespace recs {

using ptx::f32x2;

namespace {

// Bits of the kernel's ABL template parameter (operation order / detector choice / timing-only ablations).
constexpr int kAblNoDetector = 1;   // TIMING ONLY: no min-d2 tracking in the unmasked loop
constexpr int kAblNoMufu = 4;       // TIMING ONLY: d2 stands in for rsqrt(d2)
constexpr int kAblCubeFirst = 8;    // s = (r^-2 * r^-1) * G*m instead of (G*m * r^-1) * r^-2
constexpr int kAblOverflow = 16;    // range-shifted data, overflow of r^-3 is the detector (implies cube first)

struct Span {
    uint32_t begin;
    uint32_t count;
};

// Unit span of CTA c when `units` are dealt to `grid` CTAs as evenly as possible.
__host__ __device__ inline Span span_of(uint32_t c, uint32_t grid, uint32_t units) {
    const uint32_t q = units / grid, rem = units % grid;
    Span s;
    s.begin = c * q + (c < rem ? c : rem);
    s.count = q + (c < rem ? 1u : 0u);
    return s;
}
// Inverse: which CTA owns unit u.
__host__ __device__ inline uint32_t cta_of(uint32_t u, uint32_t grid, uint32_t units) {
    const uint32_t q = units / grid, rem = units % grid;
    const uint32_t big = rem * (q + 1);
    if (u < big) return u / (q + 1);
    return rem + (u - big) / q;
}

// One pair-interaction: two bodies j (the f32x2 lanes) against one body i.
//   MASKED  = true : the reference's `if distance_squared <= f32::EPSILON { zero }` applied exactly
//                    (compare + select per lane; also turns rsqrt(0) = inf of the self pair into 0).
//   MASKED  = false: no mask; instead the smallest d2 seen is tracked (two FMNMX) so the caller can
//                    detect that this tile needed the mask and redo it exactly.
template <bool MASKED>
__device__ __forceinline__ void pair_interaction(const f32x2 xj, const f32x2 yj, const f32x2 zj, const f32x2 gm,
                                                 const float nx, const float ny, const float nz, const float eps,
                                                 f32x2& ax, f32x2& ay, f32x2& az, float& dmin_a, float& dmin_b) {  // (both: this row)
    // (x_j0, x_j1) - x_i: written as an add of the negated scalar so ptxas can use FADD2's
    // broadcast-scalar operand form (`-R.F32`): one register read instead of a materialised pair
    const f32x2 dx = ptx::add2(xj, ptx::pack(-nx, -nx));
    const f32x2 dy = ptx::add2(yj, ptx::pack(-ny, -ny));
    const f32x2 dz = ptx::add2(zj, ptx::pack(-nz, -nz));
    f32x2 d2 = ptx::mul2(dx, dx);
    if (MASKED) {
        // the mask decision must be the reference's: cgmath's magnitude2 = (dx*dx + dy*dy) + dz*dz with every
        // product and sum rounded on its own (rustc never contracts), crates/n-body/src/lib.rs:116-120
        d2 = ptx::add2_unfused(d2, ptx::mul2(dy, dy));
        d2 = ptx::add2_unfused(d2, ptx::mul2(dz, dz));
    } else {
        d2 = ptx::fma2(dy, dy, d2);
        d2 = ptx::fma2(dz, dz, d2);
    }
    float d2a, d2b;
    ptx::unpack(d2, d2a, d2b);
    float ra = ptx::rsqrt_approx(d2a);
    float rb = ptx::rsqrt_approx(d2b);
    if (MASKED) {
        ra = d2a > eps ? ra : 0.f;
        rb = d2b > eps ? rb : 0.f;
    } else {
        dmin_a = fminf(dmin_a, d2a);
        dmin_b = fminf(dmin_b, d2b);
    }
    const f32x2 rinv = ptx::pack(ra, rb);
    const f32x2 rinv2 = ptx::mul2(rinv, rinv);
    f32x2 s = ptx::mul2(gm, rinv);
    s = ptx::mul2(s, rinv2);  // G*m / d2 / sqrt(d2)
    ax = ptx::fma2(s, dx, ax);
    ay = ptx::fma2(s, dy, ay);
    az = ptx::fma2(s, dz, az);
}

template <int RI, int UNROLL, bool MASKED>
__device__ __forceinline__ void tile_loop(const float4* __restrict__ t, const float (&nxi)[RI], const float (&nyi)[RI],
                                          const float (&nzi)[RI], const float eps, f32x2 (&ax)[RI], f32x2 (&ay)[RI],
                                          f32x2 (&az)[RI], float (&dmin)[RI]) {
#pragma unroll UNROLL
    for (int p = 0; p < kGravTile / 2; ++p) {
        const float4 lo = t[2 * p];      // x0 x1 y0 y1
        const float4 hi = t[2 * p + 1];  // z0 z1 gm0 gm1
        const f32x2 xj = ptx::pack(lo.x, lo.y);
        const f32x2 yj = ptx::pack(lo.z, lo.w);
        const f32x2 zj = ptx::pack(hi.x, hi.y);
        const f32x2 gm = ptx::pack(hi.z, hi.w);
#pragma unroll
        for (int r = 0; r < RI; ++r)
            pair_interaction<MASKED>(xj, yj, zj, gm, nxi[r], nyi[r], nzi[r], eps, ax[r], ay[r], az[r], dmin[r], dmin[r]);
    }
}

// The same interaction with the roles of the f32x2 lanes swapped ("i-packed"): the lanes are two bodies i
// of this thread, body j is a scalar that every packed instruction takes in the broadcast-operand form.
// One LDS.128 pair (two bodies j) then feeds 2 x RI/2 pair-interactions, i.e. per pair-interaction half
// the shared-memory reads of the j-packed form at the same number of accumulator registers.
// nx/ny/nz hold the NEGATED positions of the two rows.
template <bool MASKED, int ABL = 0>
__device__ __forceinline__ void ipair_interaction(const float xj, const float yj, const float zj, const float gm,
                                                  const f32x2 nx, const f32x2 ny, const f32x2 nz, const float eps,
                                                  f32x2& ax, f32x2& ay, f32x2& az, float& dmin_a, float& dmin_b) {  // (both: this row)
    const f32x2 dx = ptx::add2(nx, ptx::pack(xj, xj));
    const f32x2 dy = ptx::add2(ny, ptx::pack(yj, yj));
    const f32x2 dz = ptx::add2(nz, ptx::pack(zj, zj));
    f32x2 d2 = ptx::mul2(dx, dx);
    if (MASKED) {  // unfused: the reference's own d2, so `d2 <= epsilon` decides exactly as it does (see pair_interaction)
        d2 = ptx::add2_unfused(d2, ptx::mul2(dy, dy));
        d2 = ptx::add2_unfused(d2, ptx::mul2(dz, dz));
    } else {
        d2 = ptx::fma2(dy, dy, d2);
        d2 = ptx::fma2(dz, dz, d2);
    }
    float d2a, d2b;
    ptx::unpack(d2, d2a, d2b);
    float ra, rb;
    if (ABL & kAblNoMufu) {
        ra = d2a;
        rb = d2b;
    } else {
        ra = ptx::rsqrt_approx(d2a);
        rb = ptx::rsqrt_approx(d2b);
    }
    if (MASKED) {
        ra = d2a > eps ? ra : 0.f;
        rb = d2b > eps ? rb : 0.f;
    } else if (!(ABL & (kAblNoDetector | kAblOverflow))) {
        dmin_a = fminf(dmin_a, d2a);
        dmin_b = fminf(dmin_b, d2b);
    }
    const f32x2 rinv = ptx::pack(ra, rb);
    const f32x2 rinv2 = ptx::mul2(rinv, rinv);
    f32x2 s;
    if (ABL & (kAblCubeFirst | kAblOverflow)) {
        // r^-3 first, mass last: on range-shifted data r^-3 overflows for every pair the reference masks, whatever
        // its mass.  (Unshifted, r^-3 is denormal beyond 2e12 m: the exact path keeps the (G*m*r^-1)*r^-2 order.)
        s = ptx::mul2(rinv2, rinv);
        s = ptx::mul2(s, ptx::pack(gm, gm));
    } else {
        s = ptx::mul2(rinv, ptx::pack(gm, gm));
        s = ptx::mul2(s, rinv2);
    }
    ax = ptx::fma2(s, dx, ax);
    ay = ptx::fma2(s, dy, ay);
    az = ptx::fma2(s, dz, az);
}

template <int RP, int UNROLL, bool MASKED, int ABL = 0, bool UNSCALE = false>
__device__ __forceinline__ void itile_loop(const float4* __restrict__ t, const f32x2 (&nx)[RP], const f32x2 (&ny)[RP],
                                           const f32x2 (&nz)[RP], const float eps, f32x2 (&ax)[RP], f32x2 (&ay)[RP],
                                           f32x2 (&az)[RP], float (&dmin)[2 * RP]) {
#pragma unroll UNROLL
    for (int p = 0; p < kGravTile / 2; ++p) {
        float4 lo = t[2 * p];      // x0 x1 y0 y1
        float4 hi = t[2 * p + 1];  // z0 z1 gm0 gm1
        if (UNSCALE) {  // tile stored range-shifted (see kGravPosScale): exact powers of two, undone exactly
            lo.x *= kGravPosUnscale, lo.y *= kGravPosUnscale, lo.z *= kGravPosUnscale, lo.w *= kGravPosUnscale;
            hi.x *= kGravPosUnscale, hi.y *= kGravPosUnscale;
        }
#pragma unroll
        for (int q = 0; q < RP; ++q)
            ipair_interaction<MASKED, ABL>(lo.x, lo.z, hi.x, hi.z, nx[q], ny[q], nz[q], eps, ax[q], ay[q], az[q], dmin[2 * q],
                                           dmin[2 * q + 1]);
#pragma unroll
        for (int q = 0; q < RP; ++q)
            ipair_interaction<MASKED, ABL>(lo.y, lo.w, hi.y, hi.w, nx[q], ny[q], nz[q], eps, ax[q], ay[q], az[q], dmin[2 * q],
                                           dmin[2 * q + 1]);
    }
}

template <int RI, int MIN_CTAS, int UNROLL, bool DETECT, int CW = 8, bool IPACK = false, int ABL = 0>
__global__ void __launch_bounds__(32 * CW + 32, MIN_CTAS) nbody_gravity_kernel(const GravityArgs a) {
    constexpr int kGravRowsPerThread = RI;
    constexpr int kGravConsumerWarps = CW;
    constexpr int kGravConsumerThreads = 32 * CW;
    constexpr int kGravRows = kGravConsumerThreads * RI;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4* tiles = reinterpret_cast<float4*>(smem_raw);
    // running prefix of the current chunk, one private float4 slot per (row, thread): the second
    // level of the summation (level 1 = one tile of 256 bodies in registers)
    float4* run = reinterpret_cast<float4*>(smem_raw + kGravSmemBytes);
    __shared__ __align__(8) uint64_t full_bar[kGravStages];
    __shared__ __align__(8) uint64_t empty_bar[kGravStages];

    const GravityWork& w = a.work;
    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < kGravStages; ++s) {
            ptx::mbar_init(&full_bar[s], 1);
            ptx::mbar_init(&empty_bar[s], kGravConsumerWarps);
        }
        ptx::mbar_fence_init();
    }
    __syncthreads();

    if (warp == kGravConsumerWarps) {
        // ---- TMA producer: one thread streams this CTA's tiles through the ring --------------
        if (lane == 0) {
            uint32_t kk = 0;                         // position in the ring, running across the phases
            uint32_t ready = 1u << (a.own_rank & 31);  // ranks whose slice of the mirror has been seen complete
            for (uint32_t ph = a.phase_begin; ph < a.phase_end; ++ph) {
                const GravityPhase& P = w.phase[ph];
                const Span span = span_of(blockIdx.x, w.grid, P.n_iblocks * P.n_vtiles);
                if (span.count == 0) continue;
                uint32_t v = span.begin % P.n_vtiles;
                for (uint32_t k = 0; k < span.count; ++k, ++kk) {
                    const uint32_t stage = kk % kGravStages;
                    const uint32_t phase = (kk / kGravStages) & 1u;
                    ptx::mbar_wait(&empty_bar[stage], phase ^ 1u);
                    uint32_t tile = v + P.tile_offset;
                    if (tile >= w.n_tiles_total) tile -= w.n_tiles_total;
                    if (a.flags) {
                        // the slice of another rank: its tail / pack kernel pushed it into this GPU's memory and then
                        // raised flags[rank] to the generation (release.sys); acquire it before the first bulk load
                        const uint32_t r = tile / a.tiles_per_rank;
                        if (!(ready >> r & 1u)) {
                            ptx::wait_generation(a.flags + r, a.wait_gen, a.error);
                            ptx::fence_proxy_async_global();
                            ready |= 1u << r;
                        }
                    }
                    ptx::mbar_arrive_expect_tx(&full_bar[stage], kGravTileBytes);
                    ptx::bulk_g2s(tiles + stage * kGravTile, a.posm + static_cast<size_t>(tile) * kGravTile,
                                  kGravTileBytes, &full_bar[stage]);
                    if (++v == P.n_vtiles) v = 0;
                }
            }
        }
        return;
    }

    // ---- consumers ---------------------------------------------------------------------------
    const uint32_t tid = threadIdx.x;  // 0..kGravConsumerThreads-1
    float nxi[kGravRowsPerThread], nyi[kGravRowsPerThread], nzi[kGravRowsPerThread];

    // row of the i-block that register slot r of this thread holds.  i-packed shapes give a warp RI*32 consecutive
    // rows, so the warp's own bodies sit in ONE j tile (one exact-path tile per i-block instead of RI).
    auto row_in_block = [&](int r) -> uint32_t {
        return IPACK ? (warp * (32u * RI) + r * 32u + lane) : (r * kGravConsumerThreads + tid);
    };
    auto load_rows = [&](uint32_t iblock) {
#pragma unroll
        for (int r = 0; r < kGravRowsPerThread; ++r) {
            const uint32_t local = iblock * kGravRows + row_in_block(r);
            float x = 0.f, y = 0.f, z = 0.f;
            if (local < a.row_count) {
                const float* p = a.ipos + 3ull * (a.row_begin + local);
                x = __ldg(p + 0);
                y = __ldg(p + 1);
                z = __ldg(p + 2);
            }
            if ((ABL & kAblOverflow) != 0) x *= kGravPosScale, y *= kGravPosScale, z *= kGravPosScale;
            nxi[r] = x;  // kept positive; negated at the use (see pair_interaction)
            nyi[r] = y;
            nzi[r] = z;
        }
    };
    const float eps = a.eps;
    const uint32_t K = w.chunk_tiles;

    uint32_t kk = 0;
    for (uint32_t ph = a.phase_begin; ph < a.phase_end; ++ph) {
        const GravityPhase& P = w.phase[ph];
        const Span span = span_of(blockIdx.x, w.grid, P.n_iblocks * P.n_vtiles);
        if (span.count == 0) continue;
        const uint32_t chunks_per_iblock = P.n_vtiles / K;
        uint32_t b = P.iblock_begin + span.begin / P.n_vtiles;
        uint32_t v = span.begin % P.n_vtiles;
        uint32_t cv = v / K;        // chunk of the phase the tile belongs to
        uint32_t pic = v - cv * K;  // position of the tile in its chunk
        // A span that begins inside a chunk stores that chunk's remaining tile sums one by one: the CTA holding the
        // chunk's first tile owns the running prefix, and the reduce kernel continues it in tile order.
        bool prefix = false;
        // v of the tile that holds this warp's own bodies (i-packed shapes: 32 * RI consecutive rows), if the outer
        // archetype is part of the nested query and the tile belongs to this phase: it goes straight to the exact path
        auto own_tile_of = [&](uint32_t iblock) -> uint32_t {
            if (a.self_offset == 0xffffffffu) return 0xffffffffu;
            const uint32_t first = a.self_offset + a.row_begin + iblock * kGravRows + warp * (32u * RI);
            const uint32_t tile = first / kGravTile;
            if ((first + 32u * RI - 1) / kGravTile != tile) return 0xffffffffu;  // the warp's rows straddle two tiles: no hint,
                                                                                 // each row's own tile is found through its poison
            const uint32_t ov = tile >= P.tile_offset ? tile - P.tile_offset : tile + w.n_tiles_total - P.tile_offset;
            return ov < P.n_vtiles ? ov : 0xffffffffu;
        };
        uint32_t own_v = own_tile_of(b);
        // a warp of the (partly filled) last i-block that holds no row at all skips the arithmetic and the fragments; it
        // still takes part in the ring.  (i-packed shapes: a warp's rows are consecutive.)
        auto warp_has_rows = [&](uint32_t iblock) -> bool {
            return !IPACK || iblock * kGravRows + warp * (32u * RI) < a.row_count;
        };
        bool active = warp_has_rows(b);
        load_rows(b);

        for (uint32_t k = 0; k < span.count; ++k, ++kk) {
            const uint32_t stage = kk % kGravStages;
            const uint32_t phase = (kk / kGravStages) & 1u;
            if (pic == 0) {
                prefix = true;
#pragma unroll
                for (int r = 0; r < kGravRowsPerThread; ++r) run[r * kGravConsumerThreads + tid] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            ptx::mbar_wait(&full_bar[stage], phase);
            const float4* t = tiles + stage * kGravTile;
            if (!active) {  // (warp-uniform)
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&empty_bar[stage]);
                if (++pic == K) pic = 0, ++cv;
                if (++v == P.n_vtiles) {
                    v = 0;
                    cv = 0;
                    ++b;
                    if (k + 1 < span.count) {
                        own_v = own_tile_of(b);
                        active = warp_has_rows(b);
                        load_rows(b);
                    }
                }
                continue;
            }

            // Which rows take this tile's EXACT result is decided row by row, from data alone: a row's own tile, and any
            // tile whose fast sum for that row is poisoned (or, for the min-d2 detectors, holds a pair at or below
            // epsilon).  Not "the thread redoes all its rows": which rows share a thread depends on the kernel shape and
            // on where a rank's rows begin, and results must not (kernels.cuh).
            float sx[kGravRowsPerThread], sy[kGravRowsPerThread], sz[kGravRowsPerThread];  // this tile's sums per row
            float dmin[kGravRowsPerThread];
            bool exact[kGravRowsPerThread];
#pragma unroll
            for (int r = 0; r < kGravRowsPerThread; ++r) dmin[r] = 3.0e38f, exact[r] = false;
            if constexpr (IPACK) {
                // lanes = rows (2q, 2q+1) of this thread
                constexpr int RP = RI / 2;
                f32x2 nx[RP], ny[RP], nz[RP], ax[RP], ay[RP], az[RP];
#pragma unroll
                for (int q = 0; q < RP; ++q) {
                    nx[q] = ptx::pack(-nxi[2 * q], -nxi[2 * q + 1]);
                    ny[q] = ptx::pack(-nyi[2 * q], -nyi[2 * q + 1]);
                    nz[q] = ptx::pack(-nzi[2 * q], -nzi[2 * q + 1]);
                    ax[q] = ay[q] = az[q] = 0ull;
                }
                constexpr bool kShifted = (ABL & kAblOverflow) != 0;
                bool any_exact = false;
                if constexpr (kShifted || DETECT) {
                    // Range-shifted fast path (kShifted): tile and row positions are stored times 2^-32 (an exact scaling;
                    // G*m is stored as is), so r^-3 of any pair with d2 <= 2^-21.3 -- a superset of the reference's
                    // `d2 <= f32::EPSILON` -- overflows to +inf and poisons the row's tile sums, which come out times 2^64.
                    // No per-pair detector instruction at all: one finiteness test per row and tile, then the exact masked
                    // path on true-scale values.  A row for which G*m/d^3 of some pair of the tile reaches 2^32 s^-2 (a
                    // density above 1e19 kg/m^3) or whose sums reach 2^64 m/s^2 takes the exact path as well.
                    // The tile that holds the warp's own bodies is known in advance and skips the doomed fast pass.
                    const bool own_tile = kShifted && v == own_v;
                    if (!own_tile) {
                        itile_loop<RP, UNROLL, false, ABL>(t, nx, ny, nz, eps, ax, ay, az, dmin);
#pragma unroll
                        for (int q = 0; q < RP; ++q) {
                            if (kShifted) {
                                const f32x2 sc = ptx::pack(kGravOutScale, kGravOutScale);
                                ax[q] = ptx::mul2(ax[q], sc);
                                ay[q] = ptx::mul2(ay[q], sc);
                                az[q] = ptx::mul2(az[q], sc);
                            }
                            ptx::unpack(ax[q], sx[2 * q], sx[2 * q + 1]);
                            ptx::unpack(ay[q], sy[2 * q], sy[2 * q + 1]);
                            ptx::unpack(az[q], sz[2 * q], sz[2 * q + 1]);
                        }
#pragma unroll
                        for (int r = 0; r < RI; ++r) {
                            // kShifted: inf or NaN in the row's sums; else: the detector sees the FUSED d2, the mask the
                            // reference's unfused one (an ulp or two apart), so everything within 1 + 1e-6 of epsilon goes too
                            exact[r] = kShifted ? !(fabsf(sx[r]) + fabsf(sy[r]) + fabsf(sz[r]) < 3.0e38f) : dmin[r] <= eps * 1.000001f;
                            any_exact = any_exact || exact[r];
                        }
                    } else {
#pragma unroll
                        for (int r = 0; r < RI; ++r) exact[r] = true;
                        any_exact = true;
                    }
                    if (any_exact) {
                        f32x2 ux[RP], uy[RP], uz[RP], ex[RP], ey[RP], ez[RP];
                        const f32x2 un = ptx::pack(kShifted ? kGravPosUnscale : 1.f, kShifted ? kGravPosUnscale : 1.f);
#pragma unroll
                        for (int q = 0; q < RP; ++q) {
                            ux[q] = ptx::mul2(nx[q], un);
                            uy[q] = ptx::mul2(ny[q], un);
                            uz[q] = ptx::mul2(nz[q], un);
                            ex[q] = ey[q] = ez[q] = 0ull;
                        }
                        itile_loop<RP, 2, true, 0, kShifted>(t, ux, uy, uz, eps, ex, ey, ez, dmin);
#pragma unroll
                        for (int q = 0; q < RP; ++q) {
                            float x0, x1, y0, y1, z0, z1;
                            ptx::unpack(ex[q], x0, x1);
                            ptx::unpack(ey[q], y0, y1);
                            ptx::unpack(ez[q], z0, z1);
                            if (exact[2 * q]) sx[2 * q] = x0, sy[2 * q] = y0, sz[2 * q] = z0;
                            if (exact[2 * q + 1]) sx[2 * q + 1] = x1, sy[2 * q + 1] = y1, sz[2 * q + 1] = z1;
                        }
                    }
                } else {
                    itile_loop<RP, UNROLL, true>(t, nx, ny, nz, eps, ax, ay, az, dmin);
#pragma unroll
                    for (int q = 0; q < RP; ++q) {
                        ptx::unpack(ax[q], sx[2 * q], sx[2 * q + 1]);
                        ptx::unpack(ay[q], sy[2 * q], sy[2 * q + 1]);
                        ptx::unpack(az[q], sz[2 * q], sz[2 * q + 1]);
                    }
                }
            } else {
                f32x2 ax[kGravRowsPerThread], ay[kGravRowsPerThread], az[kGravRowsPerThread];
#pragma unroll
                for (int r = 0; r < kGravRowsPerThread; ++r) ax[r] = ay[r] = az[r] = 0ull;
                auto fold = [&](f32x2 (&bx)[kGravRowsPerThread], f32x2 (&by)[kGravRowsPerThread], f32x2 (&bz)[kGravRowsPerThread], bool only_exact) {
#pragma unroll
                    for (int r = 0; r < kGravRowsPerThread; ++r) {
                        if (only_exact && !exact[r]) continue;
                        float x0, x1, y0, y1, z0, z1;
                        ptx::unpack(bx[r], x0, x1);
                        ptx::unpack(by[r], y0, y1);
                        ptx::unpack(bz[r], z0, z1);
                        sx[r] = x0 + x1;  // even + odd bodies of the tile
                        sy[r] = y0 + y1;
                        sz[r] = z0 + z1;
                    }
                };
                if (DETECT) {
                    tile_loop<RI, UNROLL, false>(t, nxi, nyi, nzi, eps, ax, ay, az, dmin);
                    fold(ax, ay, az, false);
                    bool any_exact = false;
#pragma unroll
                    for (int r = 0; r < kGravRowsPerThread; ++r) {
                        exact[r] = dmin[r] <= eps * 1.000001f;
                        any_exact = any_exact || exact[r];
                    }
                    if (any_exact) {
                        // some pair of a row in this tile is at or below epsilon (its own body, or a coincident one): the
                        // unmasked sums of that row are garbage (inf/NaN) -- that row takes the exact result.
                        f32x2 ex[kGravRowsPerThread], ey[kGravRowsPerThread], ez[kGravRowsPerThread];
#pragma unroll
                        for (int r = 0; r < kGravRowsPerThread; ++r) ex[r] = ey[r] = ez[r] = 0ull;
                        tile_loop<RI, 1, true>(t, nxi, nyi, nzi, eps, ex, ey, ez, dmin);
                        fold(ex, ey, ez, true);
                    }
                } else {
                    tile_loop<RI, UNROLL, true>(t, nxi, nyi, nzi, eps, ax, ay, az, dmin);
                    fold(ax, ay, az, false);
                }
            }

            // the tile is consumed: hand the stage back before the bookkeeping below
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&empty_bar[stage]);

            if (prefix) {
                // level 2: running prefix of the chunk (private shared-memory slots, no synchronisation needed)
                if (pic + 1 == K || k + 1 == span.count) {
                    float4* dst = a.frags + static_cast<size_t>(P.frag_base + (b - P.iblock_begin) * chunks_per_iblock + cv) * kGravRows;
#pragma unroll
                    for (int r = 0; r < kGravRowsPerThread; ++r) {
                        float4 acc = run[r * kGravConsumerThreads + tid];
                        acc.x += sx[r];
                        acc.y += sy[r];
                        acc.z += sz[r];
                        __stcs(dst + row_in_block(r), acc);
                    }
                } else {
#pragma unroll
                    for (int r = 0; r < kGravRowsPerThread; ++r) {
                        float4 acc = run[r * kGravConsumerThreads + tid];
                        acc.x += sx[r];
                        acc.y += sy[r];
                        acc.z += sz[r];
                        run[r * kGravConsumerThreads + tid] = acc;
                    }
                }
            } else {
                float4* dst = a.frags + (static_cast<size_t>(P.tile_frag_base) + static_cast<size_t>(blockIdx.x) * (K - 1) + k) * kGravRows;
#pragma unroll
                for (int r = 0; r < kGravRowsPerThread; ++r) __stcs(dst + row_in_block(r), make_float4(sx[r], sy[r], sz[r], 0.f));
            }

            if (++pic == K) pic = 0, ++cv;
            if (++v == P.n_vtiles) {
                v = 0;
                cv = 0;
                ++b;
                if (k + 1 < span.count) {
                    own_v = own_tile_of(b);
                    active = warp_has_rows(b);
                    load_rows(b);
                }
            }
        }
    }
}

// Waits until every rank's previous generation has arrived here (nobody still reads the buffer this push overwrites
// on a peer: a rank that has pushed generation g-1 has finished every kernel that read generation g-2), CTA-wide.
__device__ __forceinline__ void push_begin(const PeerPush& p) {
    if (p.n == 0 || !p.wait) return;
    if (threadIdx.x < p.n) ptx::wait_generation(p.local_flags + threadIdx.x, p.gen - 1u, p.error);
    __syncthreads();
}
// After the CTA's stores: the last CTA of the grid raises this rank's flag on every peer.
__device__ __forceinline__ void push_end(const PeerPush& p) {
    if (p.n == 0) return;
    if (!p.signal) {  // a later launch of the same generation raises the flag (stream order; its fence covers these stores
        __threadfence_system();  // only if they are performed: fence here as well)
        return;
    }
    // one system-scope fence per CTA: the CTA barrier orders every thread's remote stores before thread 0's fence
    // (cumulativity), which orders them before the counter increment the last CTA's flag stores follow
    __syncthreads();
    __shared__ uint32_t is_last;
    if (threadIdx.x == 0) {
        __threadfence_system();
        is_last = atomicAdd(p.done_counter, 1u) == gridDim.x - 1 ? 1u : 0u;
    }
    __syncthreads();
    if (is_last) {
        if (threadIdx.x == 0) *p.done_counter = 0;  // the next pushing kernel follows in stream order
        __threadfence_system();
        if (threadIdx.x < p.n) ptx::st_release_sys(p.flag[threadIdx.x], p.gen);
    }
}
// One refreshed mirror entry: local mirror, or the same slot of every rank's buffer.
__device__ __forceinline__ void store_posm(const PeerPush& p, float* local, uint64_t J, float x, float y, float z, float gm) {
    const size_t o = (J >> 1) * 8 + (J & 1);
    if (p.n == 0) {
        float* d = local + o;
        d[0] = x, d[2] = y, d[4] = z, d[6] = gm;
    } else {
        for (uint32_t q = 0; q < p.n; ++q) {
            float* d = p.posm[q] + o;
            d[0] = x, d[2] = y, d[4] = z, d[6] = gm;
        }
    }
}

__device__ __forceinline__ bool valid_work(const GravityWork& w) { return w.n_iblocks != 0 && w.grid != 0; }
constexpr uint32_t kGravMaxChunks = 64;  // chunks of the summation tree: kGravTargetChunks plus the padding of up to 16 ranks

// Where the pieces of every chunk sum of i-block b are -- the same for all its rows, worked out once per CTA:
//   slot[c]            prefix fragment of chunk c
//   [ext_lo, ext_hi)   units of the chunk beyond the span of the CTA that owns its first tile (tile fragments); empty
//                      for all but the chunks a span boundary falls into
struct ChunkTable {
    uint32_t slot[kGravMaxChunks], ext_lo[kGravMaxChunks], ext_hi[kGravMaxChunks], phase[kGravMaxChunks];
};
__device__ __forceinline__ void build_chunk_table(const GravityWork& w, uint32_t b, ChunkTable& t, bool use_chunksums) {
    const uint32_t K = w.chunk_tiles;
    const uint32_t n_chunks = w.n_tiles_total / K;
    if (threadIdx.x < n_chunks && valid_work(w)) {
        const uint32_t c = threadIdx.x;
        if (use_chunksums && b == w.heavy_iblock) {
            // the chunk sums of this i-block were worked out by nbody_gravity_chunksum_kernel
            t.slot[c] = w.chunksum_base + c;
            t.ext_lo[c] = t.ext_hi[c] = 0;
            t.phase[c] = 0;
            return;
        }
        const uint32_t tile0 = c * K;
        uint32_t ph = 0, v0 = 0;
        for (; ph < w.n_phases; ++ph) {  // the phase whose i-block range holds b and whose rotated tile range holds the chunk
            const GravityPhase& Q = w.phase[ph];
            if (b < Q.iblock_begin || b >= Q.iblock_begin + Q.n_iblocks) continue;
            v0 = tile0 >= Q.tile_offset ? tile0 - Q.tile_offset : tile0 + w.n_tiles_total - Q.tile_offset;
            if (v0 < Q.n_vtiles) break;
        }
        const GravityPhase& P = w.phase[ph < w.n_phases ? ph : 0];
        const uint32_t units = P.n_iblocks * P.n_vtiles;
        const uint32_t u0 = (b - P.iblock_begin) * P.n_vtiles + v0;
        const Span s0 = span_of(cta_of(u0, w.grid, units), w.grid, units);
        t.phase[c] = ph;
        t.slot[c] = P.frag_base + (b - P.iblock_begin) * (P.n_vtiles / K) + v0 / K;
        t.ext_lo[c] = min(u0 + K, s0.begin + s0.count);
        t.ext_hi[c] = u0 + K;
    }
}
// One chunk sum of one row: the prefix fragment continued with the tile fragments other CTAs stored one by one, in tile
// order (fetched kFragBatch at a time -- independent loads in flight -- and added in order).
constexpr uint32_t kFragBatch = 8;
__device__ __forceinline__ float4 chunk_sum(const GravityReduceArgs& a, const ChunkTable& t, uint32_t c, uint32_t idx, float4 pre) {
    const GravityWork& w = a.work;
    const uint32_t rows = a.iblock_rows, K = w.chunk_tiles;
    float cx = pre.x, cy = pre.y, cz = pre.z;
    if (t.ext_lo[c] < t.ext_hi[c]) {
        const GravityPhase& P = w.phase[t.phase[c]];
        const uint32_t units = P.n_iblocks * P.n_vtiles;
        uint32_t u = t.ext_lo[c];
        uint32_t g = cta_of(u, w.grid, units);
        Span sg = span_of(g, w.grid, units);
        while (u < t.ext_hi[c]) {
            while (u >= sg.begin + sg.count) sg = span_of(++g, w.grid, units);
            const uint32_t n = min(min(t.ext_hi[c], sg.begin + sg.count) - u, kFragBatch);  // consecutive slots of one CTA
            const float4* src = a.frags + (static_cast<size_t>(P.tile_frag_base) + static_cast<size_t>(g) * (K - 1) + (u - sg.begin)) * rows + idx;
            float4 f[kFragBatch];
#pragma unroll
            for (uint32_t j = 0; j < kFragBatch; ++j)
                if (j < n) f[j] = __ldcs(src + static_cast<size_t>(j) * rows);
#pragma unroll
            for (uint32_t j = 0; j < kFragBatch; ++j)
                if (j < n) cx += f[j].x, cy += f[j].y, cz += f[j].z;
            u += n;
        }
    }
    return make_float4(cx, cy, cz, 0.f);
}

// Chunk sums of the rows of the separately dealt last i-block: its spans are a few tiles long, so nearly every tile sum
// is its own fragment and a row has thousands of pieces.  One warp per row, lane l sums chunks l and l + 32 (in tile
// order), and the sums land in the fragment workspace where the reduce kernel picks them up like prefix fragments.
__global__ void __launch_bounds__(256) nbody_gravity_chunksum_kernel(const GravityReduceArgs a) {
    __shared__ ChunkTable table;
    const GravityWork& w = a.work;
    const uint32_t b = w.heavy_iblock;
    build_chunk_table(w, b, table, false);
    __syncthreads();
    const uint32_t rows = a.iblock_rows;
    const uint32_t idx = blockIdx.x * 8 + (threadIdx.x >> 5);  // row of the i-block
    if (b * rows + idx >= a.row_count) return;
    const uint32_t n_chunks = w.n_tiles_total / w.chunk_tiles;
    float4* out = const_cast<float4*>(a.frags);
    for (uint32_t c = threadIdx.x & 31; c < n_chunks; c += 32) {
        const float4 pre = __ldcs(a.frags + static_cast<size_t>(table.slot[c]) * rows + idx);
        out[static_cast<size_t>(w.chunksum_base + c) * rows + idx] = chunk_sum(a, table, c, idx, pre);
    }
}

__global__ void __launch_bounds__(256) nbody_gravity_reduce_kernel(const GravityReduceArgs a) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = i < a.row_count;
    const GravityWork& w = a.work;
    const uint32_t rows = a.iblock_rows;  // a multiple of the CTA size: one i-block per CTA
    const uint32_t b = (blockIdx.x * blockDim.x) / rows;
    const uint32_t K = w.chunk_tiles;
    const uint32_t n_chunks = w.n_tiles_total / K;
    __shared__ ChunkTable table;
    build_chunk_table(w, b, table, true);
    __syncthreads();
    if (a.tail.posm) push_begin(a.push);
    float mx = 0.f, my = 0.f, mz = 0.f, mg = 0.f;  // this row's refreshed mirror entry
    if (valid) {
        const uint32_t idx = i - b * rows;
        // `.sum()` of Vector3 starts from zero (crates/n-body/src/lib.rs:129-133).
        float sx = 0.f, sy = 0.f, sz = 0.f;
        // prefix fragments of eight chunks are fetched together (independent loads in flight), then added in chunk order
        for (uint32_t c0 = 0; c0 < n_chunks; c0 += 8) {
            float4 pre[8];
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (c0 + u < n_chunks) pre[u] = __ldcs(a.frags + static_cast<size_t>(table.slot[c0 + u]) * rows + idx);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const uint32_t c = c0 + u;
                if (c >= n_chunks) break;
                const float4 cs = chunk_sum(a, table, c, idx, pre[u]);
                sx += cs.x;
                sy += cs.y;
                sz += cs.z;
            }
        }
        const uint64_t row = a.row_begin + i;
        float* out = a.acc + 3ull * row;
        out[0] = sx;
        out[1] = sy;
        out[2] = sz;
        if (a.tail.pos) {
            // movement: position.point += velocity * dt (crates/n-body/src/lib.rs:92), then
            // acceleration: velocity += acceleration * dt (:101) -- the order the schedule produced;
            // separate roundings like the Rust code
            float* p = a.tail.pos + 3ull * row;
            float* v = a.tail.vel + 3ull * row;
            const float vx = v[0], vy = v[1], vz = v[2];
            const float px = __fadd_rn(p[0], __fmul_rn(vx, a.tail.dt));
            const float py = __fadd_rn(p[1], __fmul_rn(vy, a.tail.dt));
            const float pz = __fadd_rn(p[2], __fmul_rn(vz, a.tail.dt));
            p[0] = px;
            p[1] = py;
            p[2] = pz;
            v[0] = __fadd_rn(vx, __fmul_rn(sx, a.tail.dt));
            v[1] = __fadd_rn(vy, __fmul_rn(sy, a.tail.dt));
            v[2] = __fadd_rn(vz, __fmul_rn(sz, a.tail.dt));
            if (a.tail.posm) {
                mx = px * a.tail.pscale;
                my = py * a.tail.pscale;
                mz = pz * a.tail.pscale;
                mg = __fmul_rn(a.tail.g, a.tail.mass[row]);
            }
        }
    }
    if (a.tail.posm) {
        // the refreshed mirror entry.  Two neighbouring rows share one 32-byte pair {x0 x1 y0 y1 | z0 z1 gm0 gm1}: lanes
        // swap halves so that every lane stores ONE aligned float4 per destination (the even body the first half, the odd
        // body the second) instead of four scattered 4-byte words -- what matters when the destinations are seven peers
        // across NVLink.
        const uint64_t J = a.tail.posm_offset + a.row_begin + i;
        const bool odd = J & 1;
        const float ox = __shfl_xor_sync(0xffffffffu, mx, 1), oy = __shfl_xor_sync(0xffffffffu, my, 1);
        const float oz = __shfl_xor_sync(0xffffffffu, mz, 1), og = __shfl_xor_sync(0xffffffffu, mg, 1);
        const bool partner_valid = __shfl_xor_sync(0xffffffffu, valid ? 1 : 0, 1) != 0;
        // the partner lane holds body J ^ 1 iff lane parity == body parity
        const bool paired = partner_valid && ((threadIdx.x & 1) != 0) == odd;
        if (valid) {
            if (paired) {
                const float4 half = odd ? make_float4(oz, mz, og, mg) : make_float4(mx, ox, my, oy);
                const size_t o = (J >> 1) * 8 + (odd ? 4 : 0);
                if (a.push.n == 0) {
                    *reinterpret_cast<float4*>(a.tail.posm + o) = half;
                } else {
                    for (uint32_t q = 0; q < a.push.n; ++q) *reinterpret_cast<float4*>(a.push.posm[q] + o) = half;
                }
            } else {
                store_posm(a.push, a.tail.posm, J, mx, my, mz, mg);
            }
        }
        push_end(a.push);
    }
}

__global__ void __launch_bounds__(256) pack_posm_kernel(const float* __restrict__ pos,
                                                        const float* __restrict__ mass, float g,
                                                        float* __restrict__ posm, uint64_t row_begin,
                                                        uint64_t n, uint64_t n_pad, uint64_t dst_offset, float pscale,
                                                        const PeerPush push) {
    const uint64_t k = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    push_begin(push);
    if (k < n) {
        const uint64_t row = row_begin + k;
        store_posm(push, posm, dst_offset + k, pos[3 * row + 0] * pscale, pos[3 * row + 1] * pscale,
                   pos[3 * row + 2] * pscale, __fmul_rn(g, mass[row]));
    } else if (k < n + n_pad) {
        store_posm(push, posm, dst_offset + k, 0.f, 0.f, 0.f, 0.f);  // null body: contributes exactly zero
    }
    push_end(push);
}

// ---- small worlds: many ticks in ONE launch ------------------------------------------------------------------------------
// At the reference's own benchmark size (1 000 bodies x 100 ticks, crates/ecs_bench_suite/benches/n_body.rs:4) a tick is a
// microsecond of arithmetic behind two kernel launches.  This kernel runs the whole fused tick (gravity -> movement ->
// acceleration, one body archetype) for n_ticks ticks: a cooperative grid of one CTA per 32 rows, warp = j tile, lane = row;
// every CTA bulk-copies the whole mirror (<= 64 KiB) into shared memory once per tick, a row's tile sums meet in shared
// memory and warp 0 adds them in tile order, integrates the row in registers (position and velocity never leave them
// between ticks) and stores the refreshed mirror entry into the OTHER mirror buffer (tick k reads buffer k & 1: a CTA may
// still be loading the current one); one grid-wide barrier per tick.
// Same arithmetic, same summation tree (tile sum from zero, K = 1 chunk = 0 + tile sum, row total from zero in tile order)
// and same roundings as nbody_gravity_kernel<2, 3, 8, false, 4> + nbody_gravity_reduce_kernel, which is what the same
// sizes run tick by tick: results are bit-identical (tests/test_gpu_nbody.py::test_resident_ticks_equal_single_ticks).
__global__ void __launch_bounds__(512, 1) nbody_resident_ticks_kernel(const ResidentTicksArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4* mirror = reinterpret_cast<float4*>(smem_raw);  // n_tiles x 256 bodies, pair-interleaved
    float4* sums = mirror + static_cast<size_t>(a.n_tiles) * kGravTile;  // [tile][row of the CTA]
    __shared__ __align__(8) uint64_t full_bar;
    __shared__ uint32_t stop;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t row = blockIdx.x * 32u + lane;
    const bool valid = row < a.n_rows;
    const size_t o = static_cast<size_t>(row >> 1) * 8 + (row & 1);  // float index of x in the mirror; y, z, gm at +2, +4, +6
    if (threadIdx.x == 0) {
        ptx::mbar_init(&full_bar, 1);
        ptx::mbar_fence_init();
        stop = 0;
    }
    __syncthreads();
    float px = 0.f, py = 0.f, pz = 0.f, vx = 0.f, vy = 0.f, vz = 0.f, sx = 0.f, sy = 0.f, sz = 0.f;
    if (warp == 0 && valid) {
        px = a.pos[3ull * row], py = a.pos[3ull * row + 1], pz = a.pos[3ull * row + 2];
        vx = a.vel[3ull * row], vy = a.vel[3ull * row + 1], vz = a.vel[3ull * row + 2];
    }
    const uint32_t bytes = a.n_tiles * kGravTileBytes;
    for (uint32_t k = 0; k < a.n_ticks; ++k) {
        const float* src = (k & 1u) ? a.posm[1] : a.posm[0];
        float* dst = (k & 1u) ? a.posm[0] : a.posm[1];
        if (threadIdx.x == 0) {
            // the other CTAs' stores of the previous tick (generic proxy, acquired at the barrier) before the bulk read
            ptx::fence_proxy_async_global();
            ptx::mbar_arrive_expect_tx(&full_bar, bytes);
            ptx::bulk_g2s(mirror, src, bytes, &full_bar);
        }
        ptx::mbar_wait(&full_bar, k & 1u);
        const float* mf = reinterpret_cast<const float*>(mirror);
        float fx = 0.f, fy = 0.f, fz = 0.f;
        if (valid) {
            const float nxi[1] = {mf[o]}, nyi[1] = {mf[o + 2]}, nzi[1] = {mf[o + 4]};
            f32x2 ax[1] = {0ull}, ay[1] = {0ull}, az[1] = {0ull};
            float dmin[1] = {3.0e38f};
            tile_loop<1, 8, true>(mirror + static_cast<size_t>(warp) * kGravTile, nxi, nyi, nzi, a.eps, ax, ay, az, dmin);
            float x0, x1, y0, y1, z0, z1;
            ptx::unpack(ax[0], x0, x1);
            ptx::unpack(ay[0], y0, y1);
            ptx::unpack(az[0], z0, z1);
            // even + odd bodies of the tile, then the chunk of one tile: its running prefix starts from zero
            fx = 0.f + (x0 + x1);
            fy = 0.f + (y0 + y1);
            fz = 0.f + (z0 + z1);
        }
        sums[warp * 32u + lane] = make_float4(fx, fy, fz, 0.f);
        __syncthreads();
        if (warp == 0 && valid) {
            sx = sy = sz = 0.f;  // `.sum()` starts from zero (crates/n-body/src/lib.rs:129-133)
            for (uint32_t t = 0; t < a.n_tiles; ++t) {
                const float4 f = sums[t * 32u + lane];
                sx += f.x, sy += f.y, sz += f.z;
            }
            // movement (:92), then acceleration (:101): separate roundings like the Rust code
            px = __fadd_rn(px, __fmul_rn(vx, a.dt));
            py = __fadd_rn(py, __fmul_rn(vy, a.dt));
            pz = __fadd_rn(pz, __fmul_rn(vz, a.dt));
            vx = __fadd_rn(vx, __fmul_rn(sx, a.dt));
            vy = __fadd_rn(vy, __fmul_rn(sy, a.dt));
            vz = __fadd_rn(vz, __fmul_rn(sz, a.dt));
            dst[o] = px, dst[o + 2] = py, dst[o + 4] = pz;  // (G*m is in both buffers and does not change)
        }
        __syncthreads();  // the mirror and the sums are free; this CTA's stores are issued
        if (k + 1 < a.n_ticks) {
            if (threadIdx.x == 0) {
                // bounded: a grid that is not co-resident (never the case for a cooperative launch) ends in an error, not a hang
                // (every CTA waits for the same count, so all of them give up at the same tick)
                if (!ptx::grid_barrier(a.barrier, a.barrier_base + (k + 1u) * gridDim.x, a.error)) stop = 1;
            }
            __syncthreads();
            if (stop) break;
        }
    }
    if (warp == 0 && valid) {
        a.pos[3ull * row] = px, a.pos[3ull * row + 1] = py, a.pos[3ull * row + 2] = pz;
        a.vel[3ull * row] = vx, a.vel[3ull * row + 1] = vy, a.vel[3ull * row + 2] = vz;
        a.acc[3ull * row] = sx, a.acc[3ull * row + 1] = sy, a.acc[3ull * row + 2] = sz;
    }
}

}  // namespace

namespace {
typedef void (*GravityKernel)(const GravityArgs);
// Compiled shapes.  profiles/r01_gravity_variant_sweep*.txt hold the sweeps they were picked from (sweep5 maps the
// numbering used while sweeping to this table).
//                    rows/thread, CTAs/SM, consumer warps, range-shifted posm
const GravityVariant kVariants[kGravVariantCount] = {
    {4, 1, 12, true},   // 0  default: i-packed, overflow detector on range-shifted data, tile loop unrolled 16 times
    {2, 1, 12, false},  // 1  j-packed, FMNMX detector (the default before the i-packed rewrite; A/B)
    {2, 3, 4, false},   // 2  small problems: j-packed, 256-row i-blocks, exact mask on every pair
    {2, 4, 4, true},    // 3  medium problems: i-packed overflow detector, 256-row i-blocks, 4 CTAs/SM
    {4, 1, 12, false},  // 4  i-packed, FMNMX detector, same operation order as 0: must equal 0 bit for bit
    {4, 1, 12, false},  // 5  i-packed, exact mask on every pair
    {4, 1, 8, true},    // 6  as 0 with 8 consumer warps
    {8, 1, 8, true},    // 7  as 0 with 4 row pairs per thread (half the LDS, worse ptxas schedule)
    {4, 1, 12, true},   // 8  as 0 with the tile loop unrolled 8 times: ptxas then rotates two accumulators through MOVs
                        //    (9 non-FMA instructions per 473 against 7 per 935): 1 % slower, profiles/r02_gravity_unroll_sweep.txt
#ifdef RECS_GPU_ABLATIONS  // never in the shipped library: tools/ builds only (make ablations)
    {4, 1, 12, true},   // 9  TIMING ONLY: 0 without MUFU.RSQ (wrong results)
    {4, 1, 12, false},  // 10 TIMING ONLY: 4 without the detector (wrong results for coincident bodies)
#endif
};
GravityKernel variant_kernel(int v) {
    switch (v) {
        case 0:
            return nbody_gravity_kernel<4, 1, 16, true, 12, true, kAblOverflow>;
        case 1:
            return nbody_gravity_kernel<2, 1, 8, true, 12>;
        case 2:
            return nbody_gravity_kernel<2, 3, 8, false, 4>;
        case 3:
            return nbody_gravity_kernel<2, 4, 8, true, 4, true, kAblOverflow>;
        case 4:
            return nbody_gravity_kernel<4, 1, 4, true, 12, true, kAblCubeFirst>;
        case 5:
            return nbody_gravity_kernel<4, 1, 4, false, 12, true>;
        case 6:
            return nbody_gravity_kernel<4, 1, 2, true, 8, true, kAblOverflow>;
        case 7:
            return nbody_gravity_kernel<8, 1, 2, true, 8, true, kAblOverflow>;
        case 8:
            return nbody_gravity_kernel<4, 1, 8, true, 12, true, kAblOverflow>;
#ifdef RECS_GPU_ABLATIONS
        case 9:
            return nbody_gravity_kernel<4, 1, 16, true, 12, true, kAblOverflow | kAblNoMufu>;
        case 10:
            return nbody_gravity_kernel<4, 1, 4, true, 12, true, kAblCubeFirst | kAblNoDetector>;
#endif
    }
    return nullptr;
}
}  // namespace

GravityVariant gravity_variant(int index) { return kVariants[index]; }
// tile ring + the running-prefix slots of one i-block
static int gravity_smem_bytes(int variant) { return kGravSmemBytes + kVariants[variant].rows() * (int)sizeof(float4); }

cudaError_t gravity_prepare() {
    for (int v = 0; v < kGravVariantCount; ++v) {
        cudaError_t e = cudaFuncSetAttribute(variant_kernel(v), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             gravity_smem_bytes(v));
        if (e != cudaSuccess) return e;
    }
    return cudaFuncSetAttribute(nbody_resident_ticks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)((size_t)kResidentMaxTiles * (kGravTileBytes + 32 * sizeof(float4))));
}

int gravity_max_ctas_per_sm(int variant) {
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, variant_kernel(variant), kVariants[variant].threads(),
                                                      gravity_smem_bytes(variant)) != cudaSuccess)
        return 0;
    return n;
}

cudaError_t launch_gravity(const GravityArgs& a, cudaStream_t s) {
    if (a.work.n_iblocks == 0 || a.phase_begin >= a.phase_end) return cudaSuccess;
    variant_kernel(a.variant)<<<a.work.grid, kVariants[a.variant].threads(), gravity_smem_bytes(a.variant), s>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_gravity_reduce(const GravityReduceArgs& a, cudaStream_t s) {
    if (a.row_count == 0 && !(a.tail.posm && a.push.n)) return cudaSuccess;
    // a pushing rank always launches (at least one CTA): its peers wait for the generation flag
    const uint32_t grid = std::max<uint32_t>((a.row_count + 255) / 256, 1u);
    nbody_gravity_reduce_kernel<<<grid, 256, 0, s>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_gravity_chunksum(const GravityReduceArgs& a, cudaStream_t s) {
    if (a.work.heavy_iblock == 0xffffffffu || a.work.heavy_iblock * a.iblock_rows >= a.row_count) return cudaSuccess;
    const uint32_t heavy_rows = a.row_count - a.work.heavy_iblock * a.iblock_rows;
    nbody_gravity_chunksum_kernel<<<(heavy_rows + 7) / 8, 256, 0, s>>>(a);
    return cudaGetLastError();
}

static size_t resident_smem_bytes(uint32_t n_tiles) { return (size_t)n_tiles * (kGravTileBytes + 32 * sizeof(float4)); }

cudaError_t launch_resident_ticks(const ResidentTicksArgs& a, cudaStream_t s) {
    if (a.n_ticks == 0 || a.n_rows == 0) return cudaSuccess;
    if (a.n_tiles == 0 || a.n_tiles > kResidentMaxTiles) return cudaErrorInvalidValue;
    ResidentTicksArgs args = a;
    void* params[1] = {&args};
    // cooperative: all CTAs resident at once, or the launch fails -- the grid barrier cannot wait for a CTA that never starts
    return cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(nbody_resident_ticks_kernel), dim3((a.n_rows + 31) / 32),
                                       dim3(32 * a.n_tiles), params, resident_smem_bytes(a.n_tiles), s);
}

cudaError_t launch_pack_posm(const float* pos, const float* mass, float g, float* posm,
                             uint64_t row_begin, uint64_t n, uint64_t n_pad, uint64_t dst_offset, float pscale,
                             const PeerPush& push, cudaStream_t s) {
    if (n + n_pad == 0 && push.n == 0) return cudaSuccess;
    // a pushing rank always launches (at least one CTA): its peers wait for the generation flag
    const uint64_t grid = std::max<uint64_t>((n + n_pad + 255) / 256, 1);
    pack_posm_kernel<<<static_cast<uint32_t>(grid), 256, 0, s>>>(pos, mass, g, posm, row_begin, n, n_pad,
                                                                dst_offset, pscale, push);
    return cudaGetLastError();
}

}  // namespace recs
