// kprod_direct_f32: a_i = sum_j k(x_i, y_j) b_j for small D, K never materialised.
//
// Replaces kernel_matrix(..., fast_sqdists=False) + K @ b of the reference
// (/root/reference/kernel_matrix_benchmarks/algorithms/bruteforce.py:53-54, 18-22, 130-153).
//
// Shape of the computation
//   * Sources are pre-packed (pack_sources kernel) into records of duplicated pairs
//       [-s*y_0, -s*y_0, ..., -s*y_{DP-1}, -s*y_{DP-1}, b_0, b_0, ..., b_{EP-1}, b_{EP-1}]
//     (s folds log2(e) into the coordinates so the exponential is a bare MUFU.EX2), padded to a
//     whole number of SB-record blocks with records that evaluate to k*b == 0.
//   * A producer warp streams record blocks global -> shared with TMA 1-D bulk copies
//     (cp.async.bulk + mbarrier complete_tx) through a STAGES-deep ring.
//   * 256 consumer threads each keep R target rows in registers as R/2 packed pairs; every source
//     record is read with broadcast LDS.128 and applied to all R rows with packed FP32
//     (FADD2/FMUL2/FFMA2: two rows per issue slot) + one MUFU per pair.
//   * Work is split stream-K style: the N/TILE_ROWS x M/SB grid of (row tile, source block) units
//     is cut into G equal contiguous ranges, one per persistent CTA (G = resident CTAs), so every SM
//     finishes at the same time whatever N is.  A row tile that straddles CTAs is combined by the
//     last CTA to arrive (fixed summation order -> bitwise deterministic output).
#pragma once
#include "kmb_common.cuh"

namespace kmb {

// Filled on the device by direct_stats_kernel (no host round trip): the bounding-box centre the
// coordinates are shifted by, and which evaluation form the Gaussian kernel may use.
struct DirectStats {
    float center[16];
    float radius2;      // log2(e) * squared half-diagonal of the bounding box of x and y
    int use_product;    // 1: product form is accurate for this data (radius2 <= kProductFormRadius2)
    unsigned int counter;
    int pad;
};
// Gaussian "product form": with u = s(x-c), v = s(y-c), s^2 = log2(e):
//   exp(-|x-y|^2) = 2^(-|u|^2) * 2^(2 u.v) * 2^(-|v|^2)
// 2^(-|v|^2) is folded into the packed signal, 2^(-|u|^2) into the final store, which leaves
// FMUL2 + 2 FFMA2 + 2 MUFU + FFMA2 per two pairs (4 FMA-pipe slots instead of 7).  The exponent
// 2u.v is computed to ~3 ulp of its magnitude, so the form is only used when the centred data
// are small: |2u.v| <= 2*radius2 <= 12 keeps the relative error of every k below ~1e-6 and all
// factors within 2^+-12 (no overflow / underflow).  Otherwise the difference form runs.
constexpr float kProductFormRadius2 = 6.0f;

struct DirectParams {
    const float* x;       // (N, D) targets, unscaled
    const DirectStats* stats;
    const float4* rec;    // packed source records: n_src_blocks * SB records of RECV float4
    float* out;           // (N, E)
    float* partial;       // G * 2 slots * TILE_ROWS * PS floats
    int* tile_counter;    // n_tiles ints, zero on entry, zero again on exit
    long long N, M;
    long long row_offset; // global index of local row 0 (inverse-distance zeroing)
    int D, E;             // true dims (D <= DP); this launch covers signal columns e0 .. e0+EP-1
    int e0;
    int n_tiles, n_src_blocks;
    float xscale;         // coordinate scale folded into x (same as the records')
};

// FORM_: 0 = difference form (any kernel), 1 = Gaussian product form
// POLY_: product form only -- every POLY_-th exponential of a thread (0 = none) is evaluated on the
//        FMA pipe (Cody-Waite + degree-5 polynomial, packed FP32) instead of MUFU.EX2, which moves the
//        kernel past the 16 exp/clk/SM of the MUFU pipe
template <int DP_, int EP_, int R_, int KID_, bool NORM_, int FORM_ = 0, int CONSUMERS_ = 512, int UNROLL_ = 4,
          int MINB_ = 0, int STAGES_ = 4, int POLY_ = 0>
struct DirectCfg {
    static constexpr int POLY = POLY_;
    static_assert(POLY_ == 0 || FORM_ == 1, "the polynomial exp2 needs the bounded exponents of the product form");
    static constexpr int DP = DP_, EP = EP_, R = R_, KID = KID_, FORM = FORM_;
    static constexpr bool NORM = NORM_;
    static_assert(FORM == 0 || KID == KMB_KERNEL_GAUSSIAN, "product form is Gaussian-only");
    // product form + row normalisation: the record carries 2^(-|v|^2) as one more signal column
    static constexpr int WCOL = (FORM == 1 && NORM) ? 1 : 0;
    static constexpr int CONSUMERS = CONSUMERS_;
    static constexpr int UNROLL = UNROLL_;
    static constexpr int THREADS = CONSUMERS + 32;  // + one producer warp
    static constexpr int TILE_ROWS = CONSUMERS * R;
    static constexpr int STAGES = STAGES_;
    static constexpr int PAIRS = DP + EP + WCOL;    // float2 per record
    static constexpr int RECV = (PAIRS + 1) / 2;    // float4 per record
    // source records per stage: ~16 KB stages whatever the record size
    static constexpr int SB = RECV <= 2 ? 512 : RECV <= 4 ? 256 : RECV <= 8 ? 128 : 64;
    static constexpr int STAGE_BYTES = SB * RECV * 16;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 2 * STAGES * 8 + 16;
    // exp-type kernels under row normalisation carry a running max (online rescale)
    // (the product form only runs on data where nothing can underflow)
    static constexpr bool ONLINE_MAX = NORM && (KID != KMB_KERNEL_INVERSE_DISTANCE) && FORM == 0;
    // floats of partial state per row when a tile is split across CTAs
    static constexpr int PS = EP + (NORM ? 1 : 0) + (ONLINE_MAX ? 1 : 0);
    // resident CTAs per SM the register allocator is asked to make room for
    // (512 consumer threads + producer warp: 2 CTAs need <= 60 registers per thread)
    static constexpr int MINB = MINB_ > 0 ? MINB_ : ((!ONLINE_MAX && KID != KMB_KERNEL_INVERSE_DISTANCE && (DP + EP + WCOL) * R <= 16) ? 2 : 1);
    static_assert(R % 2 == 0, "rows are processed as packed pairs");
};

// owner CTA of unit u when U units are cut into G ranges [U*c/G, U*(c+1)/G)
__device__ __forceinline__ int unit_owner(long long u, long long U, int G) {
    return static_cast<int>(((u + 1) * G - 1) / U);
}

// 2^s for two packed exponents, |s| < 2^22, entirely on the FMA/ALU pipes: round-to-nearest split
// s = n + r (magic-number add), degree-5 minimax polynomial of 2^r on [-0.5, 0.5] (max relative
// error 7.5e-8 before rounding, 2.3e-7 in FP32 Horner form -- the same class as MUFU.EX2's 2 ulp),
// then n is added straight into the exponent field.
__device__ __forceinline__ float2 ex2_poly2(float2 s) {
    const float2 magic = make_float2(12582912.f, 12582912.f);   // 1.5 * 2^23
    const float2 t = add2(s, magic);                            // low mantissa bits = round(s)
    const float2 n = add2(t, make_float2(-12582912.f, -12582912.f));
    const float2 r = add2(s, make_float2(-n.x, -n.y));
    float2 q = fma2(r, make_float2(1.327647129e-03f, 1.327647129e-03f), make_float2(9.675540961e-03f, 9.675540961e-03f));
    q = fma2(q, r, make_float2(5.550713092e-02f, 5.550713092e-02f));
    q = fma2(q, r, make_float2(2.402212024e-01f, 2.402212024e-01f));
    q = fma2(q, r, make_float2(6.931469440e-01f, 6.931469440e-01f));
    q = fma2(q, r, make_float2(1.000000119e+00f, 1.000000119e+00f));
    return make_float2(__int_as_float(__float_as_int(q.x) + (__float_as_int(t.x) << 23)),
                       __int_as_float(__float_as_int(q.y) + (__float_as_int(t.y) << 23)));
}

template <int KID>
__device__ __forceinline__ float kernel_value(float s) {
    if constexpr (KID == KMB_KERNEL_GAUSSIAN) return ex2_approx(-s);
    else if constexpr (KID == KMB_KERNEL_ABSOLUTE_EXPONENTIAL) return ex2_approx(-sqrt_approx(s));
    else return rsqrt_approx(s);
}
// exponent (log2 domain) of an exp-type kernel, for the online max
template <int KID>
__device__ __forceinline__ float kernel_log2(float s) {
    if constexpr (KID == KMB_KERNEL_GAUSSIAN) return -s;
    else return -sqrt_approx(s);
}

// 2^(-|u|^2) of one target row (product form), recomputed at store time to keep it out of the loop
template <int DP>
__device__ __forceinline__ float product_row_scale(const DirectParams& P, long long row) {
    float n2 = 0.f;
#pragma unroll
    for (int d = 0; d < DP; ++d) {
        const float u = d < P.D ? (__ldg(P.x + row * P.D + d) - P.stats->center[d]) * P.xscale : 0.f;
        n2 = fmaf(u, u, n2);
    }
    return exp2f(-n2);
}

// (17 warps = 5 on one SM sub-partition of 16384 registers: 96 registers per thread is the hardware limit for one CTA per
// SM, 48 for two -- __maxnreg__(120) compiles and then fails to launch)
template <class C>
__global__ void __launch_bounds__(C::THREADS, C::MINB)
kprod_direct_kernel(const DirectParams P) {
    constexpr int DP = C::DP, EP = C::EP, R = C::R, RP = C::R / 2, KID = C::KID;
    constexpr int SB = C::SB, STAGES = C::STAGES, RECV = C::RECV;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4* stage_base = reinterpret_cast<float4*>(smem_raw);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_raw + STAGES * C::STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    int* s_flag = reinterpret_cast<int*>(empty_bar + STAGES);

    // both forms of a Gaussian product are enqueued; the data decide (on the device) which one runs
    if ((P.stats->use_product != 0) != (C::FORM == 1)) return;

    const int tid = threadIdx.x;
    const int G = gridDim.x;
    const long long nsb = P.n_src_blocks;
    const long long U = static_cast<long long>(P.n_tiles) * nsb;
    const long long u0 = U * blockIdx.x / G;
    const long long u1 = U * (blockIdx.x + 1) / G;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], C::CONSUMERS / 32);
        }
        fence_mbar_init();
    }
    __syncthreads();

    if (tid >= C::CONSUMERS) {
        // ------------------------------ producer warp ------------------------------
        if (tid == C::CONSUMERS) {
            uint32_t it = 0;
            for (long long u = u0; u < u1; ++u, ++it) {
                const int stage = it % STAGES;
                const uint32_t parity = (it / STAGES) & 1;
                mbar_wait_backoff(&empty_bar[stage], parity ^ 1);
                mbar_arrive_expect_tx(&full_bar[stage], C::STAGE_BYTES);
                const long long sblk = u % nsb;
                tma_bulk_g2s(stage_base + stage * (SB * RECV), P.rec + sblk * (SB * RECV), C::STAGE_BYTES,
                             &full_bar[stage]);
            }
        }
        return;
    }

    // -------------------------------- consumers --------------------------------
    uint32_t it = 0;
    long long u = u0;
    while (u < u1) {
        const int tile = static_cast<int>(u / nsb);
        const long long sb0 = u - tile * nsb;
        const int cnt = static_cast<int>(min(nsb - sb0, u1 - u));
        const long long row_base = static_cast<long long>(tile) * C::TILE_ROWS + tid;

        // targets of this tile -> registers, rows (2p, 2p+1) packed: row = row_base + 256*r
        float2 xr[DP][RP];
        [[maybe_unused]] long long jz[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const long long row = row_base + static_cast<long long>(r) * C::CONSUMERS;
            const bool ok = row < P.N;
#pragma unroll
            for (int d = 0; d < DP; ++d) {
                // centred, scaled target: -u for the difference form (records hold +v), 2u for the product form
                float v = (ok && d < P.D) ? (__ldg(P.x + row * P.D + d) - P.stats->center[d]) * P.xscale : 0.f;
                v = (C::FORM == 1) ? 2.f * v : -v;
                if (r & 1) xr[d][r >> 1].y = v; else xr[d][r >> 1].x = v;
            }
            if constexpr (KID == KMB_KERNEL_INVERSE_DISTANCE) jz[r] = (P.row_offset + row) % (P.M + 1);
        }
        // Two-level summation: acc/ksum collect one source block (SB terms), then fold into
        // tot/ktot.  A single FP32 running sum over 10^6 sources loses ~sqrt(M) ulps (2e-5 relative
        // at M = 1M, measured); per-block partial sums keep it near 1e-6.
        float2 acc[EP][RP];
        [[maybe_unused]] float2 ksum[RP];
        [[maybe_unused]] float2 kmax[RP];
        [[maybe_unused]] float2 tot[EP][RP];
        [[maybe_unused]] float2 ktot[RP];
#pragma unroll
        for (int p = 0; p < RP; ++p) {
#pragma unroll
            for (int e = 0; e < EP; ++e) acc[e][p] = tot[e][p] = make_float2(0.f, 0.f);
            ksum[p] = ktot[p] = make_float2(0.f, 0.f);
            kmax[p] = make_float2(-INFINITY, -INFINITY);
        }

        for (int k = 0; k < cnt; ++k, ++it) {
            const int stage = it % STAGES;
            mbar_wait(&full_bar[stage], (it / STAGES) & 1);
            const float4* rec = stage_base + stage * (SB * RECV);
            [[maybe_unused]] const long long j_base = (sb0 + k) * SB;

            if constexpr (!C::ONLINE_MAX) {
#pragma unroll(C::UNROLL)
                for (int j = 0; j < SB; ++j) {
                    float4 v[RECV];
#pragma unroll
                    for (int q = 0; q < RECV; ++q) v[q] = rec[j * RECV + q];
                    const float2* pr = reinterpret_cast<const float2*>(v);
#pragma unroll
                    for (int p = 0; p < RP; ++p) {
                        float2 s, kv;
                        if constexpr (C::FORM == 1) {
                            // 2 u.v, then k' = 2^(2 u.v); 2^(-|v|^2) rides in the signal, 2^(-|u|^2) in the store
#pragma unroll
                            for (int d = 0; d < DP; ++d) s = (d == 0) ? mul2(xr[d][p], pr[d]) : fma2(xr[d][p], pr[d], s);
                            // a fixed share of the exponentials goes to the FMA pipe (see DirectCfg::POLY)
                            constexpr int kPairsPerStep = RP * C::UNROLL;
                            const bool poly = C::POLY > 0 && kPairsPerStep % C::POLY == 0 &&
                                              ((j % C::UNROLL) * RP + p) % C::POLY == C::POLY - 1;
                            if (poly) kv = ex2_poly2(s);
                            else kv = make_float2(ex2_approx(s.x), ex2_approx(s.y));
                        } else {
#pragma unroll
                            for (int d = 0; d < DP; ++d) {
                                const float2 diff = add2(xr[d][p], pr[d]);
                                s = (d == 0) ? mul2(diff, diff) : fma2(diff, diff, s);
                            }
                            kv = make_float2(kernel_value<KID>(s.x), kernel_value<KID>(s.y));
                        }
                        if constexpr (KID == KMB_KERNEL_INVERSE_DISTANCE) {
                            // the reference's zeroing rule; padding records add exactly 0 (a 0/0 row stays NaN)
                            const bool pad = j_base + j >= P.M;
                            if (pad || j_base + j == jz[2 * p]) kv.x = 0.f;
                            if (pad || j_base + j == jz[2 * p + 1]) kv.y = 0.f;
                        }
#pragma unroll
                        for (int e = 0; e < EP; ++e) acc[e][p] = fma2(kv, pr[DP + e], acc[e][p]);
                        if constexpr (C::NORM) {
                            if constexpr (C::WCOL) ksum[p] = fma2(kv, pr[DP + EP], ksum[p]);
                            else ksum[p] = add2(ksum[p], kv);
                        }
                    }
                }
            } else {
                // online max-rescale over chunks of CH sources (attention, exp-type kernels):
                // acc and ksum are kept relative to 2^kmax so rows whose kernels all underflow in
                // FP32 still normalise (the float64 reference only underflows beyond d^2 ~ 745).
                constexpr int CH = 4;
                for (int j = 0; j < SB; j += CH) {
                    float2 t[CH][RP];
                    float2 cm[RP];
#pragma unroll
                    for (int c = 0; c < CH; ++c) {
                        const float2* pr = reinterpret_cast<const float2*>(rec + (j + c) * RECV);
#pragma unroll
                        for (int p = 0; p < RP; ++p) {
                            float2 s;
#pragma unroll
                            for (int d = 0; d < DP; ++d) {
                                const float2 diff = add2(xr[d][p], pr[d]);
                                s = (d == 0) ? mul2(diff, diff) : fma2(diff, diff, s);
                            }
                            t[c][p] = make_float2(kernel_log2<KID>(s.x), kernel_log2<KID>(s.y));
                            cm[p] = (c == 0) ? t[c][p]
                                             : make_float2(fmaxf(cm[p].x, t[c][p].x), fmaxf(cm[p].y, t[c][p].y));
                        }
                    }
#pragma unroll
                    for (int p = 0; p < RP; ++p) {
                        const float2 mnew = make_float2(fmaxf(kmax[p].x, cm[p].x), fmaxf(kmax[p].y, cm[p].y));
                        // rescale factor 2^(old - new); old == -inf -> 0 (acc is 0 anyway)
                        const float2 sc = make_float2(mnew.x == -INFINITY ? 1.f : ex2_approx(kmax[p].x - mnew.x),
                                                      mnew.y == -INFINITY ? 1.f : ex2_approx(kmax[p].y - mnew.y));
                        kmax[p] = mnew;
                        ksum[p] = mul2(ksum[p], sc);
#pragma unroll
                        for (int e = 0; e < EP; ++e) acc[e][p] = mul2(acc[e][p], sc);
                        const float2 moff = make_float2(mnew.x == -INFINITY ? 0.f : mnew.x,
                                                        mnew.y == -INFINITY ? 0.f : mnew.y);
#pragma unroll
                        for (int c = 0; c < CH; ++c) {
                            const float2* pr = reinterpret_cast<const float2*>(rec + (j + c) * RECV);
                            const float2 kv = make_float2(ex2_approx(t[c][p].x - moff.x), ex2_approx(t[c][p].y - moff.y));
#pragma unroll
                            for (int e = 0; e < EP; ++e) acc[e][p] = fma2(kv, pr[DP + e], acc[e][p]);
                            ksum[p] = add2(ksum[p], kv);
                        }
                    }
                }
            }
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(&empty_bar[stage]);
            if constexpr (!C::ONLINE_MAX) {
#pragma unroll
                for (int p = 0; p < RP; ++p) {
#pragma unroll
                    for (int e = 0; e < EP; ++e) {
                        tot[e][p] = add2(tot[e][p], acc[e][p]);
                        acc[e][p] = make_float2(0.f, 0.f);
                    }
                    if constexpr (C::NORM) {
                        ktot[p] = add2(ktot[p], ksum[p]);
                        ksum[p] = make_float2(0.f, 0.f);
                    }
                }
            }
        }
        if constexpr (!C::ONLINE_MAX) {  // from here on acc/ksum hold the segment totals
#pragma unroll
            for (int p = 0; p < RP; ++p) {
#pragma unroll
                for (int e = 0; e < EP; ++e) acc[e][p] = tot[e][p];
                ksum[p] = ktot[p];
            }
        }

        // ---------------------------- write this segment ----------------------------
        const bool complete = (cnt == nsb);
        if (complete) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const long long row = row_base + static_cast<long long>(r) * C::CONSUMERS;
                if (row < P.N) {
                    const float l = (r & 1) ? ksum[r >> 1].y : ksum[r >> 1].x;
                    [[maybe_unused]] float urow = 1.f;
                    if constexpr (C::FORM == 1 && !C::NORM) urow = product_row_scale<DP>(P, row);
#pragma unroll
                    for (int e = 0; e < EP; ++e) {
                        float v = (r & 1) ? acc[e][r >> 1].y : acc[e][r >> 1].x;
                        if constexpr (C::FORM == 1 && !C::NORM) v *= urow;
                        if constexpr (C::NORM) v = v / l;
                        if (P.e0 + e < P.E) P.out[row * P.E + P.e0 + e] = v;
                    }
                }
            }
        } else {
            // partial tile: park the state, the last CTA to arrive combines all segments in order
            const int slot = (u == u0) ? 0 : 1;
            float* mine = P.partial + (static_cast<size_t>(blockIdx.x) * 2 + slot) * (C::TILE_ROWS * C::PS);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int lr = tid + r * C::CONSUMERS;
#pragma unroll
                for (int e = 0; e < EP; ++e)
                    mine[e * C::TILE_ROWS + lr] = (r & 1) ? acc[e][r >> 1].y : acc[e][r >> 1].x;
                if constexpr (C::NORM) mine[EP * C::TILE_ROWS + lr] = (r & 1) ? ksum[r >> 1].y : ksum[r >> 1].x;
                if constexpr (C::ONLINE_MAX)
                    mine[(EP + 1) * C::TILE_ROWS + lr] = (r & 1) ? kmax[r >> 1].y : kmax[r >> 1].x;
            }
            __threadfence();
            named_bar_sync(1, C::CONSUMERS);
            const long long tile_u0 = static_cast<long long>(tile) * nsb;
            const int c_first = unit_owner(tile_u0, U, G);
            const int c_last = unit_owner(tile_u0 + nsb - 1, U, G);
            if (tid == 0) {
                const int old = atomicAdd(&P.tile_counter[tile], 1);
                const int last = (old == c_last - c_first);
                if (last) P.tile_counter[tile] = 0;  // leave the workspace reusable
                *s_flag = last;
            }
            named_bar_sync(1, C::CONSUMERS);
            const bool is_last = *s_flag != 0;
            named_bar_sync(1, C::CONSUMERS);  // s_flag may be rewritten by the next segment
            if (is_last) {
                __threadfence();
#pragma unroll 1
                for (int r = 0; r < R; ++r) {
                    const int lr = tid + r * C::CONSUMERS;
                    const long long row = static_cast<long long>(tile) * C::TILE_ROWS + lr;
                    if (row >= P.N) continue;
                    float tot[EP], l = 0.f, mx = -INFINITY;
#pragma unroll
                    for (int e = 0; e < EP; ++e) tot[e] = 0.f;
                    if constexpr (C::ONLINE_MAX) {
                        for (int c = c_first; c <= c_last; ++c) {
                            const int sl = (U * c / G) / nsb == tile ? 0 : 1;
                            const float* ps = P.partial + (static_cast<size_t>(c) * 2 + sl) * (C::TILE_ROWS * C::PS);
                            mx = fmaxf(mx, __ldcg(ps + (EP + 1) * C::TILE_ROWS + lr));
                        }
                    }
                    for (int c = c_first; c <= c_last; ++c) {
                        const int sl = (U * c / G) / nsb == tile ? 0 : 1;
                        const float* ps = P.partial + (static_cast<size_t>(c) * 2 + sl) * (C::TILE_ROWS * C::PS);
                        float w = 1.f;
                        if constexpr (C::ONLINE_MAX) {
                            const float m = __ldcg(ps + (EP + 1) * C::TILE_ROWS + lr);
                            w = (m == -INFINITY) ? 0.f : ex2_approx(m - mx);
                        }
#pragma unroll
                        for (int e = 0; e < EP; ++e) tot[e] = fmaf(w, __ldcg(ps + e * C::TILE_ROWS + lr), tot[e]);
                        if constexpr (C::NORM) l = fmaf(w, __ldcg(ps + EP * C::TILE_ROWS + lr), l);
                    }
                    [[maybe_unused]] float urow = 1.f;
                    if constexpr (C::FORM == 1 && !C::NORM) urow = product_row_scale<DP>(P, row);
#pragma unroll
                    for (int e = 0; e < EP; ++e)
                        if (P.e0 + e < P.E) P.out[row * P.E + P.e0 + e] = C::NORM ? tot[e] / l : tot[e] * urow;
                }
            }
        }
        u += cnt;
    }
}

// ---- data statistics: bounding box -> centre, radius, evaluation form ------------------------------
// One pass over y (and x unless it aliases y).  Per-block partial boxes, combined by the last block.
constexpr int STATS_THREADS = 256;
constexpr int STATS_MAX_BLOCKS = 256;

static __global__ void __launch_bounds__(STATS_THREADS)
direct_stats_kernel(const float* __restrict__ x, long long N, const float* __restrict__ y, long long M, int D,
                    float* __restrict__ block_box /* [blocks][2][16] */, DirectStats* stats, int allow_product) {
    __shared__ float s_lo[STATS_THREADS / 32][16], s_hi[STATS_THREADS / 32][16];
    __shared__ bool is_last;
    float lo[16], hi[16];
#pragma unroll
    for (int d = 0; d < 16; ++d) { lo[d] = INFINITY; hi[d] = -INFINITY; }
    const long long total = M + (x == y && N == M ? 0 : N);
    for (long long i = blockIdx.x * static_cast<long long>(STATS_THREADS) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * STATS_THREADS) {
        const float* pnt = i < M ? y + i * D : x + (i - M) * D;
#pragma unroll
        for (int d = 0; d < 16; ++d)
            if (d < D) { const float v = pnt[d]; lo[d] = fminf(lo[d], v); hi[d] = fmaxf(hi[d], v); }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 0; d < 16; ++d) {
        if (d < D) {
            for (int o = 16; o > 0; o >>= 1) {
                lo[d] = fminf(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], o));
                hi[d] = fmaxf(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], o));
            }
            if (lane == 0) { s_lo[warp][d] = lo[d]; s_hi[warp][d] = hi[d]; }
        }
    }
    __syncthreads();
    if (threadIdx.x < D) {
        float l = INFINITY, h = -INFINITY;
        for (int w = 0; w < STATS_THREADS / 32; ++w) { l = fminf(l, s_lo[w][threadIdx.x]); h = fmaxf(h, s_hi[w][threadIdx.x]); }
        block_box[(blockIdx.x * 2 + 0) * 16 + threadIdx.x] = l;
        block_box[(blockIdx.x * 2 + 1) * 16 + threadIdx.x] = h;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int old = atomicAdd(&stats->counter, 1u);
        is_last = (old == gridDim.x - 1);
        if (is_last) stats->counter = 0;
    }
    __syncthreads();
    if (is_last && threadIdx.x == 0) {
        __threadfence();
        float r2 = 0.f;
        for (int d = 0; d < 16; ++d) {
            float l = INFINITY, h = -INFINITY;
            if (d < D)
                for (unsigned int b = 0; b < gridDim.x; ++b) {
                    l = fminf(l, __ldcg(&block_box[(b * 2 + 0) * 16 + d]));
                    h = fmaxf(h, __ldcg(&block_box[(b * 2 + 1) * 16 + d]));
                }
            const float c = d < D ? 0.5f * (l + h) : 0.f, half = d < D ? 0.5f * (h - l) : 0.f;
            stats->center[d] = c;
            r2 = fmaf(half, half, r2);
        }
        r2 *= 1.4426950408889634f;
        stats->radius2 = r2;
        // NaN/inf coordinates fail the comparison and fall back to the difference form
        stats->use_product = (allow_product && r2 <= kProductFormRadius2) ? 1 : 0;
    }
}

// ---- source packing -------------------------------------------------------------------------
// rec[j] = [v_jd (dup) for d < DP | b_je (dup) for e < EP | (w_j) | zero pad], v = s (y - c).
//   difference form: signal as given; padding records (j >= M) sit far away (k underflows to
//                    exactly 0 for the exp kernels) and carry b == 0.
//   product form   : signal pre-multiplied by w_j = 2^(-|v_j|^2) (and w_j itself as the extra
//                    column under row normalisation); padding records are all zero.
struct PackLayout {   // per evaluation form: the record geometry of the kernel that will read it
    long long M_pad;
    int wcol, pairs_per_rec;
};
static __global__ void pack_sources_kernel(const float* __restrict__ y, const float* __restrict__ b, float2* __restrict__ rec,
                                           const DirectStats* __restrict__ stats, long long M, int D, int E, int DP,
                                           int EP, PackLayout diff_form, PackLayout prod_form, int e0, float scale) {
    const bool product = stats->use_product != 0;
    const PackLayout L = product ? prod_form : diff_form;
    const long long j = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (j >= L.M_pad) return;
    float2* o = rec + j * L.pairs_per_rec;
    const bool live = j < M;
    float n2 = 0.f;
    for (int d = 0; d < DP; ++d) {
        float v = 0.f;
        if (live) v = d < D ? scale * (y[j * D + d] - stats->center[d]) : 0.f;
        else v = (d == 0 && !product) ? 1.0e18f : 0.f;
        n2 = fmaf(v, v, n2);
        o[d] = make_float2(v, v);
    }
    const float w = (product && live) ? exp2f(-n2) : (product ? 0.f : 1.f);
    for (int e = 0; e < EP; ++e) {
        float v = 0.f;
        if (live && e0 + e < E) v = (b ? b[j * E + e0 + e] : 1.f) * w;
        o[DP + e] = make_float2(v, v);
    }
    if (L.wcol) o[DP + EP] = make_float2(w, w);
    for (int q = DP + EP + L.wcol; q < L.pairs_per_rec; ++q) o[q] = make_float2(0.f, 0.f);
}

// ---- dispatch table entry ---------------------------------------------------------------------
struct DirectEntry {
    int DP, EP, KID, NORM, FORM, WCOL, R, SB, RECV, PS, TILE_ROWS, SMEM, THREADS;
    const void* func;
    cudaError_t (*launch)(const DirectParams&, int grid, cudaStream_t stream);
};

template <class C>
cudaError_t launch_direct(const DirectParams& P, int grid, cudaStream_t stream) {
    if (ensure_dyn_smem(reinterpret_cast<const void*>(&kprod_direct_kernel<C>), C::SMEM_BYTES) != KMB_OK)
        return cudaErrorInvalidValue;   // message already set
    kprod_direct_kernel<C><<<grid, C::THREADS, C::SMEM_BYTES, stream>>>(P);
    return cudaGetLastError();
}

template <class C>
constexpr DirectEntry make_direct_entry() {
    return DirectEntry{C::DP, C::EP, C::KID, C::NORM ? 1 : 0, C::FORM, C::WCOL, C::R, C::SB, C::RECV, C::PS, C::TILE_ROWS,
                       C::SMEM_BYTES, C::THREADS, reinterpret_cast<const void*>(&kprod_direct_kernel<C>),
                       &launch_direct<C>};
}

// rows per thread: keep the register-resident targets (DP*R) and accumulators (EP*R) bounded.
// Tuned on B200 with tools/tune_direct.cu: for D=3, E=1 the best shape is 512 consumer threads x
// 4 rows, 2 CTAs/SM (34 warps/SM, ~56 registers); larger D trade rows for registers.
constexpr int direct_rows(int DP, int EP, bool online_max) {
    int r = DP <= 4 ? 4 : DP <= 8 ? 4 : 2;
    if (online_max && r > 4) r = 4;
    if (EP * r > 16) r = 16 / EP;
    return r < 2 ? 2 : r;
}

// product form: unroll 8 sources and send every 16th exponential to the FMA pipe (tools/tune_direct.cu:
// 14.71 -> 15.08 pairs/clk/SM; larger shares co-saturate the FMA and MUFU pipes and lose throughput)
#define KMB_DIRECT_ENTRY(DP, EP, KID, NORM, FORM)                                                        \
    kmb::make_direct_entry<kmb::DirectCfg<DP, EP,                                                          \
        kmb::direct_rows(DP, EP, (NORM) && (KID) != KMB_KERNEL_INVERSE_DISTANCE && (FORM) == 0), KID, NORM, FORM, \
        512, ((FORM) == 1 ? 8 : 4), 0, 4, ((FORM) == 1 ? 16 : 0)>>()
// EP = 8 and 16 (two rows per thread): wide signals re-evaluate the kernel once per 8 / 16 columns instead of once per 4
#define KMB_DIRECT_ENTRIES_FOR_DP(DP, KID, NORM, FORM)                                  \
    KMB_DIRECT_ENTRY(DP, 1, KID, NORM, FORM), KMB_DIRECT_ENTRY(DP, 2, KID, NORM, FORM), \
        KMB_DIRECT_ENTRY(DP, 4, KID, NORM, FORM), KMB_DIRECT_ENTRY(DP, 8, KID, NORM, FORM), \
        KMB_DIRECT_ENTRY(DP, 16, KID, NORM, FORM)
#define KMB_DIRECT_TABLE(NAME, KID, NORM, FORM)                                                                      \
    namespace kmb {                                                                                                  \
    extern const DirectEntry NAME[];                                                                                 \
    extern const int NAME##_count;                                                                                   \
    const DirectEntry NAME[] = {KMB_DIRECT_ENTRIES_FOR_DP(2, KID, NORM, FORM), KMB_DIRECT_ENTRIES_FOR_DP(3, KID, NORM, FORM), \
                                KMB_DIRECT_ENTRIES_FOR_DP(4, KID, NORM, FORM), KMB_DIRECT_ENTRIES_FOR_DP(8, KID, NORM, FORM), \
                                KMB_DIRECT_ENTRIES_FOR_DP(16, KID, NORM, FORM)};                                       \
    const int NAME##_count = sizeof(NAME) / sizeof(NAME[0]);                                                         \
    }

}  // namespace kmb
