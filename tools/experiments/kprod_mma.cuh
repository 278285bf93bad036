// kprod_mma: Gaussian a_i = sum_j exp(-|x_i - y_j|^2) b_j for D <= 3, E = 1 with the exponent on the
// tensor cores.  Same arithmetic contract as kprod_direct's product form
// (/root/reference/kernel_matrix_benchmarks/algorithms/bruteforce.py:25-58, 153):
//     k_ij = 2^(-|u_i|^2) e_ij 2^(-|v_j|^2),  e_ij = 2^(2 u_i.v_j),  u = s (x - c), v = s (y - c), s^2 = log2 e
//
// Why: in kprod_direct / kprod_sym the FMA pipe and the issue slots are what is scarce -- the dot product
// 2u.v costs 3 packed FMA-pipe instructions per two pairs, every source record is fetched with two
// LDS.128 per four pairs, and (symmetric case) the column sums need a shuffle butterfly.  Here
//   * S = 2 U V^T is a K = 8 micro-GEMM on mma.sync.m16n8k8.tf32: K slots [u_hi(D) | u_lo(D) | 0] x
//     [v_hi(D) | v_hi(D) | 0] and, in a second MMA, [v_lo(D) | v_lo(D) | 0] -- all four hi/lo cross terms
//     of the TF32 split, i.e. an FP32-accurate exponent (operands are exact 2-term TF32 splits, the tensor
//     core multiplies them exactly and accumulates in FP32).  Two MMAs give a warp 128 exponents.
//   * The accumulator fragment is a 2 x 2 patch per thread (rows g, g+8; columns 2t, 2t+1): row sums are an
//     in-thread FFMA2 per two pairs, and the column sums of the symmetric case are an in-thread FFMA2 too,
//     reduced over the 8 row groups of the warp once per 8 sources x 64 rows (3 SHFL), not per source.
//   * One LDS.128 (B fragments of 8 sources) + one LDS.64 (their weights) serve 512 pairs of a warp.
// What is left per pair is ~1 MUFU.EX2 + 1/2 FFMA2 (+1/2 FFMA2 for the symmetric column sums), so a share
// of the exponentials can go to the now idle FMA pipe (packed polynomial exp2, POLY) and the kernel runs
// past the 16 exponentials/clk/SM of the MUFU pipe.
//
// Work split: units of TILE_ROWS target rows x SB sources, stream-K over the rectangular (general) or
// triangular (targets == sources, see kprod_sym.cuh) unit list, first across `n_parts` GPUs, then across
// the CTAs; mma_combine_kernel adds the row pieces (and column pieces) in a fixed order.
#pragma once
#include "../../kernel_matrix_benchmarks_b200/csrc/kprod_direct.cuh"

namespace kmb {

struct MmaParams {
    const DirectStats* stats;
    const float* x;        // (N, D) targets (== y in the symmetric case), unscaled
    const float4* recb;    // nsb * SB sources x 4 float4: B fragments (see pack_mma_kernel)
    const float* wv;       // nsb * SB weights w_j = b_j 2^(-|v_j|^2)
    float* rowsum;         // n_tiles * TILE_ROWS
    float* rowpart;        // grid * 2 * TILE_ROWS
    float* colpart;        // symmetric only: n_tiles * M_pad
    float* out;            // N
    long long N, M, M_pad;
    long long unit_begin, unit_end;
    int D, n_tiles, nsb, grid;
    float xscale;
};

__device__ __forceinline__ float tf32_rn(float x) {   // round to 10 explicit mantissa bits (ties away), pure ALU
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}

// D = A * B (+ C): m16n8k8, A row-major (16 x 8), B column-major (8 x 8), TF32 in, FP32 accumulate
__device__ __forceinline__ void mma_tf32_zero(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
        : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(0.f));
}
__device__ __forceinline__ void mma_tf32_acc(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// unit list: rectangular (every tile sees all nsb blocks) or triangular (tile I sees blocks TB*I ..)
template <bool SYM, int TB>
struct UnitMap {
    static __host__ __device__ __forceinline__ long long prefix(long long I, long long nsb) {
        return SYM ? I * nsb - (TB * I * (I - 1)) / 2 : I * nsb;
    }
    static __host__ __device__ __forceinline__ int tile_of(long long u, long long nsb, int n_tiles) {
        if (!SYM) return static_cast<int>(u / nsb);
        const double h = static_cast<double>(nsb) + 0.5 * TB;
        double disc = h * h - 2.0 * TB * static_cast<double>(u);
        if (disc < 0.0) disc = 0.0;
        long long I = static_cast<long long>((h - sqrt(disc)) / TB);
        if (I < 0) I = 0;
        if (I > n_tiles - 1) I = n_tiles - 1;
        while (I + 1 < n_tiles && prefix(I + 1, nsb) <= u) ++I;
        while (I > 0 && prefix(I, nsb) > u) --I;
        return static_cast<int>(I);
    }
    static __host__ __device__ __forceinline__ long long first_block(long long I) { return SYM ? TB * I : 0; }
};

// POLY_: every POLY_-th pair of exponentials on the FMA pipe (0 = none).  WARPS_ consumer warps x 64 rows.
template <bool SYM_, int POLY_ = 0, int WARPS_ = 16, int SB_ = 256, int STAGES_ = 4, int MINB_ = 1, int MT_ = 4>
struct MmaCfg {
    static constexpr bool SYM = SYM_;
    static constexpr int POLY = POLY_, WARPS = WARPS_, SB = SB_, STAGES = STAGES_, MINB = MINB_;
    static constexpr int MT = MT_;                     // m16 tiles per warp
    static constexpr int ROWS_PER_WARP = 16 * MT;
    static constexpr int CONSUMERS = 32 * WARPS, THREADS = CONSUMERS + 32;
    static constexpr int TILE_ROWS = WARPS * ROWS_PER_WARP;
    static_assert(TILE_ROWS % SB == 0 && SB % 8 == 0 && SB <= CONSUMERS, "tile / block geometry");
    static constexpr int TB = TILE_ROWS / SB;
    static constexpr int REC_BYTES = SB * 64, W_BYTES = SB * 4, STAGE_BYTES = REC_BYTES + W_BYTES;
    static constexpr int COLBUF_BYTES = SYM ? WARPS * SB * 4 : 0;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + COLBUF_BYTES + 2 * STAGES * 8 + 16;
    using Map = UnitMap<SYM, TB>;
};

template <class C>
__global__ void __launch_bounds__(C::THREADS, C::MINB)
kprod_mma_kernel(const MmaParams P) {
    constexpr int SB = C::SB, STAGES = C::STAGES, MT = C::MT;
    using Map = typename C::Map;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* stage_base = smem_raw;
    float* colbuf = reinterpret_cast<float*>(smem_raw + STAGES * C::STAGE_BYTES);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_raw + STAGES * C::STAGE_BYTES + C::COLBUF_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;

    if (P.stats->use_product == 0) return;   // data too spread out for the product form: kprod_direct runs instead

    const int tid = threadIdx.x;
    const int G = gridDim.x;
    const long long nsb = P.nsb;
    const long long Ur = P.unit_end - P.unit_begin;
    const long long u0 = P.unit_begin + Ur * blockIdx.x / G;
    const long long u1 = P.unit_begin + Ur * (blockIdx.x + 1) / G;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], C::WARPS);
        }
        fence_mbar_init();
    }
    __syncthreads();

    if (tid >= C::CONSUMERS) {
        // ------------------------------ producer warp ------------------------------
        if (tid == C::CONSUMERS) {
            uint32_t it = 0;
            long long u = u0;
            while (u < u1) {
                const int tile = Map::tile_of(u, nsb, P.n_tiles);
                const long long base = Map::prefix(tile, nsb);
                long long jb = Map::first_block(tile) + (u - base);
                const long long end = min(Map::prefix(tile + 1, nsb), u1);
                for (; u < end; ++u, ++jb, ++it) {
                    const int stage = it % STAGES;
                    mbar_wait_backoff(&empty_bar[stage], ((it / STAGES) & 1) ^ 1);
                    mbar_arrive_expect_tx(&full_bar[stage], C::STAGE_BYTES);
                    unsigned char* st = stage_base + stage * C::STAGE_BYTES;
                    tma_bulk_g2s(st, P.recb + jb * (SB * 4), C::REC_BYTES, &full_bar[stage]);
                    tma_bulk_g2s(st + C::REC_BYTES, P.wv + jb * SB, C::W_BYTES, &full_bar[stage]);
                }
            }
        }
        return;
    }

    // -------------------------------- consumers --------------------------------
    const int lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int D = P.D;

    uint32_t it = 0;
    long long u = u0;
    while (u < u1) {
        const int tile = Map::tile_of(u, nsb, P.n_tiles);
        const long long base = Map::prefix(tile, nsb);
        const long long tile_end = Map::prefix(tile + 1, nsb);
        const long long jb0 = Map::first_block(tile) + (u - base);
        const int cnt = static_cast<int>(min(tile_end, u1) - u);
        const long long row_w = static_cast<long long>(tile) * C::TILE_ROWS + warp * C::ROWS_PER_WARP;

        // A fragments of this warp's 4 x 16 rows: K slots [2u_hi (D) | 2u_lo (D) | 0]; thread (g, t) holds
        // slots t and t + 4 of rows g and g + 8 of every m-tile
        uint32_t afrag[MT][4];
        [[maybe_unused]] float2 wrow[MT][2];   // symmetric: (w_i, w_i) of the thread's rows
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const long long row = row_w + mt * 16 + g + 8 * h;
                float slot[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) slot[k] = 0.f;
                if (row < P.N) {
#pragma unroll
                    for (int d = 0; d < 3; ++d) {
                        if (d < D) {
                            const float uu = 2.f * (__ldg(P.x + row * D + d) - P.stats->center[d]) * P.xscale;
                            const float hi = tf32_rn(uu);
                            const float lo = tf32_rn(uu - hi);
                            // slot d <- hi, slot D + d <- lo  (D is 1..3: positions resolved with selects)
#pragma unroll
                            for (int k = 0; k < 6; ++k) {
                                if (k == d) slot[k] = hi;
                                if (k == D + d) slot[k] = lo;
                            }
                        }
                    }
                }
                float lo_slot = slot[0], hi_slot = slot[4];
#pragma unroll
                for (int k = 1; k < 4; ++k) {
                    if (t == k) { lo_slot = slot[k]; hi_slot = slot[k + 4]; }
                }
                afrag[mt][h] = __float_as_uint(lo_slot);       // a0 / a1: (row g / g+8, k = t)
                afrag[mt][2 + h] = __float_as_uint(hi_slot);   // a2 / a3: (row g / g+8, k = t + 4)
                if constexpr (C::SYM) {
                    const float w = row < P.N ? __ldg(P.wv + row) : 0.f;
                    wrow[mt][h] = make_float2(w, w);
                }
            }
        }
        float2 racc[MT][2];
        float rtot[MT][2];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
            racc[mt][0] = racc[mt][1] = make_float2(0.f, 0.f);
            rtot[mt][0] = rtot[mt][1] = 0.f;
        }

        for (int k = 0; k < cnt; ++k, ++it) {
            const int stage = it % STAGES;
            mbar_wait(&full_bar[stage], (it / STAGES) & 1);
            const float4* recb = reinterpret_cast<const float4*>(stage_base + stage * C::STAGE_BYTES);
            const float* wv = reinterpret_cast<const float*>(stage_base + stage * C::STAGE_BYTES + C::REC_BYTES);
            [[maybe_unused]] const long long jb = jb0 + k;
            [[maybe_unused]] const bool off_diagonal = C::SYM && jb >= static_cast<long long>(C::TB) * (tile + 1);

#pragma unroll 1
            for (int cb2 = 0; cb2 < SB / 16; ++cb2) {
#pragma unroll
              for (int half = 0; half < 2; ++half) {
                const int cb = cb2 * 2 + half;
                const float4 bf = recb[(cb * 8 + g) * 4 + t];                                // B fragments: source cb*8 + g
                const float2 wj = *reinterpret_cast<const float2*>(wv + cb * 8 + 2 * t);    // weights of columns 2t, 2t+1
                [[maybe_unused]] float2 cacc = make_float2(0.f, 0.f);
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    float s[4];
                    mma_tf32_zero(s, afrag[mt], __float_as_uint(bf.x), __float_as_uint(bf.y));   // u . v_hi
                    mma_tf32_acc(s, afrag[mt], __float_as_uint(bf.z), __float_as_uint(bf.w));    // u . v_lo
                    // e = 2^s for the 2 x 2 patch: rows g, g+8 x columns 2t, 2t+1
                    constexpr int kPairsPerStep = 2 * MT * 2;   // packed pairs per unrolled step (2 column blocks)
                    float2 e0, e1;
                    const bool poly0 = C::POLY > 0 && kPairsPerStep % C::POLY == 0 && (half * 2 * MT + 2 * mt) % C::POLY == C::POLY - 1;
                    const bool poly1 = C::POLY > 0 && kPairsPerStep % C::POLY == 0 && (half * 2 * MT + 2 * mt + 1) % C::POLY == C::POLY - 1;
                    if (poly0) e0 = ex2_poly2(make_float2(s[0], s[1]));
                    else e0 = make_float2(ex2_approx(s[0]), ex2_approx(s[1]));
                    if (poly1) e1 = ex2_poly2(make_float2(s[2], s[3]));
                    else e1 = make_float2(ex2_approx(s[2]), ex2_approx(s[3]));
                    racc[mt][0] = fma2(e0, wj, racc[mt][0]);
                    racc[mt][1] = fma2(e1, wj, racc[mt][1]);
                    if constexpr (C::SYM) {
                        cacc = fma2(e0, wrow[mt][0], cacc);
                        cacc = fma2(e1, wrow[mt][1], cacc);
                    }
                }
                if constexpr (C::SYM) {
                    if (off_diagonal) {
                        // column sums over the warp's 64 rows: reduce over g (lane bits 2..4); after the first
                        // (transposing) step a lane carries one of its two columns
                        const bool up = lane & 16;
                        const float send = up ? cacc.x : cacc.y, keep = up ? cacc.y : cacc.x;
                        float cs = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                        cs += __shfl_xor_sync(0xffffffffu, cs, 8);
                        cs += __shfl_xor_sync(0xffffffffu, cs, 4);
                        if ((lane & 12) == 0) colbuf[warp * SB + cb * 8 + 2 * t + (up ? 1 : 0)] = cs;
                    }
                }
            }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[stage]);
            // two-level summation of the row sums (see kprod_direct.cuh)
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    rtot[mt][h] += racc[mt][h].x + racc[mt][h].y;
                    racc[mt][h] = make_float2(0.f, 0.f);
                }
            }
            if constexpr (C::SYM) {
                if (off_diagonal) {
                    named_bar_sync(1, C::CONSUMERS);
                    if (tid < SB) {
                        float cs = 0.f;
#pragma unroll
                        for (int wv_ = 0; wv_ < C::WARPS; ++wv_) cs += colbuf[wv_ * SB + tid];
                        P.colpart[static_cast<size_t>(tile) * P.M_pad + jb * SB + tid] = cs;
                    }
                    named_bar_sync(1, C::CONSUMERS);
                }
            }
        }

        // ---------------------------- row sums of this segment ----------------------------
        const bool complete = (u == base) && (cnt == tile_end - base);
        float* dst = complete ? P.rowsum + static_cast<size_t>(tile) * C::TILE_ROWS
                              : P.rowpart + (static_cast<size_t>(blockIdx.x) * 2 + (u == u0 ? 0 : 1)) * C::TILE_ROWS;
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                float v = rtot[mt][h];
                v += __shfl_xor_sync(0xffffffffu, v, 1);
                v += __shfl_xor_sync(0xffffffffu, v, 2);
                if (t == 0) dst[warp * C::ROWS_PER_WARP + mt * 16 + g + 8 * h] = v;
            }
        }
        u += cnt;
    }
}

// out_i = 2^(-|u_i|^2) * (row pieces [+ column pieces] of this part), fixed summation order.
template <class C>
__global__ void __launch_bounds__(256)
mma_combine_kernel(const MmaParams P) {
    constexpr int TB = C::TB, SB = C::SB;
    using Map = typename C::Map;
    if (P.stats->use_product == 0) return;
    const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (i >= P.N) return;
    const long long nsb = P.nsb;
    const int G = P.grid;
    const long long ub = P.unit_begin, ue = P.unit_end, Ur = ue - ub;
    const int tile = static_cast<int>(i / C::TILE_ROWS);
    const int lr = static_cast<int>(i - static_cast<long long>(tile) * C::TILE_ROWS);
    const long long t_begin = Map::prefix(tile, nsb), t_end = Map::prefix(tile + 1, nsb);

    float rs = 0.f;
    const long long a = max(t_begin, ub), b = min(t_end, ue);
    if (a < b) {
        const int c_first = static_cast<int>(((a - ub + 1) * G - 1) / Ur);
        const int c_last = static_cast<int>(((b - ub) * G - 1) / Ur);
        for (int c = c_first; c <= c_last; ++c) {
            const long long c0 = ub + Ur * c / G, c1 = ub + Ur * (c + 1) / G;
            const long long s0 = max(c0, t_begin), s1 = min(c1, t_end);
            if (s0 >= s1) continue;
            if (s0 == t_begin && s1 == t_end) rs += __ldcg(P.rowsum + static_cast<size_t>(tile) * C::TILE_ROWS + lr);
            else rs += __ldcg(P.rowpart + (static_cast<size_t>(c) * 2 + (s0 == c0 ? 0 : 1)) * C::TILE_ROWS + lr);
        }
    }
    float cs = 0.f;
    if constexpr (C::SYM) {
        const long long jb = i / SB;
        for (int tt = 0; tt < tile; ++tt) {
            const long long uid = Map::prefix(tt, nsb) + (jb - static_cast<long long>(TB) * tt);
            if (uid >= ub && uid < ue) cs += __ldcg(P.colpart + static_cast<size_t>(tt) * P.M_pad + i);
        }
    }
    float n2 = 0.f;
    for (int d = 0; d < P.D; ++d) {
        const float uu = (__ldg(P.x + i * P.D + d) - P.stats->center[d]) * P.xscale;
        n2 = fmaf(uu, uu, n2);
    }
    P.out[i] = exp2f(-n2) * (rs + cs);
}

// Sources -> B fragments and weights.  For source j and lane-in-group t the float4 is
//   { B1[t], B1[t+4], B2[t], B2[t+4] },  B1 = [v_hi(D) | v_hi(D) | 0],  B2 = [v_lo(D) | v_lo(D) | 0]   (8 K slots)
// and w_j = b_j 2^(-|v_j|^2) (b == NULL: b = 1).  Padding sources (j >= M) are all zero.
static __global__ void pack_mma_kernel(const float* __restrict__ y, const float* __restrict__ b, float4* __restrict__ recb,
                                       float* __restrict__ wv, const DirectStats* __restrict__ stats, long long M,
                                       long long M_pad, int D, float scale) {
    const long long j = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (j >= M_pad) return;
    float b1[8], b2[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) b1[k] = b2[k] = 0.f;
    float w = 0.f;
    if (j < M) {
        float n2 = 0.f;
        for (int d = 0; d < D; ++d) {
            const float v = scale * (y[j * D + d] - stats->center[d]);
            n2 = fmaf(v, v, n2);
            const float hi = tf32_rn(v);
            const float lo = tf32_rn(v - hi);
            b1[d] = b1[D + d] = hi;
            b2[d] = b2[D + d] = lo;
        }
        w = (b ? b[j] : 1.f) * exp2f(-n2);
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) recb[j * 4 + t] = make_float4(b1[t], b1[t + 4], b2[t], b2[t + 4]);
    wv[j] = w;
}

}  // namespace kmb
