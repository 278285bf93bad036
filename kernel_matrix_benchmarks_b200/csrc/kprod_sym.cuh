// kprod_sym: Gaussian a_i = sum_j exp(-|y_i - y_j|^2) b_j when targets and sources are the same
// points (the reference's `same_points` flag, base.py:56-79; always the case in the kernel solve,
// bruteforce.py:193-199).  K is symmetric, so every kernel value is evaluated ONCE and used twice:
//     a_i += k_ij b_j   (row sum, as in kprod_direct)      a_j += k_ij b_i   (column sum)
// which halves the MUFU work -- the binding pipe of kprod_direct at D = 3.
//
// Shape of the computation
//   * Product form of kprod_direct.cuh: k_ij = 2^(-|u_i|^2) e_ij 2^(-|u_j|^2), e_ij = 2^(2 u_i.u_j), with
//     w_j = b_j 2^(-|u_j|^2) riding in the packed records.  Row sums R_i = sum_j e_ij w_j and column sums
//     C_j = sum_i e_ij w_i share e_ij; out_i = 2^(-|u_i|^2) (R_i + C_i).
//   * The N x N pair matrix is cut into units of (TILE_ROWS target rows) x (SB sources).  Only units on
//     or above the block diagonal are evaluated: the TB = TILE_ROWS / SB units that intersect a tile's own
//     rows ("diagonal" units) are evaluated in full with row sums only, the units to their right feed both
//     the row sums of the tile and the column sums of the block.
//   * Column sums of a unit: each thread reduces its R rows, each warp reduces its 32 lanes with a
//     transposing shuffle butterfly (9 SHFL per 8 sources instead of 40), the 16 warps meet in shared
//     memory and one coalesced store per unit parks the 512 sums in colpart[tile][source].
//   * The triangular unit list is cut stream-K style into equal contiguous ranges -- first across
//     `n_parts` GPUs, then across the resident CTAs -- so every SM of every GPU gets the same number of
//     kernel evaluations.  sym_combine_kernel adds, in a fixed order, the row pieces and the column pieces
//     that this part produced (bitwise deterministic); with n_parts > 1 the caller sums the parts'
//     outputs (one all-reduce of N floats).
#pragma once
#include "kprod_direct.cuh"

namespace kmb {

struct SymParams {
    const DirectStats* stats;
    const float4* rec;     // product-form records (pack_sources_kernel), nsb * SB of them
    float* rowsum;         // n_tiles * TILE_ROWS : row sums of tiles finished by one CTA
    float* rowpart;        // grid * 2 * TILE_ROWS: row sums of each CTA's first / last (cut) segment
    float* colpart;        // n_tiles * N_pad     : column sums of each tile's off-diagonal units
    float* out;            // N
    long long N, N_pad;
    long long unit_begin, unit_end;   // this launch's share of the triangular unit list
    int n_tiles, nsb, grid;
};

// units of tiles 0 .. I-1:  sum_{t<I} (nsb - TB t)
template <int TB>
__host__ __device__ __forceinline__ long long sym_prefix(long long I, long long nsb) {
    return I * nsb - (TB * I * (I - 1)) / 2;
}
// tile that owns unit u
template <int TB>
__host__ __device__ __forceinline__ int sym_tile_of(long long u, long long nsb, int n_tiles) {
    const double h = static_cast<double>(nsb) + 0.5 * TB;
    double disc = h * h - 2.0 * TB * static_cast<double>(u);
    if (disc < 0.0) disc = 0.0;
    long long I = static_cast<long long>((h - sqrt(disc)) / TB);
    if (I < 0) I = 0;
    if (I > n_tiles - 1) I = n_tiles - 1;
    while (I + 1 < n_tiles && sym_prefix<TB>(I + 1, nsb) <= u) ++I;
    while (I > 0 && sym_prefix<TB>(I, nsb) > u) --I;
    return static_cast<int>(I);
}

// POLY_: every POLY_-th exponential on the FMA pipe (see DirectCfg); CH_: sources per shuffle butterfly
template <int DP_, int POLY_ = 16, int MINB_ = 2, int CH_ = 8, int CONSUMERS_ = 512, int R_ = 4, int STAGES_ = 4>
struct SymCfg {
    static constexpr int DP = DP_, POLY = POLY_, MINB = MINB_, CH = CH_;
    static constexpr int R = R_, RP = R_ / 2, CONSUMERS = CONSUMERS_, THREADS = CONSUMERS + 32, WARPS = CONSUMERS / 32;
    static_assert(R % 2 == 0 && (CH == 8 || CH == 16), "rows are processed as packed pairs; butterflies of 8 or 16 sources");
    static constexpr int TILE_ROWS = CONSUMERS * R;
    static constexpr int STAGES = STAGES_;
    static constexpr int PAIRS = DP + 1, RECV = (PAIRS + 1) / 2;
    static_assert(RECV == 2, "32-byte records (D <= 3, E = 1)");
    static constexpr int SB = 512;
    static_assert(SB % CONSUMERS == 0 && TILE_ROWS % SB == 0, "whole sources per thread when the warps' column sums are combined");
    static constexpr int TB = TILE_ROWS / SB;
    static constexpr int STAGE_BYTES = SB * RECV * 16;
    static constexpr int COLBUF_BYTES = WARPS * SB * 4;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + COLBUF_BYTES + 2 * STAGES * 8 + 16;
};

template <class C>
__global__ void __launch_bounds__(C::THREADS, C::MINB)
kprod_sym_kernel(const SymParams P) {
    constexpr int DP = C::DP, RP = C::RP, R = C::R, SB = C::SB, STAGES = C::STAGES, RECV = C::RECV, TB = C::TB, CH = C::CH;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4* stage_base = reinterpret_cast<float4*>(smem_raw);
    float* colbuf = reinterpret_cast<float*>(smem_raw + STAGES * C::STAGE_BYTES);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_raw + STAGES * C::STAGE_BYTES + C::COLBUF_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;

    if (P.stats->use_product == 0) return;   // the data need the difference form: kprod_direct runs instead

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
                const int tile = sym_tile_of<TB>(u, nsb, P.n_tiles);
                const long long base = sym_prefix<TB>(tile, nsb);
                long long jb = static_cast<long long>(TB) * tile + (u - base);
                const long long end = min(sym_prefix<TB>(tile + 1, nsb), u1);
                for (; u < end; ++u, ++jb, ++it) {
                    const int stage = it % STAGES;
                    mbar_wait_backoff(&empty_bar[stage], ((it / STAGES) & 1) ^ 1);
                    mbar_arrive_expect_tx(&full_bar[stage], C::STAGE_BYTES);
                    tma_bulk_g2s(stage_base + stage * (SB * RECV), P.rec + jb * (SB * RECV), C::STAGE_BYTES, &full_bar[stage]);
                }
            }
        }
        return;
    }

    // -------------------------------- consumers --------------------------------
    const int lane = tid & 31, warp = tid >> 5;
    const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4, h2 = lane & 2;
    // the source of a chunk this lane ends up holding, and whether it is the lane that stores it
    const int jl = CH == 16 ? (h16 ? 8 : 0) + (h8 ? 4 : 0) + (h4 ? 2 : 0) + (h2 ? 1 : 0) : (h16 ? 4 : 0) + (h8 ? 2 : 0) + (h4 ? 1 : 0);
    float* my_col = colbuf + warp * SB + jl;
    const bool col_writer = CH == 16 ? (lane & 1) == 0 : (lane & 3) == 0;

    uint32_t it = 0;
    long long u = u0;
    while (u < u1) {
        const int tile = sym_tile_of<TB>(u, nsb, P.n_tiles);
        const long long base = sym_prefix<TB>(tile, nsb);
        const long long tile_end = sym_prefix<TB>(tile + 1, nsb);
        const long long jb0 = static_cast<long long>(TB) * tile + (u - base);
        const int cnt = static_cast<int>(min(tile_end, u1) - u);
        const long long row_base = static_cast<long long>(tile) * C::TILE_ROWS + tid;

        // this thread's rows, straight from the packed records: 2u (row operand of the exponent) and
        // w = b 2^(-|u|^2) (what the row contributes to the column sums)
        float2 xr[DP][RP], w[RP];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const long long row = row_base + static_cast<long long>(r) * C::CONSUMERS;
            float4 q0 = make_float4(0.f, 0.f, 0.f, 0.f), q1 = q0;
            if (row < P.N) {
                q0 = __ldg(P.rec + row * RECV);
                q1 = __ldg(P.rec + row * RECV + 1);
            }
            const float c[4] = {q0.x, q0.z, q1.x, q1.z};   // pairs are duplicated: [v0 v0 v1 v1 | v2 v2 w w]
#pragma unroll
            for (int d = 0; d < DP; ++d) {
                if (r & 1) xr[d][r >> 1].y = 2.f * c[d]; else xr[d][r >> 1].x = 2.f * c[d];
            }
            if (r & 1) w[r >> 1].y = c[DP]; else w[r >> 1].x = c[DP];
        }
        float2 acc[RP], tot[RP];
#pragma unroll
        for (int p = 0; p < RP; ++p) acc[p] = tot[p] = make_float2(0.f, 0.f);

        for (int k = 0; k < cnt; ++k, ++it) {
            const int stage = it % STAGES;
            mbar_wait(&full_bar[stage], (it / STAGES) & 1);
            const float4* rec = stage_base + stage * (SB * RECV);
            const long long jb = jb0 + k;
            const bool off_diagonal = jb >= static_cast<long long>(TB) * (tile + 1);

            if (!off_diagonal) {
                // sources are this tile's own rows: every (i, j) of the block is evaluated, row sums only
#pragma unroll(CH)
                for (int j = 0; j < SB; ++j) {
                    const float4 v0 = rec[j * RECV], v1 = rec[j * RECV + 1];
                    const float2 pr[4] = {make_float2(v0.x, v0.y), make_float2(v0.z, v0.w), make_float2(v1.x, v1.y), make_float2(v1.z, v1.w)};
#pragma unroll
                    for (int p = 0; p < RP; ++p) {
                        float2 s;
#pragma unroll
                        for (int d = 0; d < DP; ++d) s = (d == 0) ? mul2(xr[d][p], pr[d]) : fma2(xr[d][p], pr[d], s);
                        const bool poly = C::POLY > 0 && (RP * CH) % C::POLY == 0 && ((j % CH) * RP + p) % C::POLY == C::POLY - 1;
                        const float2 kv = poly ? ex2_poly2(s) : make_float2(ex2_approx(s.x), ex2_approx(s.y));
                        acc[p] = fma2(kv, pr[DP], acc[p]);
                    }
                }
            } else {
#pragma unroll 1
                for (int j = 0; j < SB; j += CH) {
                    float t[CH];
#pragma unroll
                    for (int c = 0; c < CH; ++c) {
                        const float4 v0 = rec[(j + c) * RECV], v1 = rec[(j + c) * RECV + 1];
                        const float2 pr[4] = {make_float2(v0.x, v0.y), make_float2(v0.z, v0.w), make_float2(v1.x, v1.y), make_float2(v1.z, v1.w)};
                        float2 kv[RP];
#pragma unroll
                        for (int p = 0; p < RP; ++p) {
                            float2 s;
#pragma unroll
                            for (int d = 0; d < DP; ++d) s = (d == 0) ? mul2(xr[d][p], pr[d]) : fma2(xr[d][p], pr[d], s);
                            const bool poly = C::POLY > 0 && (RP * CH) % C::POLY == 0 && (c * RP + p) % C::POLY == C::POLY - 1;
                            kv[p] = poly ? ex2_poly2(s) : make_float2(ex2_approx(s.x), ex2_approx(s.y));
                            acc[p] = fma2(kv[p], pr[DP], acc[p]);
                        }
                        // what this thread's R rows add to column j + c
                        float2 tt = mul2(kv[0], w[0]);
#pragma unroll
                        for (int p = 1; p < RP; ++p) tt = fma2(kv[p], w[p], tt);
                        t[c] = tt.x + tt.y;
                    }
                    // transposing butterfly: each xor step halves the columns a lane carries, until one is left
                    float cs;
                    if constexpr (CH == 16) {
                        float s8[8], s4[4], s2[2];
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const float send = h16 ? t[q] : t[q + 8], keep = h16 ? t[q + 8] : t[q];
                            s8[q] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                        }
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float send = h8 ? s8[q] : s8[q + 4], keep = h8 ? s8[q + 4] : s8[q];
                            s4[q] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
                        }
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            const float send = h4 ? s4[q] : s4[q + 2], keep = h4 ? s4[q + 2] : s4[q];
                            s2[q] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
                        }
                        const float send = h2 ? s2[0] : s2[1], keep = h2 ? s2[1] : s2[0];
                        cs = keep + __shfl_xor_sync(0xffffffffu, send, 2);
                        cs += __shfl_xor_sync(0xffffffffu, cs, 1);
                    } else {
                        float s4[4], s2[2];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float send = h16 ? t[q] : t[q + 4], keep = h16 ? t[q + 4] : t[q];
                            s4[q] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                        }
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            const float send = h8 ? s4[q] : s4[q + 2], keep = h8 ? s4[q + 2] : s4[q];
                            s2[q] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
                        }
                        const float send = h4 ? s2[0] : s2[1], keep = h4 ? s2[1] : s2[0];
                        cs = keep + __shfl_xor_sync(0xffffffffu, send, 4);
                        cs += __shfl_xor_sync(0xffffffffu, cs, 2);
                        cs += __shfl_xor_sync(0xffffffffu, cs, 1);
                    }
                    if (col_writer) my_col[j] = cs;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[stage]);
            // two-level summation of the row sums (see kprod_direct.cuh)
#pragma unroll
            for (int p = 0; p < RP; ++p) {
                tot[p] = add2(tot[p], acc[p]);
                acc[p] = make_float2(0.f, 0.f);
            }
            if (off_diagonal) {
                // the 16 warps' column sums of this unit -> one value per source -> colpart[tile][source]
                named_bar_sync(1, C::CONSUMERS);
#pragma unroll
                for (int sidx = tid; sidx < SB; sidx += C::CONSUMERS) {
                    float cs = 0.f;
#pragma unroll
                    for (int wv = 0; wv < C::WARPS; ++wv) cs += colbuf[wv * SB + sidx];
                    P.colpart[static_cast<size_t>(tile) * P.N_pad + jb * SB + sidx] = cs;
                }
                named_bar_sync(1, C::CONSUMERS);
            }
        }

        // ---------------------------- row sums of this segment ----------------------------
        const bool complete = (u == base) && (cnt == tile_end - base);
        float* dst = complete ? P.rowsum + static_cast<size_t>(tile) * C::TILE_ROWS
                              : P.rowpart + (static_cast<size_t>(blockIdx.x) * 2 + (u == u0 ? 0 : 1)) * C::TILE_ROWS;
#pragma unroll
        for (int r = 0; r < R; ++r) dst[tid + r * C::CONSUMERS] = (r & 1) ? tot[r >> 1].y : tot[r >> 1].x;
        u += cnt;
    }
}

// out_i = 2^(-|u_i|^2) * (row pieces + column pieces of this part), fixed summation order.
template <class C>
__global__ void __launch_bounds__(256)
sym_combine_kernel(const SymParams P) {
    constexpr int TB = C::TB, SB = C::SB, RECV = C::RECV, DP = C::DP;
    if (P.stats->use_product == 0) return;
    const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (i >= P.N) return;
    const long long nsb = P.nsb;
    const int G = P.grid;
    const long long ub = P.unit_begin, ue = P.unit_end, Ur = ue - ub;
    const int tile = static_cast<int>(i / C::TILE_ROWS);
    const int lr = static_cast<int>(i - static_cast<long long>(tile) * C::TILE_ROWS);
    const long long t_begin = sym_prefix<TB>(tile, nsb), t_end = sym_prefix<TB>(tile + 1, nsb);

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
    // column pieces: tile t < tile(i) evaluated unit (t, block of i) iff that unit is in this part's range
    float cs = 0.f;
    const long long jb = i / SB;
    for (int t = 0; t < tile; ++t) {
        const long long uid = sym_prefix<TB>(t, nsb) + (jb - static_cast<long long>(TB) * t);
        if (uid >= ub && uid < ue) cs += __ldcg(P.colpart + static_cast<size_t>(t) * P.N_pad + i);
    }
    const float4 q0 = __ldg(P.rec + i * RECV), q1 = __ldg(P.rec + i * RECV + 1);
    const float c[4] = {q0.x, q0.z, q1.x, q1.z};
    float n2 = 0.f;
#pragma unroll
    for (int d = 0; d < DP; ++d) n2 = fmaf(c[d], c[d], n2);
    P.out[i] = exp2f(-n2) * (rs + cs);
}

}  // namespace kmb
