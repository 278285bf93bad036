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

struct DirectParams {
    const float* x;       // (N, D) targets, unscaled
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

template <int DP_, int EP_, int R_, int KID_, bool NORM_, int CONSUMERS_ = 256, int UNROLL_ = 2, int MINB_ = 0,
          int STAGES_ = 4>
struct DirectCfg {
    static constexpr int DP = DP_, EP = EP_, R = R_, KID = KID_;
    static constexpr bool NORM = NORM_;
    static constexpr int CONSUMERS = CONSUMERS_;
    static constexpr int UNROLL = UNROLL_;
    static constexpr int THREADS = CONSUMERS + 32;  // + one producer warp
    static constexpr int TILE_ROWS = CONSUMERS * R;
    static constexpr int STAGES = STAGES_;
    static constexpr int PAIRS = DP + EP;           // float2 per record
    static constexpr int RECV = (PAIRS + 1) / 2;    // float4 per record
    // source records per stage: ~16 KB stages whatever the record size
    static constexpr int SB = RECV <= 2 ? 512 : RECV <= 4 ? 256 : RECV <= 8 ? 128 : 64;
    static constexpr int STAGE_BYTES = SB * RECV * 16;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 2 * STAGES * 8 + 16;
    // exp-type kernels under row normalisation carry a running max (online rescale)
    static constexpr bool ONLINE_MAX = NORM && (KID != KMB_KERNEL_INVERSE_DISTANCE);
    // floats of partial state per row when a tile is split across CTAs
    static constexpr int PS = EP + (NORM ? 1 : 0) + (ONLINE_MAX ? 1 : 0);
    // resident CTAs per SM the register allocator is asked to make room for
    static constexpr int MINB = MINB_ > 0 ? MINB_ : ((!ONLINE_MAX && DP * R <= 32 && EP * R <= 16) ? 2 : 1);
    static_assert(R % 2 == 0, "rows are processed as packed pairs");
};

// owner CTA of unit u when U units are cut into G ranges [U*c/G, U*(c+1)/G)
__device__ __forceinline__ int unit_owner(long long u, long long U, int G) {
    return static_cast<int>(((u + 1) * G - 1) / U);
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
                mbar_wait(&empty_bar[stage], parity ^ 1);
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
                const float v = (ok && d < P.D) ? __ldg(P.x + row * P.D + d) * P.xscale : 0.f;
                if (r & 1) xr[d][r >> 1].y = v; else xr[d][r >> 1].x = v;
            }
            if constexpr (KID == KMB_KERNEL_INVERSE_DISTANCE) jz[r] = (P.row_offset + row) % (P.M + 1);
        }
        float2 acc[EP][RP];
        [[maybe_unused]] float2 ksum[RP];
        [[maybe_unused]] float2 kmax[RP];
#pragma unroll
        for (int p = 0; p < RP; ++p) {
#pragma unroll
            for (int e = 0; e < EP; ++e) acc[e][p] = make_float2(0.f, 0.f);
            ksum[p] = make_float2(0.f, 0.f);
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
                        float2 s;
#pragma unroll
                        for (int d = 0; d < DP; ++d) {
                            const float2 diff = add2(xr[d][p], pr[d]);
                            s = (d == 0) ? mul2(diff, diff) : fma2(diff, diff, s);
                        }
                        float2 kv = make_float2(kernel_value<KID>(s.x), kernel_value<KID>(s.y));
                        if constexpr (KID == KMB_KERNEL_INVERSE_DISTANCE) {
                            if (j_base + j == jz[2 * p]) kv.x = 0.f;
                            if (j_base + j == jz[2 * p + 1]) kv.y = 0.f;
                        }
#pragma unroll
                        for (int e = 0; e < EP; ++e) acc[e][p] = fma2(kv, pr[DP + e], acc[e][p]);
                        if constexpr (C::NORM) ksum[p] = add2(ksum[p], kv);
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
        }

        // ---------------------------- write this segment ----------------------------
        const bool complete = (cnt == nsb);
        if (complete) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const long long row = row_base + static_cast<long long>(r) * C::CONSUMERS;
                if (row < P.N) {
                    const float l = (r & 1) ? ksum[r >> 1].y : ksum[r >> 1].x;
#pragma unroll
                    for (int e = 0; e < EP; ++e) {
                        float v = (r & 1) ? acc[e][r >> 1].y : acc[e][r >> 1].x;
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
#pragma unroll
                    for (int e = 0; e < EP; ++e)
                        if (P.e0 + e < P.E) P.out[row * P.E + P.e0 + e] = C::NORM ? tot[e] / l : tot[e];
                }
            }
        }
        u += cnt;
    }
}

// ---- source packing -------------------------------------------------------------------------
// rec[j] = [-s*y_jd (dup) for d < DP | b_je (dup) for e < EP | zero pad], padding records (j >= M)
// sit far away (k underflows to exactly 0 for the exp kernels) and carry b == 0.
static __global__ void pack_sources_kernel(const float* __restrict__ y, const float* __restrict__ b, float2* __restrict__ rec,
                                    long long M, long long M_pad, int D, int E, int DP, int EP, int pairs_per_rec,
                                    int e0, float scale) {
    const long long j = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (j >= M_pad) return;
    float2* o = rec + j * pairs_per_rec;
    const bool live = j < M;
    for (int d = 0; d < DP; ++d) {
        float v = 0.f;
        if (live) v = d < D ? -scale * y[j * D + d] : 0.f;
        else v = (d == 0) ? -1.0e18f : 0.f;
        o[d] = make_float2(v, v);
    }
    for (int e = 0; e < EP; ++e) {
        float v = 0.f;
        if (live && e0 + e < E) v = b ? b[j * E + e0 + e] : 1.f;
        o[DP + e] = make_float2(v, v);
    }
    for (int q = DP + EP; q < pairs_per_rec; ++q) o[q] = make_float2(0.f, 0.f);
}


// ---- dispatch table entry ---------------------------------------------------------------------
struct DirectEntry {
    int DP, EP, KID, NORM, R, SB, RECV, PS, TILE_ROWS, SMEM, THREADS;
    const void* func;
    cudaError_t (*launch)(const DirectParams&, int grid, cudaStream_t stream);
};

template <class C>
cudaError_t launch_direct(const DirectParams& P, int grid, cudaStream_t stream) {
    static bool attr_set = false;  // per instantiation; set once per process (single device family)
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kprod_direct_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             C::SMEM_BYTES);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    kprod_direct_kernel<C><<<grid, C::THREADS, C::SMEM_BYTES, stream>>>(P);
    return cudaGetLastError();
}

template <class C>
constexpr DirectEntry make_direct_entry() {
    return DirectEntry{C::DP, C::EP, C::KID, C::NORM ? 1 : 0, C::R, C::SB, C::RECV, C::PS, C::TILE_ROWS,
                       C::SMEM_BYTES, C::THREADS, reinterpret_cast<const void*>(&kprod_direct_kernel<C>),
                       &launch_direct<C>};
}

// rows per thread: keep the register-resident targets (DP*R) and accumulators (EP*R) bounded
constexpr int direct_rows(int DP, int EP, bool online_max) {
    int r = DP <= 4 ? 8 : DP <= 8 ? 4 : 2;
    if (online_max && r > 4) r = 4;
    if (EP * r > 16) r = 16 / EP;
    return r < 2 ? 2 : r;
}

#define KMB_DIRECT_ENTRY(DP, EP, KID, NORM)                                                              \
    kmb::make_direct_entry<kmb::DirectCfg<DP, EP,                                                          \
        kmb::direct_rows(DP, EP, (NORM) && (KID) != KMB_KERNEL_INVERSE_DISTANCE), KID, NORM>>()
#define KMB_DIRECT_ENTRIES_FOR_DP(DP, KID, NORM) \
    KMB_DIRECT_ENTRY(DP, 1, KID, NORM), KMB_DIRECT_ENTRY(DP, 2, KID, NORM), KMB_DIRECT_ENTRY(DP, 4, KID, NORM)
#define KMB_DIRECT_TABLE(NAME, KID, NORM)                                                              \
    namespace kmb {                                                                                    \
    extern const DirectEntry NAME[];                                                                   \
    extern const int NAME##_count;                                                                     \
    const DirectEntry NAME[] = {KMB_DIRECT_ENTRIES_FOR_DP(2, KID, NORM), KMB_DIRECT_ENTRIES_FOR_DP(3, KID, NORM), \
                                KMB_DIRECT_ENTRIES_FOR_DP(4, KID, NORM), KMB_DIRECT_ENTRIES_FOR_DP(8, KID, NORM), \
                                KMB_DIRECT_ENTRIES_FOR_DP(16, KID, NORM)};                               \
    const int NAME##_count = sizeof(NAME) / sizeof(NAME[0]);                                           \
    }

}  // namespace kmb
