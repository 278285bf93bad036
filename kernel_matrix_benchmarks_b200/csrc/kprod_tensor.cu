// kprod_tensor: a_i = sum_j k(x_i, y_j) b_j for D > 16 on the 5th-generation tensor cores.
//
//   prepass   u = s (x - c), v = s (y - c)  (c = column means of y, s folds log2 e into the data),
//             split into TF32 hi / lo parts (3xTF32: hi.hi + hi.lo + lo.hi), squared norms in FP32.
//   main      persistent CTAs, warp-specialised:
//               warp 0      TMA producer: 128 x 32-float tiles of A = 2u (hi, lo), 256 x 32 of B = v (hi, lo),
//                           SWIZZLE_128B, 2-stage mbarrier ring of 96 KB stages
//               warp 1      MMA issuer: tcgen05.mma kind::tf32, M = 128, N = 256, K = 8 per instruction,
//                           S = 2 u.v accumulated in TMEM (two accumulator stages)
//               warps 2-5   epilogue: tcgen05.ld the S tile (one target row per thread), d2 = |u|^2 + |v|^2 - S,
//                           kernel function (MUFU), reduce against b in registers; online max-rescale for
//                           the row-normalised variant
//   work      stream-K over (128-row tile x 256-source block) units, as in the direct kernel: equal
//             contiguous unit ranges per CTA, split tiles combined in CTA order by the last CTA to arrive.
#include <algorithm>

#include "kprod_tensor.cuh"
#include "tensor_common.cuh"

namespace kmb {

namespace tc {

#ifndef KMB_TC_TN
#define KMB_TC_TN 256
#endif
constexpr int TN = KMB_TC_TN;      // sources per tile      (UMMA N, one TMEM column per source): 128 or 256
static_assert(TN == 128 || TN == 256, "TN");
constexpr int STAGES = TN == 256 ? 2 : 3;
constexpr int TILE_BYTES = TM * TK * 4;        // 16 KB: one 128 x 32-float A tile
constexpr int B_TILE_BYTES = TN * TK * 4;      // one TN x 32-float B tile
constexpr int STAGE_BYTES = 2 * TILE_BYTES + 2 * B_TILE_BYTES;    // A hi, A lo, B hi, B lo
constexpr int ACC_STAGES = 2;
constexpr int TMEM_COLS = ACC_STAGES * TN;     // 256 columns of 32-bit
constexpr int EPI_THREADS = 128;
constexpr int THREADS = 64 + EPI_THREADS;      // producer warp, MMA warp, 4 epilogue warps
constexpr int MAX_EP = 4;

struct Params {
    const float* un;     // (N) |u_i|^2
    const float* vn;     // (M) |v_j|^2
    const float* b;      // (M, E) signal or nullptr (density)
    float* out;          // (N, E)
    float* partial;
    int* tile_counter;
    long long N, M, row_offset;
    int E, e0;
    int n_tiles, nsb, kblocks;
};

template <int EP, int KID, bool NORM>
struct Cfg {
    static constexpr bool ONLINE_MAX = NORM && KID != KMB_KERNEL_INVERSE_DISTANCE;
    static constexpr int PS = EP + (NORM ? 1 : 0) + (ONLINE_MAX ? 1 : 0);
    static constexpr int AUX_FLOATS = TN * (1 + EP);   // |v|^2 and the signal chunk of one source block
    static constexpr int SMEM_BYTES = 1024 /*align slack*/ + STAGES * STAGE_BYTES + 2 * AUX_FLOATS * 4 +
                                      (2 * STAGES + 2 * ACC_STAGES) * 8 + 16;
};

template <int EP, int KID, bool NORM>
__global__ void __launch_bounds__(THREADS, 1)
kprod_tensor_kernel(const __grid_constant__ CUtensorMap map_ah, const __grid_constant__ CUtensorMap map_al,
                    const __grid_constant__ CUtensorMap map_bh, const __grid_constant__ CUtensorMap map_bl,
                    const Params P) {
    using C = Cfg<EP, KID, NORM>;
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
    unsigned char* stages = smem;
    float* aux = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(aux + 2 * C::AUX_FLOATS);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* acc_full = empty_bar + STAGES;
    uint64_t* acc_empty = acc_full + ACC_STAGES;
    uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(acc_empty + ACC_STAGES);
    int* s_flag = reinterpret_cast<int*>(tmem_base_smem + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int G = gridDim.x;
    const long long nsb = P.nsb;
    const long long U = static_cast<long long>(P.n_tiles) * nsb;
    const long long u0 = U * blockIdx.x / G, u1 = U * (blockIdx.x + 1) / G;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < ACC_STAGES; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], EPI_THREADS / 32); }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(tmem_base_smem, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_smem;

    if (warp == 0) {
        // ------------------------------------ TMA producer ------------------------------------
        // all 32 lanes walk the loop (uniform control flow); one elected lane issues (see elect_one)
        {
            uint32_t it = 0;
            for (long long u = u0; u < u1; ++u) {
                const int row0 = static_cast<int>(u / nsb) * TM;
                const int src0 = static_cast<int>(u % nsb) * TN;
                for (int kb = 0; kb < P.kblocks; ++kb, ++it) {
                    const int stage = it % STAGES;
                    mbar_wait(&empty_bar[stage], ((it / STAGES) & 1) ^ 1);
                    unsigned char* st = stages + stage * STAGE_BYTES;
                    if (elect_one()) {
                        mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
                        tma_load_2d(st + 0 * TILE_BYTES, &map_ah, kb * TK, row0, &full_bar[stage]);
                        tma_load_2d(st + 1 * TILE_BYTES, &map_al, kb * TK, row0, &full_bar[stage]);
                        tma_load_2d(st + 2 * TILE_BYTES, &map_bh, kb * TK, src0, &full_bar[stage]);
                        tma_load_2d(st + 2 * TILE_BYTES + B_TILE_BYTES, &map_bl, kb * TK, src0, &full_bar[stage]);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------- MMA issuer -------------------------------------
        // all 32 lanes walk the loop and wait on the barriers; one elected lane issues (see elect_one)
        {
            uint32_t it = 0, unit = 0;
            for (long long u = u0; u < u1; ++u, ++unit) {
                const int a = unit % ACC_STAGES;
                mbar_wait(&acc_empty[a], ((unit / ACC_STAGES) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + a * TN;
                for (int kb = 0; kb < P.kblocks; ++kb, ++it) {
                    const int stage = it % STAGES;
                    mbar_wait(&full_bar[stage], (it / STAGES) & 1);
                    tc_fence_after();
                    const unsigned char* st = stages + stage * STAGE_BYTES;
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < TK / UMMA_K; ++k) {
                            const uint64_t ah = umma_desc_sw128(st + 0 * TILE_BYTES, k * UMMA_K * 4);
                            const uint64_t al = umma_desc_sw128(st + 1 * TILE_BYTES, k * UMMA_K * 4);
                            const uint64_t bh = umma_desc_sw128(st + 2 * TILE_BYTES, k * UMMA_K * 4);
                            const uint64_t bl = umma_desc_sw128(st + 2 * TILE_BYTES + B_TILE_BYTES, k * UMMA_K * 4);
                            // 3xTF32: the two small cross terms first, then hi.hi
                            umma_tf32(d_tmem, al, bh, idesc_tf32(TN), (kb | k) != 0);
                            umma_tf32(d_tmem, ah, bl, idesc_tf32(TN), 1);
                            umma_tf32(d_tmem, ah, bh, idesc_tf32(TN), 1);
                        }
                        umma_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs have read it
                    }
                    __syncwarp();
                }
                if (elect_one()) umma_commit(&acc_full[a]);   // accumulator complete
                __syncwarp();
            }
        }
    } else {
        // -------------------------------------- epilogue --------------------------------------
        const int et = tid - 64;                       // 0..127
        const int lane_group = warp & 3;               // TMEM lanes this warp may access
        const int row_in_tile = lane_group * 32 + lane;
        uint32_t unit = 0;
        long long u = u0;
        while (u < u1) {
            const int tile = static_cast<int>(u / nsb);
            const long long sb0 = u - tile * nsb;
            const int cnt = static_cast<int>(min(nsb - sb0, u1 - u));
            const long long row = static_cast<long long>(tile) * TM + row_in_tile;
            const bool row_ok = row < P.N;
            const float un = row_ok ? __ldg(P.un + row) : 0.f;
            [[maybe_unused]] const long long jz = (P.row_offset + row) % (P.M + 1);

            float tot[EP], ktot = 0.f, kmax = -INFINITY;
#pragma unroll
            for (int e = 0; e < EP; ++e) tot[e] = 0.f;

            for (int k = 0; k < cnt; ++k, ++unit) {
                const long long j0 = (sb0 + k) * TN;
                // stage |v|^2 and the signal of this source block in shared memory (double buffered)
                float* ax = aux + (unit & 1) * C::AUX_FLOATS;
#pragma unroll
                for (int jt = et; jt < TN; jt += EPI_THREADS) {
                    const long long j = j0 + jt;
                    const bool live = j < P.M;
                    ax[jt] = live ? __ldg(P.vn + j) : 1.0e30f;   // padded sources: k underflows to 0
#pragma unroll
                    for (int e = 0; e < EP; ++e) {
                        float v = 0.f;
                        if (live && P.e0 + e < P.E) v = P.b ? __ldg(P.b + j * P.E + P.e0 + e) : 1.f;
                        ax[TN + jt * EP + e] = v;
                    }
                }
                named_bar_sync(1, EPI_THREADS);
                const int a = unit % ACC_STAGES;
                mbar_wait(&acc_full[a], (unit / ACC_STAGES) & 1);
                tc_fence_after();
                const uint32_t t_addr = tmem_base + a * TN + (static_cast<uint32_t>(lane_group * 32) << 16);

                float acc[EP], ksum = 0.f;
#pragma unroll
                for (int e = 0; e < EP; ++e) acc[e] = 0.f;
#pragma unroll 1
                for (int ch = 0; ch < TN / 32; ++ch) {
                    float s[32];
                    tmem_ld_32x32(t_addr + ch * 32, s);
                    if constexpr (!C::ONLINE_MAX) {
#pragma unroll
                        for (int c = 0; c < 32; ++c) {
                            const int jj = ch * 32 + c;
                            float kv = kernel_from_parts<KID>(s[c], un, ax[jj]);
                            if constexpr (KID == KMB_KERNEL_INVERSE_DISTANCE)
                                if (j0 + jj == jz || j0 + jj >= P.M) kv = 0.f;  // zeroing rule; padding adds exactly 0
#pragma unroll
                            for (int e = 0; e < EP; ++e) acc[e] = fmaf(kv, ax[TN + jj * EP + e], acc[e]);
                            if constexpr (NORM) ksum += kv;
                        }
                    } else {
                        float cm = -INFINITY;
#pragma unroll
                        for (int c = 0; c < 32; ++c) {
                            s[c] = log2_kernel_from_parts<KID>(s[c], un, ax[ch * 32 + c]);
                            cm = fmaxf(cm, s[c]);
                        }
                        const float mnew = fmaxf(kmax, cm);
                        const float sc = (mnew == -INFINITY) ? 1.f : ex2_approx(kmax - mnew);
                        const float moff = (mnew == -INFINITY) ? 0.f : mnew;
                        kmax = mnew;
                        // rescale everything accumulated so far (this unit and the tile totals)
                        ksum *= sc;
                        ktot *= sc;
#pragma unroll
                        for (int e = 0; e < EP; ++e) { acc[e] *= sc; tot[e] *= sc; }
#pragma unroll
                        for (int c = 0; c < 32; ++c) {
                            const int jj = ch * 32 + c;
                            const float kv = ex2_approx(s[c] - moff);
#pragma unroll
                            for (int e = 0; e < EP; ++e) acc[e] = fmaf(kv, ax[TN + jj * EP + e], acc[e]);
                            ksum += kv;
                        }
                    }
                }
                // accumulator drained: hand the TMEM stage back to the MMA warp
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[a]);
                // two-level summation (see kprod_direct.cuh)
#pragma unroll
                for (int e = 0; e < EP; ++e) tot[e] += acc[e];
                ktot += ksum;
            }

            // ------------------------------ write this segment ------------------------------
            if (cnt == nsb) {
                if (row_ok) {
#pragma unroll
                    for (int e = 0; e < EP; ++e)
                        if (P.e0 + e < P.E) P.out[row * P.E + P.e0 + e] = NORM ? tot[e] / ktot : tot[e];
                }
            } else {
                const int slot = (u == u0) ? 0 : 1;
                float* mine = P.partial + (static_cast<size_t>(blockIdx.x) * 2 + slot) * (TM * C::PS);
#pragma unroll
                for (int e = 0; e < EP; ++e) mine[e * TM + row_in_tile] = tot[e];
                if constexpr (NORM) mine[EP * TM + row_in_tile] = ktot;
                if constexpr (C::ONLINE_MAX) mine[(EP + 1) * TM + row_in_tile] = kmax;
                __threadfence();
                named_bar_sync(2, EPI_THREADS);
                const long long tile_u0 = static_cast<long long>(tile) * nsb;
                const int c_first = static_cast<int>(((tile_u0 + 1) * G - 1) / U);
                const int c_last = static_cast<int>(((tile_u0 + nsb) * G - 1) / U);
                if (et == 0) {
                    const int old = atomicAdd(&P.tile_counter[tile], 1);
                    const int last = (old == c_last - c_first);
                    if (last) P.tile_counter[tile] = 0;
                    *s_flag = last;
                }
                named_bar_sync(2, EPI_THREADS);
                const bool is_last = *s_flag != 0;
                named_bar_sync(2, EPI_THREADS);
                if (is_last && row_ok) {
                    __threadfence();
                    float sum[EP], l = 0.f, mx = -INFINITY;
#pragma unroll
                    for (int e = 0; e < EP; ++e) sum[e] = 0.f;
                    if constexpr (C::ONLINE_MAX) {
                        for (int c = c_first; c <= c_last; ++c) {
                            const int sl = (U * c / G) / nsb == tile ? 0 : 1;
                            const float* ps = P.partial + (static_cast<size_t>(c) * 2 + sl) * (TM * C::PS);
                            mx = fmaxf(mx, __ldcg(ps + (EP + 1) * TM + row_in_tile));
                        }
                    }
                    for (int c = c_first; c <= c_last; ++c) {
                        const int sl = (U * c / G) / nsb == tile ? 0 : 1;
                        const float* ps = P.partial + (static_cast<size_t>(c) * 2 + sl) * (TM * C::PS);
                        float w = 1.f;
                        if constexpr (C::ONLINE_MAX) {
                            const float m = __ldcg(ps + (EP + 1) * TM + row_in_tile);
                            w = (m == -INFINITY) ? 0.f : ex2_approx(m - mx);
                        }
#pragma unroll
                        for (int e = 0; e < EP; ++e) sum[e] = fmaf(w, __ldcg(ps + e * TM + row_in_tile), sum[e]);
                        if constexpr (NORM) l = fmaf(w, __ldcg(ps + EP * TM + row_in_tile), l);
                    }
#pragma unroll
                    for (int e = 0; e < EP; ++e)
                        if (P.e0 + e < P.E) P.out[row * P.E + P.e0 + e] = NORM ? sum[e] / l : sum[e];
                }
            }
            u += cnt;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ---- prepass ---------------------------------------------------------------------------------------

// partial[blk][col] = sum over this block's rows of y[row][col]
static __global__ void __launch_bounds__(256) column_sum_kernel(const float* __restrict__ y, long long M, int D,
                                                                float* __restrict__ partial) {
    __shared__ float sm[8][32];
    const int col = blockIdx.y * 32 + (threadIdx.x & 31);
    const int rl = threadIdx.x >> 5;
    float acc = 0.f;
    if (col < D)
        for (long long r = blockIdx.x * 8 + rl; r < M; r += static_cast<long long>(gridDim.x) * 8) acc += y[r * D + col];
    sm[rl][threadIdx.x & 31] = acc;
    __syncthreads();
    if (rl == 0 && col < D) {
        float t = 0.f;
        for (int i = 0; i < 8; ++i) t += sm[i][threadIdx.x];
        partial[static_cast<size_t>(blockIdx.x) * D + col] = t;
    }
}
static __global__ void column_mean_kernel(const float* __restrict__ partial, int blocks, long long M, int D, int Dp,
                                          float* __restrict__ center) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= Dp) return;
    float t = 0.f;
    if (col < D)
        for (int b = 0; b < blocks; ++b) t += partial[static_cast<size_t>(b) * D + col];
    center[col] = col < D ? t / static_cast<float>(M) : 0.f;
}

// One warp per point: w = mult * s * (p - c); hi = tf32(w), lo = tf32(w - hi); norm2 = |s (p - c)|^2.
static __global__ void __launch_bounds__(256) split_points_kernel(const float* __restrict__ pts, long long n, int D, int Dp,
                                                                  const float* __restrict__ center, float scale, float mult,
                                                                  float* __restrict__ hi, float* __restrict__ lo,
                                                                  float* __restrict__ norm2) {
    const long long row = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    float acc = 0.f;
    for (int d = lane; d < Dp; d += 32) {
        float w = 0.f;
        if (d < D) w = scale * (pts[row * D + d] - center[d]);
        acc = fmaf(w, w, acc);
        w *= mult;
        const float h = to_tf32(w);
        hi[row * Dp + d] = h;
        lo[row * Dp + d] = to_tf32(w - h);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) norm2[row] = acc;
}

// (rows, cols) fp32 row-major; box = 32 floats x box_rows rows, 128-byte swizzle, out-of-bounds reads as zero
int make_tensor_map(CUtensorMap* map, const float* base, long long rows, int cols, int box_rows) {
    using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn enc = nullptr;
    if (!enc) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        KMB_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        if (!p || q != cudaDriverEntryPointSuccess) return set_error(KMB_ERR_CUDA, "cuTensorMapEncodeTiled not available");
        enc = reinterpret_cast<EncodeFn>(p);
    }
    const cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
    const cuuint64_t strides[1] = {static_cast<cuuint64_t>(cols) * 4};
    const cuuint32_t box[2] = {static_cast<cuuint32_t>(TK), static_cast<cuuint32_t>(box_rows)};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(KMB_ERR_CUDA, "cuTensorMapEncodeTiled failed with %d", static_cast<int>(r));
    return KMB_OK;
}

int tensor_prepass(const float* x, const float* y, int64_t N, int64_t M, int D, int Dp, int kid, float* center, float* cpart,
                   float* uh, float* ul, float* vh, float* vl, float* un, float* vn, cudaStream_t stream) {
    // scale folds log2(e) into the data as in the direct path; the A operand carries the factor 2 of 2 u.v
    const float scale = kid == KMB_KERNEL_GAUSSIAN ? 1.2011224087864498f : kid == KMB_KERNEL_ABSOLUTE_EXPONENTIAL ? 1.4426950408889634f : 1.f;
    const int cblocks = static_cast<int>(std::min<long long>(CENTER_BLOCKS, (M + 7) / 8));
    dim3 g(cblocks, (D + 31) / 32);
    column_sum_kernel<<<g, 256, 0, stream>>>(y, M, D, cpart);
    KMB_CUDA_CHECK(cudaGetLastError());
    column_mean_kernel<<<(Dp + 127) / 128, 128, 0, stream>>>(cpart, cblocks, M, D, Dp, center);
    KMB_CUDA_CHECK(cudaGetLastError());
    split_points_kernel<<<static_cast<unsigned>((N * 32 + 255) / 256), 256, 0, stream>>>(x, N, D, Dp, center, scale, 2.f, uh, ul, un);
    KMB_CUDA_CHECK(cudaGetLastError());
    split_points_kernel<<<static_cast<unsigned>((M * 32 + 255) / 256), 256, 0, stream>>>(y, M, D, Dp, center, scale, 1.f, vh, vl, vn);
    KMB_CUDA_CHECK(cudaGetLastError());
    count_launch(4);
    return KMB_OK;
}

}  // namespace tc

// ---- host side ---------------------------------------------------------------------------------------
namespace {

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct TensorPlan {
    int Dp, e_chunk, n_passes, grid_max;
    long long n_tiles, nsb;
    size_t off_center, off_cpart, off_uh, off_ul, off_vh, off_vl, off_un, off_vn, off_partial, off_counter, total;
};

int plan_tensor(int64_t N, int64_t M, int D, int E, int flags, TensorPlan* pl) {
    pl->Dp = (D + tc::TK - 1) / tc::TK * tc::TK;
    pl->e_chunk = E >= 4 ? 4 : E;
    pl->n_passes = (E + pl->e_chunk - 1) / pl->e_chunk;
    pl->n_tiles = (N + tc::TM - 1) / tc::TM;
    pl->nsb = (M + tc::TN - 1) / tc::TN;
    int dev = 0, sms = 0;
    KMB_CUDA_CHECK(cudaGetDevice(&dev));
    KMB_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    pl->grid_max = sms;
    const int PS = tc::MAX_EP + 2;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t at = o; o += align_up(bytes, 256); return at; };
    pl->off_center = take(sizeof(float) * pl->Dp);
    pl->off_cpart = take(sizeof(float) * tc::CENTER_BLOCKS * D);
    pl->off_uh = take(sizeof(float) * N * pl->Dp);
    pl->off_ul = take(sizeof(float) * N * pl->Dp);
    pl->off_vh = take(sizeof(float) * M * pl->Dp);
    pl->off_vl = take(sizeof(float) * M * pl->Dp);
    pl->off_un = take(sizeof(float) * N);
    pl->off_vn = take(sizeof(float) * M);
    pl->off_partial = take(sizeof(float) * pl->grid_max * 2 * tc::TM * PS);
    pl->off_counter = take(sizeof(int) * pl->n_tiles);
    pl->total = o;
    (void)flags;
    return KMB_OK;
}

using KernelFn = void (*)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const tc::Params);

template <int EP, int KID, bool NORM>
int launch_one(const CUtensorMap* maps, const tc::Params& P, int grid, cudaStream_t stream) {
    using C = tc::Cfg<EP, KID, NORM>;
    auto fn = tc::kprod_tensor_kernel<EP, KID, NORM>;
    static bool attr = false;
    if (!attr) {
        KMB_CUDA_CHECK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        attr = true;
    }
    fn<<<grid, tc::THREADS, C::SMEM_BYTES, stream>>>(maps[0], maps[1], maps[2], maps[3], P);
    KMB_CUDA_CHECK(cudaGetLastError());
    return KMB_OK;
}

template <int KID, bool NORM>
int launch_ep(int ep, const CUtensorMap* maps, const tc::Params& P, int grid, cudaStream_t stream) {
    if (ep == 1) return launch_one<1, KID, NORM>(maps, P, grid, stream);
    if (ep == 2) return launch_one<2, KID, NORM>(maps, P, grid, stream);
    return launch_one<4, KID, NORM>(maps, P, grid, stream);
}

int launch_any(int kid, bool norm, int ep, const CUtensorMap* maps, const tc::Params& P, int grid, cudaStream_t stream) {
    switch (kid * 2 + (norm ? 1 : 0)) {
        case 0: return launch_ep<KMB_KERNEL_GAUSSIAN, false>(ep, maps, P, grid, stream);
        case 1: return launch_ep<KMB_KERNEL_GAUSSIAN, true>(ep, maps, P, grid, stream);
        case 2: return launch_ep<KMB_KERNEL_ABSOLUTE_EXPONENTIAL, false>(ep, maps, P, grid, stream);
        case 3: return launch_ep<KMB_KERNEL_ABSOLUTE_EXPONENTIAL, true>(ep, maps, P, grid, stream);
        case 4: return launch_ep<KMB_KERNEL_INVERSE_DISTANCE, false>(ep, maps, P, grid, stream);
        default: return launch_ep<KMB_KERNEL_INVERSE_DISTANCE, true>(ep, maps, P, grid, stream);
    }
}

}  // namespace

int tensor_workspace_bytes(int64_t N, int64_t M, int D, int E, int kid, int flags, size_t* bytes) {
    (void)kid;
    if (tensor_pv_applicable(D, E)) return tensor_pv_workspace_bytes(N, M, D, E, bytes);
    TensorPlan pl{};
    if (int rc = plan_tensor(N, M, D, E, flags, &pl)) return rc;
    *bytes = pl.total;
    return KMB_OK;
}

int tensor_product(const float* x, const float* y, const float* b, float* out, int64_t N, int64_t M, int D, int E,
                   int kid, int flags, int64_t row_offset, void* workspace, size_t workspace_bytes, cudaStream_t stream,
                   cudaEvent_t ev0, cudaEvent_t ev1) {
    if (tensor_pv_applicable(D, E))
        return tensor_pv_product(x, y, b, out, N, M, D, E, kid, flags, row_offset, workspace, workspace_bytes, stream, ev0, ev1);
    TensorPlan pl{};
    if (int rc = plan_tensor(N, M, D, E, flags, &pl)) return rc;
    if (!workspace || workspace_bytes < pl.total)
        return set_error(KMB_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", pl.total, workspace_bytes);
    if (N >= (1ll << 31) - tc::TM || M >= (1ll << 31) - tc::TN)
        return set_error(KMB_ERR_UNSUPPORTED, "tensor path indexes rows with 32-bit TMA coordinates");
    char* ws = static_cast<char*>(workspace);
    float* center = reinterpret_cast<float*>(ws + pl.off_center);
    float* cpart = reinterpret_cast<float*>(ws + pl.off_cpart);
    float* uh = reinterpret_cast<float*>(ws + pl.off_uh);
    float* ul = reinterpret_cast<float*>(ws + pl.off_ul);
    float* vh = reinterpret_cast<float*>(ws + pl.off_vh);
    float* vl = reinterpret_cast<float*>(ws + pl.off_vl);
    float* un = reinterpret_cast<float*>(ws + pl.off_un);
    float* vn = reinterpret_cast<float*>(ws + pl.off_vn);
    float* partial = reinterpret_cast<float*>(ws + pl.off_partial);
    int* counters = reinterpret_cast<int*>(ws + pl.off_counter);
    const bool norm = flags & KMB_FLAG_NORMALIZE_ROWS;
    const bool density = flags & KMB_FLAG_DENSITY;

    KMB_CUDA_CHECK(cudaMemsetAsync(counters, 0, sizeof(int) * pl.n_tiles, stream));
    if (int rc = tc::tensor_prepass(x, y, N, M, D, pl.Dp, kid, center, cpart, uh, ul, vh, vl, un, vn, stream)) return rc;
    CUtensorMap maps[4];
    if (int rc = tc::make_tensor_map(&maps[0], uh, N, pl.Dp, tc::TM)) return rc;
    if (int rc = tc::make_tensor_map(&maps[1], ul, N, pl.Dp, tc::TM)) return rc;
    if (int rc = tc::make_tensor_map(&maps[2], vh, M, pl.Dp, tc::TN)) return rc;
    if (int rc = tc::make_tensor_map(&maps[3], vl, M, pl.Dp, tc::TN)) return rc;

    const long long units = pl.n_tiles * pl.nsb;
    const int grid = static_cast<int>(std::min<long long>(pl.grid_max, units));
    for (int pass = 0; pass < pl.n_passes; ++pass) {
        tc::Params P;
        P.un = un;
        P.vn = vn;
        P.b = density ? nullptr : b;
        P.out = out;
        P.partial = partial;
        P.tile_counter = counters;
        P.N = N;
        P.M = M;
        P.row_offset = row_offset;
        P.E = E;
        P.e0 = pass * pl.e_chunk;
        P.n_tiles = static_cast<int>(pl.n_tiles);
        P.nsb = static_cast<int>(pl.nsb);
        P.kblocks = pl.Dp / tc::TK;
        if (ev0 && pass == pl.n_passes - 1) KMB_CUDA_CHECK(cudaEventRecord(ev0, stream));
        if (int rc = launch_any(kid, norm, pl.e_chunk, maps, P, grid, stream)) return rc;
        if (ev1 && pass == pl.n_passes - 1) KMB_CUDA_CHECK(cudaEventRecord(ev1, stream));
        count_launch();
    }
    return KMB_OK;
}

}  // namespace kmb
