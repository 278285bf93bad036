// kprod_tensor_pv16: a_i = sum_j k(x_i, y_j) b_j for 16 < D <= 128 and E > 4, Gaussian and exponential kernels --
// both contractions on the tensor cores with FP16 hi / lo operand planes (config C4: exponential-kernel
// attention, N = M = 262144, D = E = 64).  Same algorithm as kprod_tensor_pv.cu (TF32 planes), re-shaped after
// tools/ubench_umma.cu: a 128 x 64 x 8 SS MMA takes 60 cycles for 32 cycles of math, a 128 x 128 x 16 one 75 for
// 64, and kind::f16 does twice the work of kind::tf32 per instruction.
//
//   S = 2 u.v^T   128 rows x 128 sources per block; tcgen05.mma kind::f16, three-term split
//                 (lo.hi + hi.lo + hi.hi), A = u tile (hi, lo) resident in shared memory for the whole row
//                 tile, B = v blocks streamed by TMA, FP32 accumulator in TMEM (two stages)
//   P = k(S)      16 epilogue warps = 4 column groups x 4 TMEM lane quarters.  Group g owns columns [32g, 32g+32)
//                 of every S block and runs its OWN online-softmax stream over those sources: tcgen05.ld S, log2
//                 of the kernel, a running reference exponent per (row, group) (lazy rescale of the group's O
//                 when the maximum outgrows it by 2^8, so P <= 2^8 fits FP16 and rows whose kernel values all
//                 underflow FP32 still normalise), P = 2^(log2 k - ref) split into FP16 hi + lo, packed two per
//                 32-bit column and stored IN PLACE over the thread's own S columns (tcgen05.st).  No
//                 block-level synchronisation between epilogue warps inside a row tile.
//   O_g += P_g.B  tcgen05.mma with A = P from TMEM (hi, lo), B = transposed signal block (FP16 hi, lo, scaled per
//                 signal column by a power of two) from shared memory; one accumulator O_g (128 x E) per column
//                 group, kept in TMEM for the whole row tile; the four are merged (weights 2^(ref_g - ref)) when
//                 the row tile ends -- every thread can read all four because TMEM lanes are rows.
//
// TMEM columns: S/P stage 0 [0,128) | S/P stage 1 [128,256) | O_0 .. O_3 at 256 + 64 g.
// S(n+2) overwrites stage n & 1 after PV(n) has been issued (tensor-pipe order), so P is double buffered for free.
// Work split: the wave schedule of kprod_tensor.cu -- every CTA of a wave walks the SAME source blocks at the same
// time (all of them when there are at least as many row tiles as CTAs), so v and b blocks come from L2.
#include <algorithm>
#include <cstdlib>

#include <cuda_fp16.h>

#include "tensor_common.cuh"

namespace kmb {
namespace pv16 {

using namespace tc;

constexpr int TNS = 128;               // sources per S block
constexpr int SLOT_BYTES = 32768;      // ring slot: v block of one K block (hi 16 KB | lo 16 KB) or one signal block
                                       // (CTA pairs: each CTA holds half of either, 16 KB slots)
constexpr int A_TILE_BYTES = TM * 128; // 16 KB: 128 rows of one 128-byte K block
constexpr int PANEL_BYTES = 64 * 128;  // signal: 64 signal columns x 64 sources (one swizzle atom wide)
constexpr int NG = 4;                  // epilogue column groups (4 warps each)
constexpr int CPT = TNS / NG;          // S columns per epilogue thread (32)
constexpr int EPI_WARPS = 4 * NG;
constexpr int EPI_THREADS = 32 * EPI_WARPS;
constexpr int THREADS = 64 + EPI_THREADS;
constexpr int TMEM_COLS = 512;
constexpr int COL_S = 0, COL_O = 256;
constexpr int MAX_EB = 64;             // signal columns per pass
constexpr float kLazyRescale = 8.f;    // rescale O only when the row maximum outgrew the reference by 2^8
constexpr int PS = MAX_EB + 2;         // partial record: O row, sum of weights, reference exponent
// The tensor cores add into the FP32 accumulator with truncation, not round-to-nearest: every accumulating MMA shrinks
// O_g by ~2^-25 of its magnitude, and a row tile at M = 262144 makes 12288 of them per group -- measured 2.4e-4 relative
// (every row alike, the sum of weights on the CUDA cores does not shrink with it) against the 1e-4 the path is held to.
// So O_g only collects kFlushBlocks source blocks; then the epilogue adds it (round-to-nearest, L2 reductions the thread
// does not wait for) to a per-(CTA, group) FP32 accumulator in global memory and the next P.B starts from zero.
// Measured at C4 (M = 262144): flush every 2048 blocks (= never) 2.4e-4, every 32 blocks 3.6e-6 at +6 % time (the L2
// reductions of 148 CTAs arrive together).
// Build-time variants, A/B-tested on one box with tools/pv16_ab.sh (make EXTRA="-DKMB_PV16_...=..."):
#ifndef KMB_PV16_FUSED
#define KMB_PV16_FUSED 0     // 1: one fused pass per block instead of two phases (t, then P)
#endif
#ifndef KMB_PV16_SETS
#define KMB_PV16_SETS 0      // 1: the 16 epilogue warps form two sets of eight that work on ALTERNATE source blocks (see the epilogue)
#endif
#ifndef KMB_PV16_VN_LDG
#define KMB_PV16_VN_LDG 0    // 1: |v|^2 of a block straight from global memory (eight uniform LDG.128 per thread, L1 broadcast) instead of
                             // one coalesced load parked in a shared-memory line: STS / LDS share the MIO queue with the MUFU instructions
#endif
#ifndef KMB_PV16_SKEW_NS
#define KMB_PV16_SKEW_NS 0   // > 0: column group g starts every row tile g * KMB_PV16_SKEW_NS ns late, so that the four epilogue
                             // warps of an SM sub-partition (one per group) are not all between their MUFU phases at once
#endif
constexpr int kFlushBlocks = 128;      // default of Params::flush_blocks: 768 accumulating MMAs per flush, ~1.5e-5

struct Params {
    const float* un;
    const float* vn;
    const float* sscale;       // [1] = 2^-2p: S = sscale[1] * accumulator
    const float* binv;         // (Ep) 2^-q_e: undoes the per-column scale of the signal planes
    float* out;
    float* partial;
    float* olong;              // grid x NG x MAX_EB x TM: long accumulators of the O_g, [column][row] (see kFlushBlocks)
    int* tile_counter;
    long long N, M;
    int E, e0, eb, ebp;        // this pass covers signal columns e0 .. e0+eb-1; ebp = eb rounded up to 32
    int n_tiles, nsb, kblocks, ksteps_last, stages, ep_rows, flush_blocks;
    int R, C, W, R_last, C_last, slots_per_wave;
};

// `unit`: the CTA (single-CTA kernel) or the cluster (CTA pairs: the wave plan then counts pairs of row tiles)
struct WaveWork { int tile, sb_lo, sb_hi, c, Cw, tile_in_wave; };
__device__ __forceinline__ bool wave_work(const Params& P, int w, int cta, WaveWork& ww) {
    const bool last = (w == P.W - 1);
    const int Rw = last ? P.R_last : P.R, Cw = last ? P.C_last : P.C;
    if (cta >= Rw * Cw) return false;
    ww.Cw = Cw;
    ww.tile_in_wave = cta / Cw;
    ww.c = cta - ww.tile_in_wave * Cw;
    ww.tile = w * P.R + ww.tile_in_wave;
    ww.sb_lo = static_cast<int>(static_cast<long long>(P.nsb) * ww.c / Cw);
    ww.sb_hi = static_cast<int>(static_cast<long long>(P.nsb) * (ww.c + 1) / Cw);
    return true;
}

__device__ __forceinline__ void umma_f16_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d),
                 "l"(a), "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void umma_f16_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d),
                 "r"(a), "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
}
__host__ __device__ constexpr uint32_t idesc_f16(int n) {   // D = F32, A = B = F16, K-major
    return (1u << 4) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(TM >> 4) << 24);
}
// two floats -> one 32-bit word of two halves: `even` in bits [0,16), `odd` in bits [16,32)
__device__ __forceinline__ uint32_t pack_half2(float even, float odd) {
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(odd), "f"(even));
    return r;
}

// t = -log2 k for two sources at once (packed FP32: FFMA2 / FADD2 / FMUL2), from the raw accumulators:
// s_raw sscale = 2 u.v on log2-scaled data, w = |u|^2 + |v|^2
template <int KID>
__device__ __forceinline__ float2 neg_log2_kernel2(float2 s_raw, float2 nsscale, float2 w) {
    const float2 d2 = fma2(s_raw, nsscale, w);          // log2(e) |x - y|^2 (Gaussian) or (log2(e) |x - y|)^2
    if constexpr (KID == KMB_KERNEL_GAUSSIAN) return d2;
    else {
        // bruteforce.py:21: sqrt(maximum(sqdists, 0)); the clamp away from zero keeps rsqrt finite (see sqrt_approx).
        // d2 is finite: |v|^2 of padded sources is 3.39e38 and |2 u.v| is far below the 1.3e36 left to FLT_MAX.
        const float2 c = make_float2(fmaxf(d2.x, 1.0e-30f), fmaxf(d2.y, 1.0e-30f));
        return mul2(c, make_float2(rsqrt_approx(c.x), rsqrt_approx(c.y)));
    }
}

// 2^d for an integer-valued d <= 0 (exact; 0 below the normal range, which is also what d = -inf gives)
__device__ __forceinline__ float pow2_int(float d) {
    return d > -126.f ? __int_as_float((127 + static_cast<int>(d)) << 23) : 0.f;
}

// PAIR: two CTAs of a cluster (cta_group::2) work on two adjacent row tiles and the same source blocks: every MMA is
// 256 rows tall (half as many instructions per row tile -- the issuing warp is what bounds the single-CTA kernel),
// each CTA streams half of every v block (64 sources) and half of every signal block (32 signal columns), the leader
// CTA issues, tcgen05.commit multicasts to both CTAs' barriers and the peer's epilogue warps arrive on the leader's.
template <int KID, bool NORM, bool PAIR>
__device__ __forceinline__ void pv16_body(const CUtensorMap& map_ah, const CUtensorMap& map_al, const CUtensorMap& map_bh,
                                          const CUtensorMap& map_bl, const CUtensorMap& map_sh, const CUtensorMap& map_sl,
                                          const Params& P) {
    constexpr int SLOT = PAIR ? SLOT_BYTES / 2 : SLOT_BYTES;       // V: hi | lo halves of the slot
    constexpr int PANEL = PAIR ? PANEL_BYTES / 2 : PANEL_BYTES;    // signal: 4 panels (hi 0, hi 1, lo 0, lo 1)
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
    unsigned char* u_region = smem;                                   // kblocks x [A hi 16 KB | A lo 16 KB]
    unsigned char* ring = u_region + P.kblocks * 2 * A_TILE_BYTES;    // stages x 32 KB
    float* vline = reinterpret_cast<float*>(ring + P.stages * SLOT);    // EPI_WARPS x 2 x 32: per-warp |v|^2 lines
    float* refbuf = vline + EPI_WARPS * 2 * 64;                                // NG x TM: per-group reference exponents
    float* ksbuf = refbuf + NG * TM;                                           // NG x TM: per-group sums of weights
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(ksbuf + NG * TM);
    uint64_t* empty_bar = full_bar + P.stages;
    uint64_t* acc_full = empty_bar + P.stages;     // [2] S(n) complete in stage n & 1
    uint64_t* p_ready = acc_full + 2;              // [2] P(n) stored in stage n & 1
    uint64_t* pv_done = p_ready + 2;               // [2] PV(n) complete
    uint64_t* u_full = pv_done + 2;
    uint64_t* u_free = u_full + 1;
    uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(u_free + 1);
    int* s_flag = reinterpret_cast<int*>(tmem_base_smem + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = PAIR ? pair::cluster_ctarank() : 0;   // 0 = leader (issues the MMAs)
    const int cta = PAIR ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);   // unit of the wave plan
    const int ST = P.stages;
    constexpr int NCTA = PAIR ? 2 : 1;

    if (tid == 0) {
        for (int s = 0; s < ST; ++s) { mbar_init(&full_bar[s], NCTA); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&p_ready[a], NCTA * (KMB_PV16_SETS ? EPI_WARPS / 2 : EPI_WARPS)); mbar_init(&pv_done[a], 1); }
        mbar_init(u_full, NCTA);
        mbar_init(u_free, 1);
        fence_mbar_init();
    }
    if (warp == 1) {
        if constexpr (PAIR) pair::tmem_alloc2(tmem_base_smem, TMEM_COLS);
        else tmem_alloc(tmem_base_smem, TMEM_COLS);
    }
    tc_fence_before();
    __syncthreads();
    if constexpr (PAIR) pair::cluster_sync_all();   // the peer's barriers exist before anything arrives on them
    tc_fence_after();
    // barrier-side helpers: where TMA bytes are counted / how the producers arm a full barrier
    auto arm_full = [&](uint64_t* bar, uint32_t bytes_per_cta) {   // elected lane only
        if constexpr (PAIR) {
            if (rank == 0) mbar_arrive_expect_tx(bar, 2 * bytes_per_cta);
            else pair::mbar_arrive_cluster(pair::map_to_cta(bar, 0));
        } else {
            mbar_arrive_expect_tx(bar, bytes_per_cta);
        }
    };
    auto load2d = [&](void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
        if constexpr (PAIR) pair::tma_load_2d_pair(dst, map, c0, c1, pair::map_to_cta(bar, 0));
        else tma_load_2d(dst, map, c0, c1, bar);
    };
    auto commit = [&](uint64_t* bar) {
        if constexpr (PAIR) pair::umma2_commit_both(bar);
        else umma_commit(bar);
    };
    const uint32_t tmem_base = *tmem_base_smem;

    if (warp == 0) {
        // ------------------------------------ TMA producer ------------------------------------
        // all 32 lanes walk the loop (uniform control flow); one elected lane issues (see elect_one)
        uint32_t it = 0, seg = 0;
        auto emit_signal = [&](int sb) {   // one slot: hi panels 0, 1 | lo panels 0, 1
            // slab sb, signal column e0 (CTA pairs: this CTA's half of the pass's signal columns)
            const int row0 = sb * P.ep_rows + P.e0 + (PAIR ? static_cast<int>(rank) * (P.ebp / 2) : 0);
            const int slot = it % ST;
            mbar_wait(&empty_bar[slot], ((it / ST) & 1) ^ 1);
            unsigned char* dst = ring + slot * SLOT;
            if (elect_one()) {
                arm_full(&full_bar[slot], PAIR ? 4u * (P.ebp / 2) * 128u : 4u * PANEL_BYTES);
                load2d(dst + 0 * PANEL, &map_sh, 0, row0, &full_bar[slot]);
                load2d(dst + 1 * PANEL, &map_sh, 64, row0, &full_bar[slot]);
                load2d(dst + 2 * PANEL, &map_sl, 0, row0, &full_bar[slot]);
                load2d(dst + 3 * PANEL, &map_sl, 64, row0, &full_bar[slot]);
            }
            __syncwarp();
            ++it;
        };
        int prev = -1;
        WaveWork ww;
        for (int w = 0; w < P.W; ++w) {
            if (!wave_work(P, w, cta, ww)) continue;
            // new row tile: (re)load the resident u tile once the last S of the previous tile has read it
            mbar_wait(u_free, (seg & 1) ^ 1);
            const int my_row0 = (PAIR ? ww.tile * 2 + static_cast<int>(rank) : ww.tile) * TM;
            if (elect_one()) {
                arm_full(u_full, P.kblocks * 2 * A_TILE_BYTES);
                for (int kb = 0; kb < P.kblocks; ++kb) {
                    load2d(u_region + (kb * 2 + 0) * A_TILE_BYTES, &map_ah, kb * 64, my_row0, u_full);
                    load2d(u_region + (kb * 2 + 1) * A_TILE_BYTES, &map_al, kb * 64, my_row0, u_full);
                }
            }
            __syncwarp();
            ++seg;
            for (int sb = ww.sb_lo; sb < ww.sb_hi; ++sb) {
                for (int kb = 0; kb < P.kblocks; ++kb, ++it) {
                    const int slot = it % ST;
                    mbar_wait(&empty_bar[slot], ((it / ST) & 1) ^ 1);
                    unsigned char* dst = ring + slot * SLOT;
                    const int src0 = sb * TNS + (PAIR ? static_cast<int>(rank) * (TNS / 2) : 0);   // pairs: this CTA's 64 sources
                    if (elect_one()) {
                        arm_full(&full_bar[slot], SLOT);
                        load2d(dst, &map_bh, kb * 64, src0, &full_bar[slot]);
                        load2d(dst + SLOT / 2, &map_bl, kb * 64, src0, &full_bar[slot]);
                    }
                    __syncwarp();
                }
                if (prev >= 0) emit_signal(prev);   // consumed by PV(n-1), issued after S(n)
                prev = sb;
            }
        }
        if (prev >= 0) emit_signal(prev);
    } else if (warp == 1 && rank == 0) {
        // ------------------------------------- MMA issuer (pairs: the leader CTA's) -------------------------------------
        // All 32 lanes walk the loop and wait on the barriers; one elected lane issues (see elect_one).
        // Order: S(0), S(1), PV(0), S(2), PV(1), ...
        uint32_t it = 0, n = 0, seg = 0;
        const uint32_t d_o = tmem_base + COL_O;
        constexpr uint32_t m_bits = PAIR ? (static_cast<uint32_t>(256 >> 4) << 24) : (static_cast<uint32_t>(TM >> 4) << 24);
        const uint32_t idesc_s = (1u << 4) | (static_cast<uint32_t>(TNS >> 3) << 17) | m_bits;
        const uint32_t idesc_o = (1u << 4) | (static_cast<uint32_t>(P.ebp >> 3) << 17) | m_bits;
        auto mma_ss = [&](uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
            if constexpr (PAIR) pair::umma2_f16(d, a, b, idesc, acc);
            else umma_f16_ss(d, a, b, idesc, acc);
        };
        auto mma_ts = [&](uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
            if constexpr (PAIR) pair::umma2_f16_ts(d, a, b, idesc, acc);
            else umma_f16_ts(d, a, b, idesc, acc);
        };
#ifdef KMB_PV16_TIMING
        long long macc[6] = {0, 0, 0, 0, 0, 0}, mprev = clock64();
#define KMB_M(i) do { const long long t_ = clock64(); macc[i] += t_ - mprev; mprev = t_; } while (0)
#else
#define KMB_M(i) do { } while (0)
#endif
        auto issue_pv = [&](uint32_t m, bool from_zero) {   // from_zero: the O_g were flushed (or the tile starts)
            const int a = m & 1;
            KMB_M(0);
            mbar_wait(&p_ready[a], (m >> 1) & 1);
            KMB_M(1);
            const int slot = it % ST;
            mbar_wait(&full_bar[slot], (it / ST) & 1);
            ++it;
            tc_fence_after();
            KMB_M(2);
            const unsigned char* sg = ring + slot * SLOT;
            const uint32_t p_base = tmem_base + COL_S + a * TNS;
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < TNS / 16; ++k) {   // 16 sources per instruction; 32-column chunk g = k / 2
                    const int g = k >> 1, panel = k >> 2, koff = (k & 3) * 32;
                    const uint64_t bh = umma_desc_sw128(sg + panel * PANEL, koff);
                    const uint64_t bl = umma_desc_sw128(sg + (2 + panel) * PANEL, koff);
                    const uint32_t a_hi = p_base + g * CPT + (k & 1) * 8, a_lo = a_hi + CPT / 2;
#if KMB_PV16_SETS
                    // stream = set (m & 1) + 2 * column half (k / 4): both 32-column chunks of a half go to one accumulator
                    const uint32_t d_g = d_o + ((m & 1) + 2 * (k >> 2)) * MAX_EB;
                    mma_ts(d_g, a_lo, bh, idesc_o, !(from_zero && (k & 3) == 0));
#else
                    const uint32_t d_g = d_o + g * MAX_EB;
                    mma_ts(d_g, a_lo, bh, idesc_o, !(from_zero && (k & 1) == 0));
#endif
                    mma_ts(d_g, a_hi, bl, idesc_o, 1);
                    mma_ts(d_g, a_hi, bh, idesc_o, 1);
                }
                commit(&empty_bar[slot]);
                commit(&pv_done[a]);
            }
            __syncwarp();
        };
        bool prev_first = false, pv_pending = false;
        WaveWork ww;
        for (int w = 0; w < P.W; ++w) {
            if (!wave_work(P, w, cta, ww)) continue;
            mbar_wait(u_full, seg & 1);
            ++seg;
#if KMB_PV16_SETS
            int until_flush2[2] = {0, 0};   // per set: its own blocks until its streams are flushed again
            const int flush_own = max(1, P.flush_blocks / 2);
#else
            int until_flush = 0;   // blocks until the epilogue flushes the O_g again (a countdown: no division in the loop)
#endif
            for (int sb = ww.sb_lo; sb < ww.sb_hi; ++sb, ++n) {
#if KMB_PV16_SETS
                int& until_flush = until_flush2[n & 1];
                const bool first = (until_flush == 0);
                until_flush = first ? flush_own - 1 : until_flush - 1;
#else
                const bool first = (until_flush == 0);
                until_flush = first ? P.flush_blocks - 1 : until_flush - 1;
#endif
                const int a = n & 1;   // stage a held P(n-2): PV(n-2) was issued in the previous iteration
                const uint32_t d_s = tmem_base + COL_S + a * TNS;
                for (int kb = 0; kb < P.kblocks; ++kb, ++it) {
                    const int slot = it % ST;
                    KMB_M(3);
                    mbar_wait(&full_bar[slot], (it / ST) & 1);
                    tc_fence_after();
                    KMB_M(4);
                    const unsigned char* bt = ring + slot * SLOT;
                    const unsigned char* at = u_region + kb * 2 * A_TILE_BYTES;
                    const int ksteps = (kb == P.kblocks - 1) ? P.ksteps_last : 4;
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            if (k < ksteps) {
                                const uint64_t ah = umma_desc_sw128(at, k * 32);
                                const uint64_t al = umma_desc_sw128(at + A_TILE_BYTES, k * 32);
                                const uint64_t bh = umma_desc_sw128(bt, k * 32);
                                const uint64_t bl = umma_desc_sw128(bt + SLOT / 2, k * 32);
                                mma_ss(d_s, al, bh, idesc_s, (kb | k) != 0);
                                mma_ss(d_s, ah, bl, idesc_s, 1);
                                mma_ss(d_s, ah, bh, idesc_s, 1);
                            }
                        }
                        commit(&empty_bar[slot]);
                    }
                    __syncwarp();
                }
                const bool last_of_tile = (sb + 1 == ww.sb_hi);
                if (elect_one()) {
                    commit(&acc_full[a]);
                    if (last_of_tile) commit(u_free);   // last S of this row tile
                }
                __syncwarp();
                if (pv_pending) issue_pv(n - 1, prev_first);
                pv_pending = true;
                prev_first = first;
            }
        }
        if (pv_pending) issue_pv(n - 1, prev_first);
#ifdef KMB_PV16_TIMING
        if (blockIdx.x == 0 && lane == 0) {
            for (int i = 0; i < 5; ++i) P.out[8 + i] = static_cast<float>(macc[i]) / n;
        }
#endif
    } else if (warp >= 2) {
        // -------------------------------------- epilogue --------------------------------------
        const int et = tid - 64;
        const int lane_group = warp & 3;             // TMEM lane quarter this warp may touch
        const int cg = (warp - 2) >> 2;              // column group = online-softmax stream (own reference, sum, O_g)
#if KMB_PV16_SETS
        // Two sets of eight warps work on ALTERNATE source blocks: stream cg belongs to set cg & 1 (blocks whose global
        // index n has n & 1 == set, i.e. S / P stage `set`) and covers columns [64 half, 64 half + 64) of them, half = cg >> 1,
        // in two passes of 32.  The four warps of an SM sub-partition (same TMEM lane quarter) are then two from each set,
        // half a block period apart: while one pair waits for its S tile, loads it, stores P or arrives, the other pair keeps
        // the MUFU pipe busy -- with all 16 warps on the same block those ~900 cycles per block leave it idle.
        const int set = cg & 1, half = cg >> 1;
#else
        const int col0 = cg * CPT;                   // first S column of this thread
#endif
        const int row_in_tile = lane_group * 32 + lane;
        const uint32_t lane_addr = static_cast<uint32_t>(lane_group * 32) << 16;
        const uint32_t o_mine = tmem_base + COL_O + cg * MAX_EB + lane_addr;   // this group's accumulator
        const float sscale = __ldg(P.sscale + 1);
        uint32_t n = 0;   // blocks this CTA has processed (all waves)
#ifdef KMB_PV16_TIMING
        long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tprev = clock64();
#define KMB_T(i) do { const long long t_ = clock64(); tacc[i] += t_ - tprev; tprev = t_; } while (0)
#else
#define KMB_T(i) do { } while (0)
#endif
        auto wait_pv = [&](uint32_t m) {   // PV(m) has read P(m) and finished accumulating into the O_g
            mbar_wait(&pv_done[m & 1], (m >> 1) & 1);
            tc_fence_after();
        };

        // |v|^2 of this group's 32 sources: lane l fetches source l of the NEXT block (one coalesced load, a block
        // ahead of its use), parks it in the warp's shared-memory line and every lane reads the line back as float4s
#if KMB_PV16_SETS
        float* my_line = vline + (warp - 2) * 2 * 64;   // 64 sources per own block, double buffered
        float vn_a = 0.f, vn_b = 0.f;                   // prefetched |v|^2 of the next own block (sources lane, 32 + lane)
        uint32_t pref_n = 0xffffffffu, own = 0, n_base = 0;   // block they belong to; own blocks so far; first block of the tile
#else
        float* my_line = vline + (warp - 2) * 2 * 32;
        float vn_next = 0.f;
        bool primed = false;
#endif
        WaveWork ww;
        for (int w = 0; w < P.W; ++w) {
            if (!wave_work(P, w, cta, ww)) continue;
            const int tile = PAIR ? ww.tile * 2 + static_cast<int>(rank) : ww.tile;
            const long long row = static_cast<long long>(tile) * TM + row_in_tile;
            const bool row_ok = row < P.N;
            const float un = row_ok ? __ldg(P.un + row) : 0.f;
            float ksum = 0.f, ref = -INFINITY;   // this group's stream
            // long accumulator of this thread's O_g row: zero, then only ever touched by this thread until the merge
            float* olong = P.olong + (static_cast<size_t>(blockIdx.x) * NG + cg) * (MAX_EB * TM) + row_in_tile;
            for (int c = 0; c < P.ebp; ++c) __stcg(olong + c * TM, 0.f);
            auto flush_o = [&]() {   // olong += O_g (after the last P.B into it has completed)
                for (int c0 = 0; c0 < P.ebp; c0 += 16) {
                    float o[16];
                    tmem_ld_cols<16>(o_mine + c0, o);
#pragma unroll
                    for (int c = 0; c < 16; ++c) atomicAdd(olong + (c0 + c) * TM, o[c]);   // RED: nothing to wait for
                }
            };
#if KMB_PV16_SETS
            const int cnt = ww.sb_hi - ww.sb_lo;
            int next_lo = -1, next_cnt = 0;   // the next wave this CTA works in (prefetch across the tile boundary)
            {
                WaveWork wn;
                for (int w2 = w + 1; w2 < P.W && next_lo < 0; ++w2)
                    if (wave_work(P, w2, cta, wn)) { next_lo = wn.sb_lo; next_cnt = wn.sb_hi - wn.sb_lo; }
            }
            const int flush_own = max(1, P.flush_blocks / 2);
            int until_flush = flush_own;   // own blocks until this thread moves its O_g row to the long accumulator
            bool flushed = false, any = false;   // a flush has happened in this tile; the stream has had a block in this tile
            uint32_t m_last = 0;                 // its last block
            for (int i = ((n_base & 1u) == static_cast<uint32_t>(set)) ? 0 : 1; i < cnt; i += 2) {
                const int sb = ww.sb_lo + i;
                const uint32_t n = n_base + i;   // n & 1 == set
                const long long jb = static_cast<long long>(sb) * TNS + 64 * half;
                float* line = my_line + (own & 1) * 64;
                ++own;
                if (pref_n != n) {   // nothing prefetched for this block (first block of the kernel, or a tile of one block)
                    vn_a = __ldg(P.vn + jb + lane);   // padded to whole blocks with 3.4e38
                    vn_b = __ldg(P.vn + jb + 32 + lane);
                }
                line[lane] = vn_a;
                line[32 + lane] = vn_b;
                {
                    int sbn = -1;
                    uint32_t nn = 0;
                    if (i + 2 < cnt) { sbn = sb + 2; nn = n + 2; }
                    else if (next_lo >= 0) {
                        const uint32_t nb = n_base + cnt;
                        const int i2 = ((nb & 1u) == static_cast<uint32_t>(set)) ? 0 : 1;
                        if (i2 < next_cnt) { sbn = next_lo + i2; nn = nb + i2; }
                    }
                    pref_n = 0xffffffffu;
                    if (sbn >= 0) {
                        const long long jn = static_cast<long long>(sbn) * TNS + 64 * half;
                        vn_a = __ldg(P.vn + jn + lane);
                        vn_b = __ldg(P.vn + jn + 32 + lane);
                        pref_n = nn;
                    }
                }
                __syncwarp();
                KMB_T(0);
                mbar_wait(&acc_full[set], (n >> 1) & 1);
                tc_fence_after();
                KMB_T(1);
#pragma unroll 1
                for (int pass = 0; pass < 2; ++pass) {
                    const int col0 = 64 * half + 32 * pass;
                    const long long j0 = static_cast<long long>(sb) * TNS + col0;
                    const uint32_t st_addr = tmem_base + COL_S + set * TNS + col0 + lane_addr;
                    const float4* vnq = reinterpret_cast<const float4*>(line + 32 * pass);
                    float2 t2[CPT / 2];   // S as pairs of sources, then t = -log2 k
                    tmem_ld_cols<CPT>(st_addr, reinterpret_cast<float(&)[CPT]>(t2));
                    KMB_T(2);
                    // t = -log2 of the kernel values (packed pairs) and their minimum over this thread's columns
                    float tmin = INFINITY;
                    if (j0 < P.M) {
                        const float2 nss2 = make_float2(-sscale, -sscale), un2 = make_float2(un, un);
#pragma unroll
                        for (int c = 0; c < CPT / 4; ++c) {
                            const float4 vq = vnq[c];   // broadcast read of the warp's line
                            const float2 wa = make_float2(vq.x, vq.y), wb = make_float2(vq.z, vq.w);
                            const float2 ta = neg_log2_kernel2<KID>(t2[2 * c], nss2, KID == KMB_KERNEL_GAUSSIAN ? wa : add2(wa, un2));
                            const float2 tb = neg_log2_kernel2<KID>(t2[2 * c + 1], nss2, KID == KMB_KERNEL_GAUSSIAN ? wb : add2(wb, un2));
                            t2[2 * c] = ta;
                            t2[2 * c + 1] = tb;
                            tmin = fminf(fminf(tmin, ta.x), ta.y);
                            tmin = fminf(fminf(tmin, tb.x), tb.y);
                        }
                    } else {   // every source of this chunk is padding (warp-uniform): all weights are zero
#pragma unroll
                        for (int c = 0; c < CPT / 2; ++c) t2[c] = make_float2(INFINITY, INFINITY);
                    }
                    constexpr bool kRowTermOut = (KID == KMB_KERNEL_GAUSSIAN);   // t2 lacks the row's |u|^2
                    const float cm = kRowTermOut ? -(tmin + un) : -tmin;   // largest log2 k of the chunk
                    KMB_T(3);
                    // lazy rescale with INTEGER reference exponents: every rescale factor is an exact power of two, so the
                    // P of pass 0 that is already in tensor memory can follow a reference that moves in pass 1 without error
                    {
                        bool need = false;
                        if (ref == -INFINITY) ref = (cm == -INFINITY) ? cm : ceilf(cm);   // nothing but zero weights so far
                        else need = cm > ref + kLazyRescale;
                        if (__any_sync(0xffffffffu, need)) {
                            const float ref_new = need ? ceilf(cm) : ref;
                            const float sc = need ? pow2_int(ref - ref_new) : 1.f;
                            if (any) {   // O_g holds this tile's sums
                                wait_pv(m_last);
                                for (int c0 = 0; c0 < P.ebp; c0 += 16) {
                                    float o[16];
                                    tmem_ld_cols<16>(o_mine + c0, o);
#pragma unroll
                                    for (int c = 0; c < 16; ++c) o[c] *= sc;
                                    tmem_st_cols<16>(o_mine + c0, o);
                                }
                                if (flushed && sc != 1.f)   // something was flushed already
                                    for (int c = 0; c < P.ebp; ++c) __stcg(olong + c * TM, sc * __ldcg(olong + c * TM));
                            }
                            if (pass == 1) {   // the P of pass 0 (16 hi + 16 lo columns of packed halves) waits for the same P.B
                                float pq[CPT];
                                tmem_ld_cols<CPT>(st_addr - CPT, pq);
                                const __half2 sc2 = __float2half2_rn(sc);
#pragma unroll
                                for (int c = 0; c < CPT; ++c) {
                                    __half2 h = *reinterpret_cast<__half2*>(&pq[c]);
                                    h = __hmul2(h, sc2);
                                    pq[c] = *reinterpret_cast<float*>(&h);
                                }
                                tmem_st_cols<CPT>(st_addr - CPT, pq);
                            }
                            tmem_st_wait();
                            ksum *= sc;
                            ref = ref_new;
                        }
                    }
                    KMB_T(4);
                    // P = 2^(log2 k - ref) = 2^(-ref - t), FP16 hi / lo, two sources per TMEM column, over this thread's own S columns
                    {
                        uint32_t ph[CPT / 2], pl[CPT / 2];
                        float2 kacc = make_float2(0.f, 0.f);   // two-level sum of the weights (see kprod_direct.cuh)
                        const float nref = (ref == -INFINITY) ? 0.f : (kRowTermOut ? -ref - un : -ref);
                        const float2 nref2 = make_float2(nref, nref);
#pragma unroll
                        for (int c = 0; c < CPT / 2; ++c) {
                            const float2 e = sub2(nref2, t2[c]);
                            const float2 pw = make_float2(ex2_approx(e.x), ex2_approx(e.y));
                            kacc = add2(kacc, pw);
                            const float2 h = make_float2(__uint_as_float(__float_as_uint(pw.x) & 0xffffe000u),
                                                         __uint_as_float(__float_as_uint(pw.y) & 0xffffe000u));
                            const float2 l = sub2(pw, h);
                            ph[c] = pack_half2(h.x, h.y);
                            pl[c] = pack_half2(l.x, l.y);
                        }
                        tmem_st_32x16(st_addr, reinterpret_cast<const float*>(ph));
                        tmem_st_32x16(st_addr + CPT / 2, reinterpret_cast<const float*>(pl));
                        ksum += kacc.x + kacc.y;
                    }
                    KMB_T(5);
                }
                tmem_st_wait();
                if (until_flush == 0) {   // P.B(n) starts this stream's O_g from zero
                    wait_pv(m_last);
                    flush_o();
                    until_flush = flush_own;
                    flushed = true;
                }
                --until_flush;
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (PAIR && rank != 0) pair::mbar_arrive_cluster(pair::map_to_cta(&p_ready[set], 0));
                    else mbar_arrive(&p_ready[set]);
                }
                any = true;
                m_last = n;
                KMB_T(6);
            }
#ifdef KMB_PV16_TIMING
            if (blockIdx.x == 0 && et == 0) {
                for (int i = 0; i < 7; ++i) P.out[i] = static_cast<float>(tacc[i]) / max(1u, own);
                P.out[7] = static_cast<float>(own);
            }
#endif
#else
            if (!primed) {
                vn_next = __ldg(P.vn + static_cast<long long>(ww.sb_lo) * TNS + col0 + lane);   // padded to whole blocks with 3.4e38
                primed = true;
            }
            // first block of the next wave this CTA works in (for the prefetch across the tile boundary)
            int sb_next_tile = -1;
            {
                WaveWork wn;
                for (int w2 = w + 1; w2 < P.W && sb_next_tile < 0; ++w2)
                    if (wave_work(P, w2, cta, wn)) sb_next_tile = wn.sb_lo;
            }

#if KMB_PV16_SKEW_NS > 0
            if (cg > 0) __nanosleep(cg * KMB_PV16_SKEW_NS);
#endif
            int until_flush = P.flush_blocks;   // blocks until this thread moves its O_g row to the long accumulator
            bool flushed = false;               // ... which has happened at least once in this tile
            for (int sb = ww.sb_lo; sb < ww.sb_hi; ++sb, ++n) {
                const long long j0 = static_cast<long long>(sb) * TNS + col0;
                const int a = n & 1;
                const uint32_t st_addr = tmem_base + COL_S + a * TNS + col0 + lane_addr;
#if KMB_PV16_VN_LDG
                const float4* vnq = reinterpret_cast<const float4*>(P.vn + j0);   // 128-byte aligned; padded to whole blocks
#else
                float* line = my_line + (n & 1) * 32;
                line[lane] = vn_next;
                {
                    const int sbn = (sb + 1 < ww.sb_hi) ? sb + 1 : sb_next_tile;
                    if (sbn >= 0) vn_next = __ldg(P.vn + static_cast<long long>(sbn) * TNS + col0 + lane);
                }
                __syncwarp();
                const float4* vnq = reinterpret_cast<const float4*>(line);
#endif
                KMB_T(0);
                mbar_wait(&acc_full[a], (n >> 1) & 1);
                tc_fence_after();
                KMB_T(1);
#if KMB_PV16_FUSED
                // One fused pass per block: S -> t = -log2 k -> P = 2^(-ref - t) with the reference exponent the row ALREADY has
                // (it is lazy: it only moves when the block maximum outgrows it by 2^8), FP16 hi / lo, packed.  rsqrt / ex2
                // (MUFU) and the packed FP32 work of different columns interleave instead of forming two phases in which the
                // four warps of an SM sub-partition first all wait for the FMA pipe and then all for the MUFU pipe.  If the
                // reference has to move (first block of a tile; afterwards almost never) the O_g are rescaled and the pass is
                // repeated from the S columns, which are still in tensor memory.
                uint32_t ph[CPT / 2], pl[CPT / 2];
                float2 kacc;
                bool again;
                do {
                    float2 t2[CPT / 2];   // S as pairs of sources
                    tmem_ld_cols<CPT>(st_addr, reinterpret_cast<float(&)[CPT]>(t2));
                    KMB_T(2);
                    constexpr bool kRowTermOut = (KID == KMB_KERNEL_GAUSSIAN);   // t lacks the row's |u|^2
                    // all -inf so far: every weight is 2^-inf = 0
                    const float nref = (ref == -INFINITY) ? 0.f : (kRowTermOut ? -ref - un : -ref);
                    const float2 nref2 = make_float2(nref, nref);
                    float2 tmin2[2] = {make_float2(INFINITY, INFINITY), make_float2(INFINITY, INFINITY)};
                    kacc = make_float2(0.f, 0.f);   // two-level sum of the weights (see kprod_direct.cuh)
                    if (j0 < P.M) {
                        const float2 nss2 = make_float2(-sscale, -sscale), un2 = make_float2(un, un);
#pragma unroll
                        for (int c = 0; c < CPT / 2; ++c) {
                            const float4 vq = vnq[c >> 1];   // broadcast read of the warp's line
                            // Gaussian: |u|^2 is the same for the whole row, so it is left out of t here and added to the
                            // block minimum / subtracted with the reference exponent (one FADD2 per four values less)
                            const float2 w = (c & 1) ? make_float2(vq.z, vq.w) : make_float2(vq.x, vq.y);
                            const float2 t = neg_log2_kernel2<KID>(t2[c], nss2, KID == KMB_KERNEL_GAUSSIAN ? w : add2(w, un2));
                            tmin2[c & 1] = make_float2(fminf(tmin2[c & 1].x, t.x), fminf(tmin2[c & 1].y, t.y));
                            const float2 e = sub2(nref2, t);
                            const float2 pw = make_float2(ex2_approx(e.x), ex2_approx(e.y));
                            kacc = add2(kacc, pw);
                            // 11 significant bits: exact in FP16
                            const float2 h = make_float2(__uint_as_float(__float_as_uint(pw.x) & 0xffffe000u),
                                                         __uint_as_float(__float_as_uint(pw.y) & 0xffffe000u));
                            const float2 l = sub2(pw, h);
                            ph[c] = pack_half2(h.x, h.y);
                            pl[c] = pack_half2(l.x, l.y);
                        }
                    } else {   // every source of this group is padding (warp-uniform): all weights are zero
#pragma unroll
                        for (int c = 0; c < CPT / 2; ++c) ph[c] = pl[c] = 0u;
                    }
                    const float tmin = fminf(fminf(tmin2[0].x, tmin2[0].y), fminf(tmin2[1].x, tmin2[1].y));
                    const float cm = kRowTermOut ? -(tmin + un) : -tmin;   // largest log2 k of the block
                    KMB_T(3);
                    // lazy rescale: keep the reference exponent unless the maximum outgrew it by 2^8 (P <= 2^8 fits FP16;
                    // a first block, ref == -inf, always adopts its maximum)
                    const bool need = (ref == -INFINITY) ? (cm != -INFINITY) : (cm > ref + kLazyRescale);
                    again = __any_sync(0xffffffffu, need);
                    if (again) {
                        const float sc = need ? ex2_approx(ref - cm) : 1.f;   // ref == -inf: 0 (nothing accumulated yet)
                        if (sb > ww.sb_lo) {   // O_g holds this tile's sums
                            wait_pv(n - 1);
                            for (int c0 = 0; c0 < P.ebp; c0 += 16) {
                                float o[16];
                                tmem_ld_cols<16>(o_mine + c0, o);
#pragma unroll
                                for (int c = 0; c < 16; ++c) o[c] *= sc;
                                tmem_st_cols<16>(o_mine + c0, o);
                            }
                            tmem_st_wait();
                            if (flushed && sc != 1.f)   // something was flushed already
                                for (int c = 0; c < P.ebp; ++c) __stcg(olong + c * TM, sc * __ldcg(olong + c * TM));
                        }
                        ksum *= sc;
                        if (need) ref = cm;
                    }
                    KMB_T(4);
                } while (again);
                // P over this thread's own S columns: FP16 hi / lo, two sources per TMEM column
                tmem_st_32x16(st_addr, reinterpret_cast<const float*>(ph));
                tmem_st_32x16(st_addr + CPT / 2, reinterpret_cast<const float*>(pl));
                ksum += kacc.x + kacc.y;
#else   // two phases: all t = -log2 k of the block, then (reference settled) all P
                float2 t2[CPT / 2];   // S as pairs of sources, then t = -log2 k
                tmem_ld_cols<CPT>(st_addr, reinterpret_cast<float(&)[CPT]>(t2));
                KMB_T(2);

                // t = -log2 of the kernel values (packed pairs) and their minimum over this thread's columns
                float tmin = INFINITY;
                if (j0 < P.M) {
                    const float2 nss2 = make_float2(-sscale, -sscale), un2 = make_float2(un, un);
#pragma unroll
                    for (int c = 0; c < CPT / 4; ++c) {
                        const float4 vq = KMB_PV16_VN_LDG ? __ldg(vnq + c) : vnq[c];   // broadcast read (L1 / the warp's line)
                        // Gaussian: |u|^2 is the same for the whole row, so it is left out of t here and added to the
                        // block minimum / subtracted with the reference exponent below (one FADD2 per four values less)
                        const float2 wa = make_float2(vq.x, vq.y), wb = make_float2(vq.z, vq.w);
                        const float2 ta = neg_log2_kernel2<KID>(t2[2 * c], nss2, KID == KMB_KERNEL_GAUSSIAN ? wa : add2(wa, un2));
                        const float2 tb = neg_log2_kernel2<KID>(t2[2 * c + 1], nss2, KID == KMB_KERNEL_GAUSSIAN ? wb : add2(wb, un2));
                        t2[2 * c] = ta;
                        t2[2 * c + 1] = tb;
                        tmin = fminf(fminf(tmin, ta.x), ta.y);
                        tmin = fminf(fminf(tmin, tb.x), tb.y);
                    }
                } else {   // every source of this group is padding (warp-uniform): all weights are zero
#pragma unroll
                    for (int c = 0; c < CPT / 2; ++c) t2[c] = make_float2(INFINITY, INFINITY);
                }
                constexpr bool kRowTermOut = (KID == KMB_KERNEL_GAUSSIAN);   // t2 lacks the row's |u|^2
                const float cm = kRowTermOut ? -(tmin + un) : -tmin;   // largest log2 k of the block
                KMB_T(3);
                // lazy rescale: keep the reference exponent unless the maximum outgrew it by 2^8
                {
                    bool need = false;
                    if (ref == -INFINITY) ref = cm;   // nothing but zero weights so far
                    else need = cm > ref + kLazyRescale;
                    if (__any_sync(0xffffffffu, need)) {
                        const float sc = need ? ex2_approx(ref - cm) : 1.f;
                        if (sb > ww.sb_lo) {   // O_g holds this tile's sums
                            wait_pv(n - 1);
                            for (int c0 = 0; c0 < P.ebp; c0 += 16) {
                                float o[16];
                                tmem_ld_cols<16>(o_mine + c0, o);
#pragma unroll
                                for (int c = 0; c < 16; ++c) o[c] *= sc;
                                tmem_st_cols<16>(o_mine + c0, o);
                            }
                            tmem_st_wait();
                            if (flushed && sc != 1.f)   // something was flushed already
                                for (int c = 0; c < P.ebp; ++c) __stcg(olong + c * TM, sc * __ldcg(olong + c * TM));
                        }
                        ksum *= sc;
                        if (need) ref = cm;
                    }
                }
                KMB_T(4);
                // P = 2^(log2 k - ref) = 2^(-ref - t), FP16 hi / lo, two sources per TMEM column, over this thread's own S columns
                {
                    uint32_t ph[CPT / 2], pl[CPT / 2];
                    float2 kacc = make_float2(0.f, 0.f);   // two-level sum of the weights (see kprod_direct.cuh)
                    // all -inf so far: every weight is 2^-inf = 0
                    const float nref = (ref == -INFINITY) ? 0.f : (kRowTermOut ? -ref - un : -ref);
                    const float2 nref2 = make_float2(nref, nref);
#pragma unroll
                    for (int c = 0; c < CPT / 2; ++c) {
                        const float2 e = sub2(nref2, t2[c]);
                        const float2 pw = make_float2(ex2_approx(e.x), ex2_approx(e.y));
                        kacc = add2(kacc, pw);
                        // 11 significant bits: exact in FP16
                        const float2 h = make_float2(__uint_as_float(__float_as_uint(pw.x) & 0xffffe000u),
                                                     __uint_as_float(__float_as_uint(pw.y) & 0xffffe000u));
                        const float2 l = sub2(pw, h);
                        ph[c] = pack_half2(h.x, h.y);
                        pl[c] = pack_half2(l.x, l.y);
                    }
                    tmem_st_32x16(st_addr, reinterpret_cast<const float*>(ph));
                    tmem_st_32x16(st_addr + CPT / 2, reinterpret_cast<const float*>(pl));
                    ksum += kacc.x + kacc.y;
                }
#endif
                KMB_T(5);
                tmem_st_wait();
                if (until_flush == 0) {   // P.B(n) starts the O_g from zero
                    wait_pv(n - 1);
                    flush_o();
                    until_flush = P.flush_blocks;
                    flushed = true;
                }
                --until_flush;
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (PAIR && rank != 0) pair::mbar_arrive_cluster(pair::map_to_cta(&p_ready[a], 0));
                    else mbar_arrive(&p_ready[a]);
                }
                KMB_T(6);
            }
#ifdef KMB_PV16_TIMING
            if (blockIdx.x == 0 && et == 0) {
                for (int i = 0; i < 7; ++i) P.out[i] = static_cast<float>(tacc[i]) / n;
                P.out[7] = static_cast<float>(n);
            }
#endif

#endif

            // ------------------------------ row tile done: merge the four streams ------------------------------
            refbuf[cg * TM + row_in_tile] = ref;
            ksbuf[cg * TM + row_in_tile] = ksum;
#if KMB_PV16_SETS
            if (any) {   // (a stream without a block in this tile has nothing in its O_g: its long accumulator stays zero)
                wait_pv(m_last);
                flush_o();
            }
            n_base += cnt;
#else
            wait_pv(n - 1);   // the tile's last PV
            flush_o();        // the long accumulators now hold the whole tile
#endif
            __threadfence();
            named_bar_sync(2, EPI_THREADS);
            float rmax = -INFINITY, wg[NG], ktot = 0.f;
#pragma unroll
            for (int g = 0; g < NG; ++g) rmax = fmaxf(rmax, refbuf[g * TM + row_in_tile]);
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                const float rg = refbuf[g * TM + row_in_tile];
                wg[g] = (rg == -INFINITY) ? 0.f : ex2_approx(rg - rmax);
                ktot = fmaf(wg[g], ksbuf[g * TM + row_in_tile], ktot);   // fixed order
            }
            const bool complete = (ww.Cw == 1);
            const bool ghost = tile >= P.n_tiles;   // pairs: an odd number of row tiles leaves the last peer without one
            // partial records of one wave: [row tile or pair in wave][range c]([rank])
            const size_t slot0 = static_cast<size_t>(w) * P.slots_per_wave + static_cast<size_t>(ww.tile_in_wave) * ww.Cw * NCTA + rank;
            float* mine = P.partial + (slot0 + static_cast<size_t>(ww.c) * NCTA) * (TM * PS);
            // plain product: undo the reference exponent (2^ref may underflow exactly where FP32 K b would)
            const float row_scale = NORM ? 1.f / ktot : ((rmax == -INFINITY) ? 0.f : ex2_approx(rmax));
            for (int c0 = cg * 16; c0 < P.ebp; c0 += NG * 16) {   // this thread merges 16-column chunks c0 of all four O_g
                float o[16];
#pragma unroll
                for (int c = 0; c < 16; ++c) o[c] = 0.f;
#pragma unroll
                for (int g = 0; g < NG; ++g) {
                    const float* og = P.olong + (static_cast<size_t>(blockIdx.x) * NG + g) * (MAX_EB * TM) + row_in_tile;
#pragma unroll
                    for (int c = 0; c < 16; ++c) o[c] = fmaf(wg[g], __ldcg(og + (c0 + c) * TM), o[c]);
                }
                if (ghost) continue;
                if (complete) {
                    if (row_ok) {
#pragma unroll
                        for (int c = 0; c < 16; ++c)
                            if (c0 + c < P.eb) P.out[row * P.E + P.e0 + c0 + c] = o[c] * __ldg(P.binv + P.e0 + c0 + c) * row_scale;
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < 16; ++c) mine[(c0 + c) * TM + row_in_tile] = o[c];
                }
            }
            tc_fence_before();
            if (!complete && !ghost) {
                if (cg == 0) {
                    mine[MAX_EB * TM + row_in_tile] = ktot;
                    mine[(MAX_EB + 1) * TM + row_in_tile] = rmax;
                }
                __threadfence();
                named_bar_sync(2, EPI_THREADS);
                if (et == 0) {
                    const int old = atomicAdd(&P.tile_counter[tile], 1);
                    const int last = (old == ww.Cw - 1);
                    if (last) P.tile_counter[tile] = 0;
                    *s_flag = last;
                }
                named_bar_sync(2, EPI_THREADS);
                const bool is_last = *s_flag != 0;
                named_bar_sync(2, EPI_THREADS);
                if (is_last && row_ok) {
                    __threadfence();
                    float mx = -INFINITY;
                    for (int c = 0; c < ww.Cw; ++c)
                        mx = fmaxf(mx, __ldcg(P.partial + (slot0 + static_cast<size_t>(c) * NCTA) * (TM * PS) + (MAX_EB + 1) * TM + row_in_tile));
                    for (int e = cg; e < P.eb; e += NG) {   // the groups share the signal columns of the row
                        float sum = 0.f, l = 0.f;
                        for (int c = 0; c < ww.Cw; ++c) {
                            const float* ps = P.partial + (slot0 + static_cast<size_t>(c) * NCTA) * (TM * PS);
                            const float m = __ldcg(ps + (MAX_EB + 1) * TM + row_in_tile);
                            const float wgt = (m == -INFINITY) ? 0.f : ex2_approx(m - mx);
                            sum = fmaf(wgt, __ldcg(ps + e * TM + row_in_tile), sum);
                            l = fmaf(wgt, __ldcg(ps + MAX_EB * TM + row_in_tile), l);
                        }
                        const float rs = NORM ? 1.f / l : ((mx == -INFINITY) ? 0.f : ex2_approx(mx));
                        P.out[row * P.E + P.e0 + e] = sum * __ldg(P.binv + P.e0 + e) * rs;
                    }
                }
            } else {
                named_bar_sync(2, EPI_THREADS);   // refbuf / ksbuf are rewritten at the end of the next tile
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if constexpr (PAIR) pair::cluster_sync_all();   // neither CTA frees tensor memory (or exits) while the other may signal it
    if (warp == 1) {
        if constexpr (PAIR) pair::tmem_dealloc2(tmem_base, TMEM_COLS);
        else tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

template <int KID, bool NORM>
__global__ void __launch_bounds__(THREADS, 1)
kprod_tensor_pv16_kernel(const __grid_constant__ CUtensorMap map_ah, const __grid_constant__ CUtensorMap map_al,
                         const __grid_constant__ CUtensorMap map_bh, const __grid_constant__ CUtensorMap map_bl,
                         const __grid_constant__ CUtensorMap map_sh, const __grid_constant__ CUtensorMap map_sl,
                         const Params P) {
    pv16_body<KID, NORM, false>(map_ah, map_al, map_bh, map_bl, map_sh, map_sl, P);
}
template <int KID, bool NORM>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
kprod_tensor_pv16_pair_kernel(const __grid_constant__ CUtensorMap map_ah, const __grid_constant__ CUtensorMap map_al,
                              const __grid_constant__ CUtensorMap map_bh, const __grid_constant__ CUtensorMap map_bl,
                              const __grid_constant__ CUtensorMap map_sh, const __grid_constant__ CUtensorMap map_sl,
                              const Params P) {
    pv16_body<KID, NORM, true>(map_ah, map_al, map_bh, map_bl, map_sh, map_sl, P);
}

// ---- signal planes -----------------------------------------------------------------------------------
// per block: column maxima of |b|
static __global__ void __launch_bounds__(256) signal_absmax_kernel(const float* __restrict__ b, long long M, int E,
                                                                   float* __restrict__ pmax) {
    __shared__ float sm[8][32];
    const int col = blockIdx.y * 32 + (threadIdx.x & 31);
    const int rl = threadIdx.x >> 5;
    float hi = 0.f;
    if (col < E)
        for (long long r = blockIdx.x * 8 + rl; r < M; r += static_cast<long long>(gridDim.x) * 8) hi = fmaxf(hi, fabsf(b[r * E + col]));
    sm[rl][threadIdx.x & 31] = hi;
    __syncthreads();
    if (rl == 0 && col < E) {
        float h = 0.f;
        for (int i = 0; i < 8; ++i) h = fmaxf(h, sm[i][threadIdx.x]);
        pmax[static_cast<size_t>(blockIdx.x) * E + col] = h;
    }
}
// bscale[e] = 2^q with 2^q max_j |b[j][e]| in [2^13, 2^14); binv[e] = 2^-q
static __global__ void signal_scale_kernel(const float* __restrict__ pmax, int blocks, int E, int Ep, float* __restrict__ bscale,
                                           float* __restrict__ binv) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= Ep) return;
    float m = 0.f;
    if (e < E)
        for (int b = 0; b < blocks; ++b) m = fmaxf(m, pmax[static_cast<size_t>(b) * E + e]);
    int q = 0;
    if (m > 0.f && m < INFINITY) q = 13 - ilogbf(m);
    q = max(-100, min(100, q));
    bscale[e] = exp2f(static_cast<float>(q));
    binv[e] = exp2f(static_cast<float>(-q));
}
// hi/lo[sb][e][jj] = FP16 hi/lo of 2^q_e b[128 sb + jj][e]: the K-major signal of the P.B contraction, one contiguous
// (Ep x 128) slab per source block (a TMA box then reads 64 rows of 128 bytes 256 bytes apart, not 64 rows that are
// 2 Mp bytes apart), zero padded
static __global__ void transpose_split_signal_f16_kernel(const float* __restrict__ b, long long M, long long Mp, int E, int Ep,
                                                         const float* __restrict__ bscale, __half* __restrict__ hi,
                                                         __half* __restrict__ lo) {
    const long long j = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    const int e = blockIdx.y;
    if (j >= Mp) return;
    const float v = (j < M && e < E) ? b[j * E + e] * bscale[e] : 0.f;
    const __half h = __float2half_rn(v);
    const long long at = ((j / TNS) * Ep + e) * TNS + (j % TNS);
    hi[at] = h;
    lo[at] = __float2half_rn(v - __half2float(h));
}

}  // namespace pv16

namespace {

size_t align_up_pv16(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct Pv16Plan {
    bool pair;   // CTA pairs (cta_group::2): at least two row tiles
    int Dp, Ep, kblocks, ksteps_last, stages, grid, smem;
    long long n_tiles, nsb, Mp;
    tc::WavePlan waves;
    size_t off_center, off_stats, off_sscale, off_uh, off_ul, off_vh, off_vl, off_un, off_vn, off_sh, off_sl, off_bmax, off_bscale,
        off_binv, off_partial, off_olong, off_counter, total;
};

int plan_pv16(int64_t N, int64_t M, int D, int E, Pv16Plan* pl) {
    pl->Dp = (D + 15) / 16 * 16;
    pl->kblocks = (pl->Dp + 63) / 64;
    pl->ksteps_last = (pl->Dp - (pl->kblocks - 1) * 64) / 16;
    pl->Ep = (E + pv16::MAX_EB - 1) / pv16::MAX_EB * pv16::MAX_EB;
    pl->Mp = (M + pv16::TNS - 1) / pv16::TNS * pv16::TNS;
    pl->n_tiles = (N + tc::TM - 1) / tc::TM;
    pl->nsb = (M + pv16::TNS - 1) / pv16::TNS;
    int dev = 0, sms = 0, smem_max = 0;
    KMB_CUDA_CHECK(cudaGetDevice(&dev));
    KMB_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    KMB_CUDA_CHECK(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    static const bool pair_enabled = [] {   // tuning knob: KMB_TENSOR_PAIR=0 keeps the single-CTA kernel
        const char* e = getenv("KMB_TENSOR_PAIR");
        return !(e && e[0] == '0');
    }();
    pl->pair = pair_enabled && pl->n_tiles >= 2 && sms >= 2;
    pl->grid = pl->pair ? sms / 2 * 2 : sms;
    const int slot = pl->pair ? pv16::SLOT_BYTES / 2 : pv16::SLOT_BYTES;
    const int fixed = 1024 + pl->kblocks * 2 * pv16::A_TILE_BYTES + pv16::EPI_WARPS * 2 * 64 * 4 + 2 * pv16::NG * tc::TM * 4 + 512;
    pl->stages = std::min(pl->pair ? 10 : 6, (smem_max - fixed) / slot);
    if (pl->stages < 3) return set_error(KMB_ERR_UNSUPPORTED, "not enough shared memory for D=%d", D);
    pl->smem = fixed + pl->stages * slot;
    if (pl->pair) {   // the wave plan counts pairs of row tiles and clusters
        tc::plan_waves((pl->n_tiles + 1) / 2, pl->nsb, pl->grid / 2, static_cast<size_t>(tc::TM) * pl->Dp * 8, &pl->waves);
        pl->waves.slots_per_wave *= 2;
        pl->waves.partial_slots *= 2;
    } else {
        tc::plan_waves(pl->n_tiles, pl->nsb, pl->grid, static_cast<size_t>(tc::TM) * pl->Dp * 4, &pl->waves);
    }
    F16PointsLayout L;   // the points-only head shared with kprod_tensor (kmb_product_prepare_f32)
    f16_points_layout(N, M, D, &L);
    pl->off_center = L.off_center; pl->off_stats = L.off_stats; pl->off_sscale = L.off_sscale;
    pl->off_uh = L.off_uh; pl->off_ul = L.off_ul; pl->off_vh = L.off_vh; pl->off_vl = L.off_vl;
    pl->off_un = L.off_un; pl->off_vn = L.off_vn;
    size_t o = L.end;
    auto take = [&](size_t bytes) { size_t at = o; o += align_up_pv16(bytes, 256); return at; };
    pl->off_sh = take(2 * static_cast<size_t>(pl->Ep) * pl->Mp);
    pl->off_sl = take(2 * static_cast<size_t>(pl->Ep) * pl->Mp);
    pl->off_bmax = take(sizeof(float) * tc::CENTER_BLOCKS * E);
    pl->off_bscale = take(sizeof(float) * pl->Ep);
    pl->off_binv = take(sizeof(float) * pl->Ep);
    pl->off_partial = take(sizeof(float) * pl->waves.partial_slots * tc::TM * pv16::PS);
    pl->off_olong = take(sizeof(float) * pl->grid * pv16::NG * pv16::MAX_EB * tc::TM);
    pl->off_counter = take(sizeof(int) * pl->n_tiles);
    pl->total = o;
    return KMB_OK;
}

template <int KID, bool NORM>
int launch_pv16(const CUtensorMap* m, const pv16::Params& P, int grid, int smem, bool pair, cudaStream_t stream) {
    if (pair) {
        auto fn = pv16::kprod_tensor_pv16_pair_kernel<KID, NORM>;
        if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(fn), smem)) return rc;
        fn<<<grid, pv16::THREADS, smem, stream>>>(m[0], m[1], m[2], m[3], m[4], m[5], P);
    } else {
        auto fn = pv16::kprod_tensor_pv16_kernel<KID, NORM>;
        if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(fn), smem)) return rc;
        fn<<<grid, pv16::THREADS, smem, stream>>>(m[0], m[1], m[2], m[3], m[4], m[5], P);
    }
    KMB_CUDA_CHECK(cudaGetLastError());
    return KMB_OK;
}

}  // namespace

bool tensor_pv16_applicable(int D, int E, int kid) {
    return E > 4 && D <= 128 && (kid == KMB_KERNEL_GAUSSIAN || kid == KMB_KERNEL_ABSOLUTE_EXPONENTIAL);
}

int tensor_pv16_workspace_bytes(int64_t N, int64_t M, int D, int E, size_t* bytes) {
    Pv16Plan pl{};
    if (int rc = plan_pv16(N, M, D, E, &pl)) return rc;
    *bytes = pl.total;
    return KMB_OK;
}

int tensor_pv16_product(const float* x, const float* y, const float* b, float* out, int64_t N, int64_t M, int D, int E, int kid,
                        int flags, void* workspace, size_t workspace_bytes, cudaStream_t stream, cudaEvent_t ev0,
                        cudaEvent_t ev1, bool prepared) {
    Pv16Plan pl{};
    if (int rc = plan_pv16(N, M, D, E, &pl)) return rc;
    if (!workspace || workspace_bytes < pl.total)
        return set_error(KMB_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", pl.total, workspace_bytes);
    if (N >= (1ll << 31) - tc::TM || M >= (1ll << 31) - pv16::TNS)
        return set_error(KMB_ERR_UNSUPPORTED, "tensor path indexes rows with 32-bit TMA coordinates");
    if (!b) return set_error(KMB_ERR_INVALID, "signal is NULL");
    char* ws = static_cast<char*>(workspace);
    auto F = [&](size_t off) { return reinterpret_cast<float*>(ws + off); };
    void *uh = ws + pl.off_uh, *ul = ws + pl.off_ul, *vh = ws + pl.off_vh, *vl = ws + pl.off_vl, *sh = ws + pl.off_sh, *sl = ws + pl.off_sl;
    int* counters = reinterpret_cast<int*>(ws + pl.off_counter);
    const bool norm = flags & KMB_FLAG_NORMALIZE_ROWS;

    KMB_CUDA_CHECK(cudaMemsetAsync(counters, 0, sizeof(int) * pl.n_tiles, stream));
    if (!prepared) {
        F16PointsLayout L;
        f16_points_layout(N, M, D, &L);
        if (int rc = f16_points_prepass(x, y, N, M, D, kid, L, ws, stream)) return rc;
    }
    {
        const int blocks = static_cast<int>(std::min<long long>(tc::CENTER_BLOCKS, (M + 7) / 8));
        pv16::signal_absmax_kernel<<<dim3(blocks, (E + 31) / 32), 256, 0, stream>>>(b, M, E, F(pl.off_bmax));
        KMB_CUDA_CHECK(cudaGetLastError());
        pv16::signal_scale_kernel<<<(pl.Ep + 127) / 128, 128, 0, stream>>>(F(pl.off_bmax), blocks, E, pl.Ep, F(pl.off_bscale), F(pl.off_binv));
        KMB_CUDA_CHECK(cudaGetLastError());
        dim3 g(static_cast<unsigned>((pl.Mp + 255) / 256), pl.Ep);
        pv16::transpose_split_signal_f16_kernel<<<g, 256, 0, stream>>>(b, M, pl.Mp, E, pl.Ep, F(pl.off_bscale), static_cast<__half*>(sh),
                                                                        static_cast<__half*>(sl));
        KMB_CUDA_CHECK(cudaGetLastError());
        count_launch(3);
    }
    CUtensorMap maps[6];
    if (int rc = tc::make_tensor_map_f16(&maps[0], uh, N, pl.Dp, tc::TM)) return rc;
    if (int rc = tc::make_tensor_map_f16(&maps[1], ul, N, pl.Dp, tc::TM)) return rc;
    // CTA pairs: each CTA loads 64 of a block's 128 sources and half of the pass's signal columns
    if (int rc = tc::make_tensor_map_f16(&maps[2], vh, M, pl.Dp, pl.pair ? pv16::TNS / 2 : pv16::TNS)) return rc;
    if (int rc = tc::make_tensor_map_f16(&maps[3], vl, M, pl.Dp, pl.pair ? pv16::TNS / 2 : pv16::TNS)) return rc;

    const int n_passes = pl.Ep / pv16::MAX_EB;
    for (int pass = 0; pass < n_passes; ++pass) {
        pv16::Params P;
        P.un = F(pl.off_un);
        P.vn = F(pl.off_vn);
        P.sscale = F(pl.off_sscale);
        P.binv = F(pl.off_binv);
        P.out = out;
        P.partial = F(pl.off_partial);
        P.olong = F(pl.off_olong);
        P.tile_counter = counters;
        P.N = N;
        P.M = M;
        P.E = E;
        P.e0 = pass * pv16::MAX_EB;
        P.eb = std::min(pv16::MAX_EB, E - P.e0);
        P.ebp = (P.eb + 31) / 32 * 32;
        {
            const int box_rows = pl.pair ? P.ebp / 2 : pv16::MAX_EB;
            if (int rc = tc::make_tensor_map_f16(&maps[4], sh, pl.nsb * pl.Ep, pv16::TNS, box_rows)) return rc;
            if (int rc = tc::make_tensor_map_f16(&maps[5], sl, pl.nsb * pl.Ep, pv16::TNS, box_rows)) return rc;
        }
        P.n_tiles = static_cast<int>(pl.n_tiles);
        P.nsb = static_cast<int>(pl.nsb);
        P.kblocks = pl.kblocks;
        P.ksteps_last = pl.ksteps_last;
        P.stages = pl.stages;
        P.ep_rows = pl.Ep;
        static const int flush_blocks = [] {   // tuning knob
            const char* e = getenv("KMB_PV16_FLUSH_BLOCKS");
            const int v = e ? atoi(e) : pv16::kFlushBlocks;
            return v < 1 ? 1 : v;
        }();
        P.flush_blocks = flush_blocks;
        P.R = pl.waves.R;
        P.C = pl.waves.C;
        P.W = pl.waves.W;
        P.R_last = pl.waves.R_last;
        P.C_last = pl.waves.C_last;
        P.slots_per_wave = pl.waves.slots_per_wave;
        if (ev0 && pass == n_passes - 1) KMB_CUDA_CHECK(cudaEventRecord(ev0, stream));
        int rc;
        switch (kid * 2 + (norm ? 1 : 0)) {
            case 0: rc = launch_pv16<KMB_KERNEL_GAUSSIAN, false>(maps, P, pl.grid, pl.smem, pl.pair, stream); break;
            case 1: rc = launch_pv16<KMB_KERNEL_GAUSSIAN, true>(maps, P, pl.grid, pl.smem, pl.pair, stream); break;
            case 2: rc = launch_pv16<KMB_KERNEL_ABSOLUTE_EXPONENTIAL, false>(maps, P, pl.grid, pl.smem, pl.pair, stream); break;
            default: rc = launch_pv16<KMB_KERNEL_ABSOLUTE_EXPONENTIAL, true>(maps, P, pl.grid, pl.smem, pl.pair, stream); break;
        }
        if (rc) return rc;
        if (ev1 && pass == n_passes - 1) KMB_CUDA_CHECK(cudaEventRecord(ev1, stream));
        count_launch();
    }
    return KMB_OK;
}

}  // namespace kmb
