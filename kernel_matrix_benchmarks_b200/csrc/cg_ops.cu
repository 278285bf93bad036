// Fused vector steps of conjugate gradients on (K + lambda I) b = a (include/kmb_b200.h).
// The matvec is kmb_product_f32; these kernels do everything else in one pass each over the
// row shard, with deterministic reductions (per-block partials, combined in block order by the
// last block to arrive).  Scalars stay on the device: no host synchronisation in the loop.
#include "kmb_common.cuh"

namespace kmb {

constexpr int CG_THREADS = 256;
constexpr int CG_MAX_BLOCKS = 1024;
constexpr int CG_MAX_E = 16;

struct CgScratch {        // caller-provided, zero-initialised once (kmb_cg_scratch_bytes)
    float* partial;       // CG_MAX_BLOCKS * E
    unsigned int* counter;
};
__host__ __device__ inline CgScratch carve(void* scratch) {
    CgScratch s;
    s.counter = static_cast<unsigned int*>(scratch);
    s.partial = reinterpret_cast<float*>(static_cast<char*>(scratch) + 256);
    return s;
}

// Sum `local[e]` over the grid into out[e], deterministically.
template <int MAXE>
__device__ __forceinline__ void grid_reduce(float (&local)[MAXE], int E, CgScratch s, float* out) {
    __shared__ float warp_part[CG_THREADS / 32][MAXE];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int e = 0; e < E; ++e) {
        float v = local[e];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) warp_part[warp][e] = v;
    }
    __syncthreads();
    if (threadIdx.x < E) {
        float v = 0.f;
        for (int w = 0; w < CG_THREADS / 32; ++w) v += warp_part[w][threadIdx.x];
        s.partial[blockIdx.x * E + threadIdx.x] = v;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int old = atomicAdd(s.counter, 1u);
        is_last = (old == gridDim.x - 1);
        if (is_last) *s.counter = 0;
    }
    __syncthreads();
    if (is_last && threadIdx.x < E) {
        __threadfence();
        float v = 0.f;
        for (unsigned int b = 0; b < gridDim.x; ++b) v += __ldcg(&s.partial[b * E + threadIdx.x]);
        out[threadIdx.x] = v;
    }
}

__global__ void __launch_bounds__(CG_THREADS) cg_init_kernel(const float* __restrict__ a, float* __restrict__ x,
                                                             float* __restrict__ r, float* __restrict__ p,
                                                             float* rs, long long n, int E, CgScratch s) {
    float local[CG_MAX_E];
    for (int e = 0; e < E; ++e) local[e] = 0.f;
    for (long long i = blockIdx.x * static_cast<long long>(CG_THREADS) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * CG_THREADS)
        for (int e = 0; e < E; ++e) {
            const float v = a[i * E + e];
            x[i * E + e] = 0.f;
            r[i * E + e] = v;
            p[i * E + e] = v;
            local[e] = fmaf(v, v, local[e]);
        }
    grid_reduce(local, E, s, rs);
}

__global__ void __launch_bounds__(CG_THREADS) cg_shift_dot_kernel(float* __restrict__ Ap, const float* __restrict__ p,
                                                                  float lambda, float* pAp, long long n, int E,
                                                                  CgScratch s) {
    float local[CG_MAX_E];
    for (int e = 0; e < E; ++e) local[e] = 0.f;
    for (long long i = blockIdx.x * static_cast<long long>(CG_THREADS) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * CG_THREADS)
        for (int e = 0; e < E; ++e) {
            const float pv = p[i * E + e];
            const float v = fmaf(lambda, pv, Ap[i * E + e]);
            Ap[i * E + e] = v;
            local[e] = fmaf(pv, v, local[e]);
        }
    grid_reduce(local, E, s, pAp);
}

__global__ void __launch_bounds__(CG_THREADS) cg_update_kernel(float* __restrict__ x, float* __restrict__ r,
                                                               const float* __restrict__ p, const float* __restrict__ Ap,
                                                               const float* rs, const float* pAp, float* rs_new,
                                                               long long n, int E, CgScratch s) {
    float local[CG_MAX_E], alpha[CG_MAX_E];
    for (int e = 0; e < E; ++e) {
        local[e] = 0.f;
        const float den = pAp[e];
        alpha[e] = den != 0.f ? rs[e] / den : 0.f;  // r == 0 already: stay put
    }
    for (long long i = blockIdx.x * static_cast<long long>(CG_THREADS) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * CG_THREADS)
        for (int e = 0; e < E; ++e) {
            x[i * E + e] = fmaf(alpha[e], p[i * E + e], x[i * E + e]);
            const float rv = fmaf(-alpha[e], Ap[i * E + e], r[i * E + e]);
            r[i * E + e] = rv;
            local[e] = fmaf(rv, rv, local[e]);
        }
    grid_reduce(local, E, s, rs_new);
}

__global__ void __launch_bounds__(CG_THREADS) cg_direction_kernel(float* __restrict__ p, const float* __restrict__ r,
                                                                  const float* rs_new, const float* rs, long long n,
                                                                  int E) {
    float beta[CG_MAX_E];
    for (int e = 0; e < E; ++e) beta[e] = rs[e] != 0.f ? rs_new[e] / rs[e] : 0.f;
    for (long long i = blockIdx.x * static_cast<long long>(CG_THREADS) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * CG_THREADS)
        for (int e = 0; e < E; ++e) p[i * E + e] = fmaf(beta[e], p[i * E + e], r[i * E + e]);
}

static int cg_grid(long long n) {
    long long g = (n + CG_THREADS * 4 - 1) / (CG_THREADS * 4);
    return static_cast<int>(g < 1 ? 1 : (g > CG_MAX_BLOCKS ? CG_MAX_BLOCKS : g));
}
static int cg_check(long long n, int E, const void* scratch) {
    if (n < 0 || E < 1) return set_error(KMB_ERR_INVALID, "bad sizes n=%lld E=%d", n, E);
    if (E > CG_MAX_E) return set_error(KMB_ERR_UNSUPPORTED, "CG supports E <= %d right-hand sides per call (got %d)", CG_MAX_E, E);
    if (!scratch) return set_error(KMB_ERR_INVALID, "scratch is NULL");
    return KMB_OK;
}

}  // namespace kmb

using namespace kmb;

extern "C" {

size_t kmb_cg_scratch_bytes(void) { return 256 + sizeof(float) * CG_MAX_BLOCKS * CG_MAX_E; }

int kmb_cg_init_f32(const float* a, float* x, float* r, float* p_shard, float* rs_local, int64_t n, int E,
                    void* scratch, void* stream) {
    if (int rc = cg_check(n, E, scratch)) return rc;
    cg_init_kernel<<<cg_grid(n), CG_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(a, x, r, p_shard, rs_local, n, E,
                                                                                     carve(scratch));
    KMB_CUDA_CHECK(cudaGetLastError());
    count_launch();
    return KMB_OK;
}
int kmb_cg_shift_dot_f32(float* Ap, const float* p_shard, float lambda, float* pAp_local, int64_t n, int E,
                         void* scratch, void* stream) {
    if (int rc = cg_check(n, E, scratch)) return rc;
    cg_shift_dot_kernel<<<cg_grid(n), CG_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(Ap, p_shard, lambda, pAp_local,
                                                                                          n, E, carve(scratch));
    KMB_CUDA_CHECK(cudaGetLastError());
    count_launch();
    return KMB_OK;
}
int kmb_cg_update_f32(float* x, float* r, const float* p_shard, const float* Ap, const float* rs, const float* pAp,
                      float* rs_new_local, int64_t n, int E, void* scratch, void* stream) {
    if (int rc = cg_check(n, E, scratch)) return rc;
    cg_update_kernel<<<cg_grid(n), CG_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(x, r, p_shard, Ap, rs, pAp,
                                                                                       rs_new_local, n, E, carve(scratch));
    KMB_CUDA_CHECK(cudaGetLastError());
    count_launch();
    return KMB_OK;
}
int kmb_cg_direction_f32(float* p_shard, const float* r, const float* rs_new, const float* rs, int64_t n, int E,
                         void* stream) {
    if (n < 0 || E < 1 || E > CG_MAX_E) return set_error(KMB_ERR_INVALID, "bad sizes n=%lld E=%d", (long long)n, E);
    cg_direction_kernel<<<cg_grid(n), CG_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(p_shard, r, rs_new, rs, n, E);
    KMB_CUDA_CHECK(cudaGetLastError());
    count_launch();
    return KMB_OK;
}

}  // extern "C"
