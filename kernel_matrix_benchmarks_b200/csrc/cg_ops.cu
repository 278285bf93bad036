// Fused vector steps of conjugate gradients on (K + lambda I) b = a (include/kmb_b200.h).
// The matvec is kmb_product_f32; these kernels do everything else in one pass each over the
// row shard, with deterministic reductions (per-block partials, combined in block order by the
// last block to arrive).  Scalars stay on the device: no host synchronisation in the loop.
#include "kmb_common.cuh"

namespace kmb {

constexpr int CG_THREADS = 256;
constexpr int CG_MAX_BLOCKS = 1024;
constexpr int CG_MAX_E = 16;

template <class T>
struct CgScratch {        // caller-provided, zero-initialised once (kmb_cg_scratch_bytes)
    T* partial;           // CG_MAX_BLOCKS * E
    unsigned int* counter;
};
template <class T>
__host__ __device__ inline CgScratch<T> carve(void* scratch) {
    CgScratch<T> s;
    s.counter = static_cast<unsigned int*>(scratch);
    s.partial = reinterpret_cast<T*>(static_cast<char*>(scratch) + 256);
    return s;
}
__device__ __forceinline__ float fma_t(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ double fma_t(double a, double b, double c) { return fma(a, b, c); }

// Sum `local[e]` over the grid into out[e], deterministically.
template <class T, int MAXE>
__device__ __forceinline__ void grid_reduce(T (&local)[MAXE], int E, CgScratch<T> s, T* out) {
    __shared__ T warp_part[CG_THREADS / 32][MAXE];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int e = 0; e < E; ++e) {
        T v = local[e];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) warp_part[warp][e] = v;
    }
    __syncthreads();
    if (threadIdx.x < E) {
        T v = T(0);
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
        T v = T(0);
        for (unsigned int b = 0; b < gridDim.x; ++b) v += __ldcg(&s.partial[b * E + threadIdx.x]);
        out[threadIdx.x] = v;
    }
}

template <class T>
__global__ void __launch_bounds__(CG_THREADS) cg_init_kernel(const T* __restrict__ a, T* __restrict__ x,
                                                             T* __restrict__ r, T* __restrict__ p,
                                                             T* rs, long long n, int E, CgScratch<T> s) {
    T local[CG_MAX_E];
    for (int e = 0; e < E; ++e) local[e] = T(0);
    for (long long i = blockIdx.x * static_cast<long long>(CG_THREADS) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * CG_THREADS)
        for (int e = 0; e < E; ++e) {
            const T v = a[i * E + e];
            x[i * E + e] = T(0);
            r[i * E + e] = v;
            p[i * E + e] = v;
            local[e] = fma_t(v, v, local[e]);
        }
    grid_reduce(local, E, s, rs);
}

template <class T>
__global__ void __launch_bounds__(CG_THREADS) cg_shift_dot_kernel(T* __restrict__ Ap, const T* __restrict__ p,
                                                                  T lambda, T* pAp, long long n, int E,
                                                                  CgScratch<T> s) {
    T local[CG_MAX_E];
    for (int e = 0; e < E; ++e) local[e] = T(0);
    for (long long i = blockIdx.x * static_cast<long long>(CG_THREADS) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * CG_THREADS)
        for (int e = 0; e < E; ++e) {
            const T pv = p[i * E + e];
            const T v = fma_t(lambda, pv, Ap[i * E + e]);
            Ap[i * E + e] = v;
            local[e] = fma_t(pv, v, local[e]);
        }
    grid_reduce(local, E, s, pAp);
}

template <class T>
__global__ void __launch_bounds__(CG_THREADS) cg_update_kernel(T* __restrict__ x, T* __restrict__ r,
                                                               const T* __restrict__ p, const T* __restrict__ Ap,
                                                               const T* rs, const T* pAp, T* rs_new,
                                                               long long n, int E, CgScratch<T> s) {
    T local[CG_MAX_E], alpha[CG_MAX_E];
    for (int e = 0; e < E; ++e) {
        local[e] = T(0);
        const T den = pAp[e];
        alpha[e] = den != T(0) ? rs[e] / den : T(0);  // r == 0 already: stay put
    }
    for (long long i = blockIdx.x * static_cast<long long>(CG_THREADS) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * CG_THREADS)
        for (int e = 0; e < E; ++e) {
            x[i * E + e] = fma_t(alpha[e], p[i * E + e], x[i * E + e]);
            const T rv = fma_t(-alpha[e], Ap[i * E + e], r[i * E + e]);
            r[i * E + e] = rv;
            local[e] = fma_t(rv, rv, local[e]);
        }
    grid_reduce(local, E, s, rs_new);
}

template <class T>
__global__ void __launch_bounds__(CG_THREADS) cg_direction_kernel(T* __restrict__ p, const T* __restrict__ r,
                                                                  const T* rs_new, const T* rs, long long n,
                                                                  int E) {
    T beta[CG_MAX_E];
    for (int e = 0; e < E; ++e) beta[e] = rs[e] != T(0) ? rs_new[e] / rs[e] : T(0);
    for (long long i = blockIdx.x * static_cast<long long>(CG_THREADS) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * CG_THREADS)
        for (int e = 0; e < E; ++e) p[i * E + e] = fma_t(beta[e], p[i * E + e], r[i * E + e]);
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


template <class T>
int cg_init(const T* a, T* x, T* r, T* p, T* rs, int64_t n, int E, void* scratch, void* stream) {
    if (int rc = cg_check(n, E, scratch)) return rc;
    cg_init_kernel<T><<<cg_grid(n), CG_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(a, x, r, p, rs, n, E, carve<T>(scratch));
    KMB_CUDA_CHECK(cudaGetLastError());
    count_launch();
    return KMB_OK;
}
template <class T>
int cg_shift_dot(T* Ap, const T* p, T lambda, T* pAp, int64_t n, int E, void* scratch, void* stream) {
    if (int rc = cg_check(n, E, scratch)) return rc;
    cg_shift_dot_kernel<T><<<cg_grid(n), CG_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(Ap, p, lambda, pAp, n, E, carve<T>(scratch));
    KMB_CUDA_CHECK(cudaGetLastError());
    count_launch();
    return KMB_OK;
}
template <class T>
int cg_update(T* x, T* r, const T* p, const T* Ap, const T* rs, const T* pAp, T* rs_new, int64_t n, int E, void* scratch, void* stream) {
    if (int rc = cg_check(n, E, scratch)) return rc;
    cg_update_kernel<T><<<cg_grid(n), CG_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(x, r, p, Ap, rs, pAp, rs_new, n, E, carve<T>(scratch));
    KMB_CUDA_CHECK(cudaGetLastError());
    count_launch();
    return KMB_OK;
}
template <class T>
int cg_direction(T* p, const T* r, const T* rs_new, const T* rs, int64_t n, int E, void* stream) {
    if (n < 0 || E < 1 || E > CG_MAX_E) return set_error(KMB_ERR_INVALID, "bad sizes n=%lld E=%d", (long long)n, E);
    cg_direction_kernel<T><<<cg_grid(n), CG_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(p, r, rs_new, rs, n, E);
    KMB_CUDA_CHECK(cudaGetLastError());
    count_launch();
    return KMB_OK;
}

}  // namespace kmb

using namespace kmb;

extern "C" {

size_t kmb_cg_scratch_bytes(void) { return 256 + sizeof(double) * CG_MAX_BLOCKS * CG_MAX_E; }

int kmb_cg_init_f32(const float* a, float* x, float* r, float* p_shard, float* rs_local, int64_t n, int E, void* scratch, void* stream) {
    return cg_init<float>(a, x, r, p_shard, rs_local, n, E, scratch, stream);
}
int kmb_cg_shift_dot_f32(float* Ap, const float* p_shard, float lambda, float* pAp_local, int64_t n, int E, void* scratch, void* stream) {
    return cg_shift_dot<float>(Ap, p_shard, lambda, pAp_local, n, E, scratch, stream);
}
int kmb_cg_update_f32(float* x, float* r, const float* p_shard, const float* Ap, const float* rs, const float* pAp,
                      float* rs_new_local, int64_t n, int E, void* scratch, void* stream) {
    return cg_update<float>(x, r, p_shard, Ap, rs, pAp, rs_new_local, n, E, scratch, stream);
}
int kmb_cg_direction_f32(float* p_shard, const float* r, const float* rs_new, const float* rs, int64_t n, int E, void* stream) {
    return cg_direction<float>(p_shard, r, rs_new, rs, n, E, stream);
}
int kmb_cg_init_f64(const double* a, double* x, double* r, double* p_shard, double* rs_local, int64_t n, int E, void* scratch, void* stream) {
    return cg_init<double>(a, x, r, p_shard, rs_local, n, E, scratch, stream);
}
int kmb_cg_shift_dot_f64(double* Ap, const double* p_shard, double lambda, double* pAp_local, int64_t n, int E, void* scratch, void* stream) {
    return cg_shift_dot<double>(Ap, p_shard, lambda, pAp_local, n, E, scratch, stream);
}
int kmb_cg_update_f64(double* x, double* r, const double* p_shard, const double* Ap, const double* rs, const double* pAp,
                      double* rs_new_local, int64_t n, int E, void* scratch, void* stream) {
    return cg_update<double>(x, r, p_shard, Ap, rs, pAp, rs_new_local, n, E, scratch, stream);
}
int kmb_cg_direction_f64(double* p_shard, const double* r, const double* rs_new, const double* rs, int64_t n, int E, void* stream) {
    return cg_direction<double>(p_shard, r, rs_new, rs, n, E, stream);
}

}  // extern "C"
