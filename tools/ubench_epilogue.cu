// Which pipe the instructions of the pv16 epilogue occupy (kprod_tensor_pv16.cu): warp-instructions per clock per SM
// sub-partition for each instruction alone and for pairs of them interleaved.  If two instructions share a pipe the
// interleaved loop takes the SUM of the two times, otherwise the maximum.  Prints one JSON object per line.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o kmb_ubench_epilogue tools/ubench_epilogue.cu
#include <cstdint>
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr int ITERS = 20000;
constexpr int ILP = 8;

__device__ __forceinline__ float ex2(float x) { float r; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float sqrtm(float x) { float r; asm volatile("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float rsqrtm(float x) { float r; asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ uint32_t f2fp(float lo, float hi) { uint32_t r; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }
__device__ __forceinline__ uint32_t f2fp_rz(float lo, float hi) { uint32_t r; asm volatile("cvt.rz.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }
__device__ __forceinline__ uint32_t f2fp_bf(float lo, float hi) { uint32_t r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }
__device__ __forceinline__ uint32_t lop(uint32_t a, uint32_t b) { uint32_t r; asm volatile("and.b32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b) { uint32_t r; asm volatile("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ float fmin3(float a, float b, float c) { float r; asm volatile("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float2 add2v(float2 a, float2 b) {
    float2 r;
    asm volatile("{\n.reg .b64 ra, rb, rc;\nmov.b64 ra, {%2, %3};\nmov.b64 rb, {%4, %5};\nadd.rn.f32x2 rc, ra, rb;\nmov.b64 {%0, %1}, rc;\n}"
                 : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ uint32_t hadd2(uint32_t a, uint32_t b) { uint32_t r; asm volatile("add.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ float h2f_lo(uint32_t a) { float r; asm volatile("{\n.reg .b16 l, h;\nmov.b32 {l, h}, %1;\ncvt.f32.f16 %0, l;\n}" : "=f"(r) : "r"(a)); return r; }

template <int MODE>
__global__ void __launch_bounds__(512) pipe_kernel(float* out, float seed, long long* cycles) {
    float2 a[ILP];
    uint32_t u[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { a[i] = make_float2(seed + i + threadIdx.x, seed * i + 1.f); u[i] = threadIdx.x * 7 + i; }
    const float2 c2 = make_float2(seed, seed * 0.5f);
    const long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (MODE == 0) { u[i] = f2fp(a[i].x, __uint_as_float(u[i])); }                    // 1 F2FP.F16.F32.PACK_AB
            if (MODE == 1) { a[i].x = ex2(a[i].x); }                                          // 1 MUFU.EX2
            if (MODE == 2) { a[i].x = sqrtm(a[i].x); }                                        // 1 MUFU.SQRT
            if (MODE == 3) { a[i].x = ex2(a[i].x); u[i] = f2fp(a[i].y, __uint_as_float(u[i])); }   // MUFU.EX2 + F2FP
            if (MODE == 4) { u[i] = lop(u[i], 0xffffe000u + i); }                             // 1 LOP3
            if (MODE == 5) { a[i].x = fmin3(a[i].x, a[i].y, c2.x); }                          // 1 FMNMX3
            if (MODE == 6) { a[i] = add2v(a[i], c2); }                                        // 1 FADD2
            if (MODE == 7) {   // the P phase per two sources: FADD2, 2 MUFU.EX2, FADD2, 2 LOP3, FADD2, 2 F2FP
                const float2 e = add2v(a[i], c2);
                const float2 pw = make_float2(ex2(e.x), ex2(e.y));
                a[(i + 1) % ILP] = add2v(a[(i + 1) % ILP], pw);
                const float2 h = make_float2(__uint_as_float(lop(__float_as_uint(pw.x), 0xffffe000u)), __uint_as_float(lop(__float_as_uint(pw.y), 0xffffe000u)));
                const float2 l = add2v(pw, make_float2(-h.x, -h.y));
                u[i] ^= f2fp(h.x, h.y) + f2fp(l.x, l.y);
            }
            if (MODE == 8) { u[i] = f2fp(a[i].x, __uint_as_float(u[i])); a[i].y = __uint_as_float(lop(__float_as_uint(a[i].y), 0xffffe000u + i)); }   // F2FP + LOP3
            if (MODE == 9) { u[i] = f2fp(a[i].x, __uint_as_float(u[i])); a[i] = add2v(a[i], c2); }   // F2FP + FADD2
            if (MODE == 10) { u[i] = prmt(u[i], __float_as_uint(a[i].x)); }                    // 1 PRMT
            if (MODE == 11) { u[i] = f2fp_rz(a[i].x, __uint_as_float(u[i])); }                 // F2FP .RZ
            if (MODE == 12) { u[i] = f2fp_bf(a[i].x, __uint_as_float(u[i])); }                 // F2FP.BF16
            if (MODE == 13) { u[i] = hadd2(u[i], 0x3c003c00u); }                               // HADD2
            if (MODE == 14) { a[i].x = h2f_lo(u[i]) + a[i].x; }                                // HADD2.F32 (half -> float) + FADD
            if (MODE == 15) { a[i].x = rsqrtm(a[i].x); }                                       // 1 MUFU.RSQ
            if (MODE == 16) { a[i].x = ex2(a[i].x); a[i].y = __uint_as_float(lop(__float_as_uint(a[i].y), 0xffffe000u + i)); }   // MUFU + LOP3
        }
    }
    const long long t1 = clock64();
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < ILP; ++i) r += a[i].x + a[i].y + __uint_as_float(u[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
int run(const char* name, int sms, int warps_per_smsp) {
    const int blocks = sms, threads = 128 * warps_per_smsp;
    float* out; long long* cyc;
    CHECK(cudaMalloc(&out, sizeof(float) * blocks * threads));
    CHECK(cudaMalloc(&cyc, sizeof(long long) * blocks));
    pipe_kernel<MODE><<<blocks, threads>>>(out, 1e-3f, cyc);
    CHECK(cudaDeviceSynchronize());
    pipe_kernel<MODE><<<blocks, threads>>>(out, 1e-3f, cyc);
    CHECK(cudaDeviceSynchronize());
    static long long h[4096]; CHECK(cudaMemcpy(h, cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost));
    double mean = 0; for (int i = 0; i < blocks; ++i) mean += h[i]; mean /= blocks;
    // cycles of one SM sub-partition per inner step of ONE warp (warps_per_smsp warps share it)
    const double per_step = mean / (static_cast<double>(ITERS) * ILP * warps_per_smsp);
    printf("{\"mode\": \"%s\", \"warps_per_smsp\": %d, \"smsp_cycles_per_warp_step\": %.2f}\n", name, warps_per_smsp, per_step);
    cudaFree(out); cudaFree(cyc);
    return 0;
}

int main() {
    cudaDeviceProp p; CHECK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    for (int w : {2, 4}) {
        if (run<0>("F2FP.F16.F32.PACK_AB", sms, w)) return 1;
        if (run<11>("F2FP.RZ", sms, w)) return 1;
        if (run<12>("F2FP.BF16", sms, w)) return 1;
        if (run<1>("MUFU.EX2", sms, w)) return 1;
        if (run<2>("MUFU.SQRT", sms, w)) return 1;
        if (run<15>("MUFU.RSQ", sms, w)) return 1;
        if (run<3>("MUFU.EX2 + F2FP", sms, w)) return 1;
        if (run<4>("LOP3", sms, w)) return 1;
        if (run<5>("FMNMX3", sms, w)) return 1;
        if (run<6>("FADD2", sms, w)) return 1;
        if (run<10>("PRMT", sms, w)) return 1;
        if (run<13>("HADD2", sms, w)) return 1;
        if (run<14>("HADD2.F32 + FADD", sms, w)) return 1;
        if (run<8>("F2FP + LOP3", sms, w)) return 1;
        if (run<9>("F2FP + FADD2", sms, w)) return 1;
        if (run<16>("MUFU.EX2 + LOP3", sms, w)) return 1;
        if (run<7>("P phase: FADD2 2MUFU FADD2 2LOP3 FADD2 2F2FP", sms, w)) return 1;
    }
    return 0;
}
