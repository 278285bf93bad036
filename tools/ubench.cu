// Pipe micro-benchmarks for the roofline denominators (SURVEY.md section 8d marks them as
// assumptions to verify on the box): per-SM per-clock throughput of FFMA, packed FFMA2,
// MUFU.EX2 and the instruction mix of the D=3 Gaussian pair.  Prints one JSON object.
#include <cstdio>
#include <cuda_runtime.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr int ITERS = 200000;
constexpr int ILP = 8;

template <int MODE>
__global__ void __launch_bounds__(512) pipe_kernel(float* out, float seed, long long* cycles) {
    float a[ILP], b = seed, c = seed * 0.5f;
    float2 a2[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { a[i] = seed + i + threadIdx.x; a2[i] = make_float2(a[i], a[i] + 1.f); }
    const float2 b2 = make_float2(b, b), c2 = make_float2(c, c);
    const long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (MODE == 0) a[i] = fmaf(a[i], b, c);                                   // FFMA
            if (MODE == 1) a2[i] = __ffma2_rn(a2[i], b2, c2);                         // FFMA2
            if (MODE == 2) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));   // MUFU.EX2
            if (MODE == 3) {  // one Gaussian pair-pair in packed form: 3 FADD2 1 FMUL2 3 FFMA2 2 MUFU
                float2 d0 = __fadd2_rn(a2[i], b2), d1 = __fadd2_rn(a2[i], c2), d2 = __fadd2_rn(a2[i], a2[(i + 1) % ILP]);
                float2 s = __fmul2_rn(d0, d0);
                s = __ffma2_rn(d1, d1, s);
                s = __ffma2_rn(d2, d2, s);
                float ex, ey;
                asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(-s.x));
                asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(ey) : "f"(-s.y));
                a2[i] = __ffma2_rn(make_float2(ex, ey), b2, a2[i]);
            }
            if (MODE == 5) {  // expanded form, packed: 3 FFMA2 1 FADD2 1 FFMA2 2 MUFU per 2 pairs
                float2 s = __ffma2_rn(a2[i], b2, c2);
                s = __ffma2_rn(a2[(i + 1) % ILP], c2, s);
                s = __ffma2_rn(a2[(i + 2) % ILP], b2, s);
                s = __fadd2_rn(s, c2);
                float ex, ey;
                asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(-s.x));
                asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(ey) : "f"(-s.y));
                a2[i] = __ffma2_rn(make_float2(ex, ey), b2, a2[i]);
            }
            if (MODE == 4) {  // same pair in scalar form: 3 FADD 1 FMUL 3 FFMA 1 MUFU
                float d0 = a[i] + b, d1 = a[i] + c, d2 = a[i] + a[(i + 1) % ILP];
                float s = d0 * d0;
                s = fmaf(d1, d1, s);
                s = fmaf(d2, d2, s);
                float ex;
                asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(-s));
                a[i] = fmaf(ex, b, a[i]);
            }
        }
    }
    const long long t1 = clock64();
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < ILP; ++i) r += a[i] + a2[i].x + a2[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
int run(const char* name, double lane_ops_per_inner, int sms, int khz, bool last) {
    const int blocks = sms * 2, threads = 512;  // 32 warps / SM
    float* out; long long* cyc;
    CHECK(cudaMalloc(&out, sizeof(float) * blocks * threads));
    CHECK(cudaMalloc(&cyc, sizeof(long long) * blocks));
    cudaEvent_t e0, e1;
    CHECK(cudaEventCreate(&e0)); CHECK(cudaEventCreate(&e1));
    pipe_kernel<MODE><<<blocks, threads>>>(out, 1e-3f, cyc);
    CHECK(cudaDeviceSynchronize());
    CHECK(cudaEventRecord(e0));
    pipe_kernel<MODE><<<blocks, threads>>>(out, 1e-3f, cyc);
    CHECK(cudaEventRecord(e1));
    CHECK(cudaDeviceSynchronize());
    float ms; CHECK(cudaEventElapsedTime(&ms, e0, e1));
    long long h[4096]; CHECK(cudaMemcpy(h, cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost));
    double mean = 0; for (int i = 0; i < blocks; ++i) mean += h[i]; mean /= blocks;
    // per SM: 2 CTAs * 512 threads * ITERS * ILP inner steps
    const double inner = 2.0 * threads * ITERS * ILP;
    // two estimates: per-CTA clock64 spans (biased high when CTAs do not overlap perfectly) and
    // event time at the nominal clock (biased low by launch overhead and any clock droop)
    printf("  \"%s\": {\"per_sm_per_clk_clock64\": %.2f, \"per_sm_per_clk_event_at_nominal\": %.2f, \"ms\": %.3f, \"eff_mhz\": %.0f}%s\n", name,
           inner * lane_ops_per_inner / mean, inner * lane_ops_per_inner / (ms * 1e-3 * khz * 1e3), ms, mean / (ms * 1e-3) / 1e6, last ? "" : ",");
    cudaFree(out); cudaFree(cyc);
    return 0;
}

int main() {
    cudaDeviceProp p; CHECK(cudaGetDeviceProperties(&p, 0));
    int khz = 0; CHECK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
    printf("{\n  \"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d,\n", p.name, p.multiProcessorCount, khz);
    // units: lane-level operations (one FFMA2 = 2 FMA lane-ops; MODE 3 = 2 pairs, MODE 4 = 1 pair)
    if (run<0>("ffma_lane_ops", 1, p.multiProcessorCount, khz, false)) return 1;
    if (run<1>("ffma2_lane_ops", 2, p.multiProcessorCount, khz, false)) return 1;
    if (run<2>("mufu_ex2_lane_ops", 1, p.multiProcessorCount, khz, false)) return 1;
    if (run<3>("gauss_pairs_packed", 2, p.multiProcessorCount, khz, false)) return 1;
    if (run<4>("gauss_pairs_scalar", 1, p.multiProcessorCount, khz, false)) return 1;
    if (run<5>("gauss_pairs_expanded_packed", 2, p.multiProcessorCount, khz, true)) return 1;
    printf("}\n");
    return 0;
}
