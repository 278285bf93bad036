// Tuning harness for kprod_direct_kernel: times several configurations of the D=3, E=1 Gaussian
// product on one GPU (device-resident inputs, CUDA events) and prints one line per configuration.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o tune_direct tools/tune_direct.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../kernel_matrix_benchmarks_b200/csrc/kprod_direct.cuh"

namespace kmb {
int set_error(int code, const char*, ...) { return code; }
void count_launch(int) {}
}  // namespace kmb
using namespace kmb;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

template <class C>
void run(const char* name, long long N, long long M, const float* x, const float* y, const float* b, float* out, int sms) {
    const long long nsb = (M + C::SB - 1) / C::SB, M_pad = nsb * C::SB;
    const long long n_tiles = (N + C::TILE_ROWS - 1) / C::TILE_ROWS;
    float2* rec; float* partial; int* counters; DirectStats* stats; float* box;
    CK(cudaMalloc(&rec, M_pad * C::RECV * 16));
    CK(cudaMalloc(&stats, sizeof(DirectStats))); CK(cudaMemset(stats, 0, sizeof(DirectStats)));
    CK(cudaMalloc(&box, sizeof(float) * STATS_MAX_BLOCKS * 32));
    CK(cudaFuncSetAttribute(kprod_direct_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kprod_direct_kernel<C>, C::THREADS, C::SMEM_BYTES));
    cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, kprod_direct_kernel<C>));
    long long grid = (long long)sms * per_sm;
    if (grid > n_tiles * nsb) grid = n_tiles * nsb;
    CK(cudaMalloc(&partial, (size_t)grid * 2 * C::TILE_ROWS * C::PS * 4));
    CK(cudaMalloc(&counters, n_tiles * 4));
    CK(cudaMemset(counters, 0, n_tiles * 4));
    const float scale = 1.2011224087864498f;
    direct_stats_kernel<<<64, STATS_THREADS>>>(x, N, y, M, 3, box, stats, C::FORM);
    PackLayout L{M_pad, C::WCOL, C::RECV * 2};
    pack_sources_kernel<<<(unsigned)((M_pad + 255) / 256), 256>>>(y, b, rec, stats, M, 3, 1, C::DP, C::EP, L, L, 0, scale);
    DirectParams P{};
    P.x = x; P.stats = stats; P.rec = (const float4*)rec; P.out = out; P.partial = partial; P.tile_counter = counters;
    P.N = N; P.M = M; P.row_offset = 0; P.D = 3; P.E = 1; P.e0 = 0; P.n_tiles = (int)n_tiles; P.n_src_blocks = (int)nsb; P.xscale = scale;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 2; ++i) kprod_direct_kernel<C><<<(int)grid, C::THREADS, C::SMEM_BYTES>>>(P);
    CK(cudaDeviceSynchronize());
    const int reps = 3;
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; ++i) kprod_direct_kernel<C><<<(int)grid, C::THREADS, C::SMEM_BYTES>>>(P);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= reps;
    std::vector<float> h(N); CK(cudaMemcpy(h.data(), out, N * 4, cudaMemcpyDeviceToHost));
    double cs = 0; for (long long i = 0; i < N; ++i) cs += h[i];
    const double gp = (double)N * M / (ms * 1e-3) / 1e9;
    printf("%-40s regs=%3d ctas/sm=%d grid=%4lld smem=%6d  %8.3f ms  %7.1f Gpairs/s  %.2f pairs/clk/SM@1965  checksum=%.6e\n", name,
           fa.numRegs, per_sm, grid, C::SMEM_BYTES, ms, gp, gp * 1e9 / (sms * 1.965e9), cs);
    cudaFree(rec); cudaFree(partial); cudaFree(counters); cudaFree(stats); cudaFree(box);
}

int main(int argc, char** argv) {
    const long long N = argc > 1 ? atoll(argv[1]) : 262144, M = N;
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    std::vector<float> hy(M * 3), hb(M);
    srand(1);
    for (auto& v : hy) v = rand() / (float)RAND_MAX;
    for (auto& v : hb) v = rand() / (float)RAND_MAX - 0.5f;
    float *y, *b, *out;
    CK(cudaMalloc(&y, M * 12)); CK(cudaMalloc(&b, M * 4)); CK(cudaMalloc(&out, N * 4));
    CK(cudaMemcpy(y, hy.data(), M * 12, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(b, hb.data(), M * 4, cudaMemcpyHostToDevice));
    printf("N=M=%lld on %s (%d SMs)\n", N, p.name, p.multiProcessorCount);
    const int sms = p.multiProcessorCount;
    //   DP EP R KID NORM FORM CONS UNR MINB STAGES POLY
#define RUN(...) run<DirectCfg<__VA_ARGS__>>(#__VA_ARGS__, N, M, y, y, b, out, sms)
    RUN(3, 1, 4, 0, false, 1, 512, 4, 2, 4, 0);
    RUN(3, 1, 4, 0, false, 1, 512, 4, 2, 4, 8);
    RUN(3, 1, 4, 0, false, 1, 512, 4, 2, 4, 4);
    RUN(3, 1, 4, 0, false, 1, 512, 2, 2, 4, 4);
    RUN(3, 1, 4, 0, false, 1, 512, 4, 2, 4, 2);
    RUN(3, 1, 4, 0, false, 1, 512, 4, 1, 4, 4);
    RUN(3, 1, 8, 0, false, 1, 512, 4, 1, 4, 4);
    RUN(3, 1, 8, 0, false, 1, 512, 2, 1, 4, 4);
    RUN(3, 1, 8, 0, false, 1, 512, 2, 1, 4, 8);
    RUN(3, 1, 8, 0, false, 1, 256, 4, 2, 4, 4);
    RUN(3, 1, 6, 0, false, 1, 512, 4, 1, 4, 4);
    RUN(3, 1, 6, 0, false, 1, 512, 4, 1, 4, 3);
    RUN(3, 1, 4, 0, false, 1, 768, 4, 1, 4, 4);
    RUN(3, 1, 4, 0, false, 1, 992, 4, 1, 4, 4);
    RUN(3, 1, 4, 0, false, 1, 512, 8, 2, 4, 4);
    RUN(3, 1, 4, 0, false, 1, 512, 8, 2, 4, 16);
    return 0;
}
