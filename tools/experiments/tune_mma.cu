// Tuning harness for kprod_mma_kernel (exponent on mma.sync TF32): times configurations of the D = 3
// Gaussian product (general and symmetric) on one GPU and prints one line each.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o tune_mma tools/experiments/tune_mma.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "kprod_mma.cuh"

namespace kmb {
int set_error(int code, const char*, ...) { return code; }
void count_launch(int) {}
}  // namespace kmb
using namespace kmb;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

template <class C>
void run(const char* name, long long N, const float* y, const float* b, float* out, int sms) {
    using Map = typename C::Map;
    const long long nsb = (N + C::SB - 1) / C::SB, M_pad = nsb * C::SB;
    const long long n_tiles = (N + C::TILE_ROWS - 1) / C::TILE_ROWS;
    float4* recb; float* wv; DirectStats* stats; float* box; float *rowsum, *rowpart, *colpart = nullptr;
    CK(cudaMalloc(&recb, M_pad * 64)); CK(cudaMalloc(&wv, M_pad * 4));
    CK(cudaMalloc(&stats, sizeof(DirectStats))); CK(cudaMemset(stats, 0, sizeof(DirectStats)));
    CK(cudaMalloc(&box, sizeof(float) * STATS_MAX_BLOCKS * 32));
    CK(cudaFuncSetAttribute(kprod_mma_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kprod_mma_kernel<C>, C::THREADS, C::SMEM_BYTES));
    cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, kprod_mma_kernel<C>));
    const long long units = Map::prefix(n_tiles, nsb);
    long long grid = (long long)sms * per_sm;
    if (grid > units) grid = units;
    CK(cudaMalloc(&rowsum, n_tiles * C::TILE_ROWS * 4));
    CK(cudaMalloc(&rowpart, grid * 2 * C::TILE_ROWS * 4));
    if (C::SYM) CK(cudaMalloc(&colpart, (size_t)n_tiles * M_pad * 4));
    const float scale = 1.2011224087864498f;
    direct_stats_kernel<<<64, STATS_THREADS>>>(y, N, y, N, 3, box, stats, 1);
    pack_mma_kernel<<<(unsigned)((M_pad + 255) / 256), 256>>>(y, b, recb, wv, stats, N, M_pad, 3, scale);
    MmaParams P{};
    P.stats = stats; P.x = y; P.recb = recb; P.wv = wv; P.rowsum = rowsum; P.rowpart = rowpart; P.colpart = colpart; P.out = out;
    P.N = N; P.M = N; P.M_pad = M_pad; P.unit_begin = 0; P.unit_end = units; P.D = 3; P.n_tiles = (int)n_tiles; P.nsb = (int)nsb;
    P.grid = (int)grid; P.xscale = scale;
    cudaEvent_t e0, e1, e2; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreate(&e2));
    for (int i = 0; i < 2; ++i) kprod_mma_kernel<C><<<(int)grid, C::THREADS, C::SMEM_BYTES>>>(P);
    CK(cudaDeviceSynchronize());
    const int reps = 3;
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; ++i) kprod_mma_kernel<C><<<(int)grid, C::THREADS, C::SMEM_BYTES>>>(P);
    CK(cudaEventRecord(e1));
    mma_combine_kernel<C><<<(unsigned)((N + 255) / 256), 256>>>(P);
    CK(cudaEventRecord(e2));
    CK(cudaDeviceSynchronize());
    float ms, ms_c; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= reps; CK(cudaEventElapsedTime(&ms_c, e1, e2));
    std::vector<float> h(N); CK(cudaMemcpy(h.data(), out, N * 4, cudaMemcpyDeviceToHost));
    double cs = 0; for (long long i = 0; i < N; ++i) cs += h[i];
    const double gp = (double)N * N / ((ms + ms_c) * 1e-3) / 1e9;
    const double kev = (double)units * C::TILE_ROWS * C::SB / (ms * 1e-3);
    printf("%-30s regs=%3d ctas/sm=%d grid=%4lld smem=%6d  main %8.3f ms + combine %6.3f ms  %7.1f Gpairs/s  %.2f k-evals/clk/SM@1965  checksum=%.6e out[5]=%.6e\n",
           name, fa.numRegs, per_sm, grid, C::SMEM_BYTES, ms, ms_c, gp, kev / (sms * 1.965e9), cs, (double)h[5]);
    cudaFree(recb); cudaFree(wv); cudaFree(rowsum); cudaFree(rowpart); if (colpart) cudaFree(colpart); cudaFree(stats); cudaFree(box);
}

int main(int argc, char** argv) {
    const long long N = argc > 1 ? atoll(argv[1]) : 262144;
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    std::vector<float> hy(N * 3), hb(N);
    srand(1);
    for (auto& v : hy) v = rand() / (float)RAND_MAX;
    for (auto& v : hb) v = rand() / (float)RAND_MAX - 0.5f;
    float *y, *b, *out;
    CK(cudaMalloc(&y, N * 12)); CK(cudaMalloc(&b, N * 4)); CK(cudaMalloc(&out, N * 4));
    CK(cudaMemcpy(y, hy.data(), N * 12, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(b, hb.data(), N * 4, cudaMemcpyHostToDevice));
    printf("N=%lld on %s (%d SMs)   (reference checksums of tune_direct / tune_sym: -2.353711e+06 at 262144, -9.753223e+06 at 1000000)\n", N, p.name, p.multiProcessorCount);
    const int sms = p.multiProcessorCount;
    //   SYM POLY WARPS SB STAGES MINB MT
#define RUN(...) run<MmaCfg<__VA_ARGS__>>(#__VA_ARGS__, N, y, b, out, sms)
    RUN(false, 0, 16, 256, 4, 1, 4);
    RUN(false, 4, 16, 256, 4, 1, 4);
    RUN(false, 0, 16, 256, 4, 2, 2);
    RUN(false, 4, 16, 256, 4, 2, 2);
    RUN(false, 2, 16, 256, 4, 2, 2);
    RUN(false, 8, 16, 256, 4, 2, 2);
    RUN(false, 4, 8, 256, 4, 4, 2);
    RUN(false, 4, 16, 256, 3, 2, 2);
    RUN(true, 0, 16, 256, 4, 2, 2);
    RUN(true, 4, 16, 256, 4, 2, 2);
    RUN(true, 8, 16, 256, 4, 2, 2);
    RUN(true, 0, 8, 256, 4, 4, 2);
    RUN(true, 8, 8, 256, 4, 4, 2);
    return 0;
}
