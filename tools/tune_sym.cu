// Tuning harness for kprod_sym_kernel (same_points Gaussian product, D = 3): times several
// configurations on one GPU (device-resident inputs, CUDA events) and prints one line each.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o tune_sym tools/tune_sym.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../kernel_matrix_benchmarks_b200/csrc/kprod_sym.cuh"

namespace kmb {
int set_error(int code, const char*, ...) { return code; }
void count_launch(int) {}
}  // namespace kmb
using namespace kmb;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

template <class C>
void run(const char* name, long long N, const float* y, const float* b, float* out, int sms, double ref_checksum) {
    const long long nsb = (N + C::SB - 1) / C::SB, N_pad = nsb * C::SB;
    const long long n_tiles = (N + C::TILE_ROWS - 1) / C::TILE_ROWS;
    float2* rec; DirectStats* stats; float* box; float *rowseg, *rowpart, *colpart;
    CK(cudaMalloc(&rec, N_pad * C::RECV * 16));
    CK(cudaMalloc(&stats, sizeof(DirectStats))); CK(cudaMemset(stats, 0, sizeof(DirectStats)));
    CK(cudaMalloc(&box, sizeof(float) * STATS_MAX_BLOCKS * 32));
    CK(cudaFuncSetAttribute(kprod_sym_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kprod_sym_kernel<C>, C::THREADS, C::SMEM_BYTES));
    cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, kprod_sym_kernel<C>));
    long long grid = (long long)sms * per_sm;
    SymGeom geom;
    sym_build_geom<C::TB>(n_tiles, nsb, grid, &geom);
    const long long units = geom.strip_prefix[geom.n_strips];
    if (grid > units) grid = units;
    CK(cudaMalloc(&rowseg, (size_t)geom.seg_prefix[geom.n_strips] * C::TILE_ROWS * 4));
    CK(cudaMalloc(&rowpart, grid * 2 * C::TILE_ROWS * 4));
    CK(cudaMalloc(&colpart, (size_t)(grid + geom.n_strips) * geom.Wb * C::SB * 4));
    const float scale = 1.2011224087864498f;
    direct_stats_kernel<<<64, STATS_THREADS>>>(y, N, y, N, 3, box, stats, 1);
    PackLayout L{N_pad, 0, C::RECV * 2};
    pack_sources_kernel<<<(unsigned)((N_pad + 255) / 256), 256>>>(y, b, rec, stats, N, 3, 1, C::DP, 1, L, L, 0, scale);
    SymParams P{};
    P.stats = stats; P.rec = (const float4*)rec; P.rowseg = rowseg; P.rowpart = rowpart; P.colpart = colpart; P.out = out;
    P.N = N; P.M = N; P.unit_begin = 0; P.unit_end = units; P.piece_floats = (long long)geom.Wb * C::SB; P.grid = (int)grid;
    P.seg_base = 0; P.strip_base = 0; P.g = geom;
    cudaEvent_t e0, e1, e2; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreate(&e2));
    for (int i = 0; i < 2; ++i) kprod_sym_kernel<C><<<(int)grid, C::THREADS, C::SMEM_BYTES>>>(P);
    CK(cudaDeviceSynchronize());
    const int reps = 3;
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; ++i) kprod_sym_kernel<C><<<(int)grid, C::THREADS, C::SMEM_BYTES>>>(P);
    CK(cudaEventRecord(e1));
    sym_combine_kernel<C><<<(unsigned)((N + 255) / 256), 256>>>(P);
    CK(cudaEventRecord(e2));
    CK(cudaDeviceSynchronize());
    float ms, ms_c; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= reps; CK(cudaEventElapsedTime(&ms_c, e1, e2));
    std::vector<float> h(N); CK(cudaMemcpy(h.data(), out, N * 4, cudaMemcpyDeviceToHost));
    double cs = 0; for (long long i = 0; i < N; ++i) cs += h[i];
    const double gp = (double)N * N / ((ms + ms_c) * 1e-3) / 1e9;
    const double kev = (double)units * C::TILE_ROWS * C::SB / (ms * 1e-3);
    printf("%-28s regs=%3d ctas/sm=%d grid=%4lld smem=%6d  main %8.3f ms + combine %6.3f ms  %7.1f Gpairs/s  %.2f k-evals/clk/SM@1965  checksum=%.6e (ref %.6e)\n",
           name, fa.numRegs, per_sm, grid, C::SMEM_BYTES, ms, ms_c, gp, kev / (sms * 1.965e9), cs, ref_checksum);
    cudaFree(rec); cudaFree(rowseg); cudaFree(rowpart); cudaFree(colpart); cudaFree(stats); cudaFree(box);
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
    printf("N=%lld on %s (%d SMs)\n", N, p.name, p.multiProcessorCount);
    const int sms = p.multiProcessorCount;
    //   DP KID FORM POLY MINB CH CONSUMERS R STAGES   (product form of the Gaussian kernel: KID 0, FORM 1)
#define RUN(...) run<SymCfg<__VA_ARGS__>>(#__VA_ARGS__, N, y, b, out, sms, 0.0)
    RUN(3, 0, 1, 0, 2, 8, 512, 4, 4);
    RUN(3, 0, 1, 0, 2, 16, 512, 4, 4);
    RUN(3, 0, 1, 0, 1, 16, 512, 8, 4);
    RUN(3, 0, 1, 0, 1, 8, 512, 8, 4);
    RUN(3, 0, 1, 32, 2, 16, 512, 4, 4);
    RUN(3, 0, 1, 32, 1, 16, 512, 8, 4);
    return 0;
}
