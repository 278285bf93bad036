// Instantiations and launchers of the symmetric (same_points) Gaussian product, D <= 3, E = 1.
#include "kprod_sym.cuh"

namespace kmb {

namespace {

template <class C>
int sym_grid_of(int sms, int* grid) {
    int per_sm = 0;
    if (int rc = resident_ctas(reinterpret_cast<const void*>(&kprod_sym_kernel<C>), C::THREADS, C::SMEM_BYTES, &per_sm)) return rc;
    *grid = sms * (per_sm > 2 ? 2 : per_sm);
    return KMB_OK;
}

template <class C>
int sym_launch_of(SymParams P, cudaStream_t stream) {
    const long long units = P.unit_end - P.unit_begin;
    if (units > 0) {
        const int grid = static_cast<int>(std::min<long long>(P.grid, units));
        P.grid = grid;
        kprod_sym_kernel<C><<<grid, C::THREADS, C::SMEM_BYTES, stream>>>(P);
        KMB_CUDA_CHECK(cudaGetLastError());
        count_launch();
    }
    sym_combine_kernel<C><<<static_cast<unsigned>((P.N + 255) / 256), 256, 0, stream>>>(P);
    KMB_CUDA_CHECK(cudaGetLastError());
    count_launch();
    return KMB_OK;
}

// Production shape (tools/tune_sym.cu on B200, profiles/r1_tune_sym.log): 512 consumer threads x 8 rows,
// one CTA per SM, butterflies of 16 sources, every exponential on the MUFU pipe -- 13.1 kernel
// evaluations/clk/SM = 26 pairs/clk/SM at N = 10^6 (4 rows x 2 CTAs/SM: 12.5; with 1/16 of the
// exponentials on the FMA pipe: 12.0 -- here the FMA pipe and the issue slots are the scarce resource).
using Sym2 = SymCfg<2, 0, 1, 16, 512, 8, 4>;
using Sym3 = SymCfg<3, 0, 1, 16, 512, 8, 4>;

}  // namespace

bool sym_supported(int D) { return D >= 1 && D <= 3; }
int sym_padded_dim(int D) { return D <= 2 ? 2 : 3; }
int sym_tile_rows() { return Sym3::TILE_ROWS; }
int sym_block_sources() { return Sym3::SB; }
long long sym_total_units(long long n_tiles, long long nsb) { return sym_prefix<Sym3::TB>(n_tiles, nsb); }

int sym_grid(int D, int sms, int* grid) {
    return sym_padded_dim(D) == 2 ? sym_grid_of<Sym2>(sms, grid) : sym_grid_of<Sym3>(sms, grid);
}

// the main kernel over P.unit_begin .. P.unit_end on P.grid CTAs, then the combine kernel
int sym_launch(int D, const SymParams& P, cudaStream_t stream) {
    return sym_padded_dim(D) == 2 ? sym_launch_of<Sym2>(P, stream) : sym_launch_of<Sym3>(P, stream);
}

}  // namespace kmb
