// Instantiations of kprod_direct_kernel: kernel gauss, normalize_rows=0, difference form
// (split per file to build in parallel).
#include "kprod_direct.cuh"
KMB_DIRECT_TABLE(kDirect_gauss_n0, 0, false, 0)
