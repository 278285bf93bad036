// Instantiations of kprod_direct_kernel: kernel invdist, normalize_rows=0, difference form
// (split per file to build in parallel).
#include "kprod_direct.cuh"
KMB_DIRECT_TABLE(kDirect_invdist_n0, 2, false, 0)
