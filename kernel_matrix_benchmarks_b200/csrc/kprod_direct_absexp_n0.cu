// Instantiations of kprod_direct_kernel: kernel absexp, normalize_rows=0, difference form
// (split per file to build in parallel).
#include "kprod_direct.cuh"
KMB_DIRECT_TABLE(kDirect_absexp_n0, 1, false, 0)
