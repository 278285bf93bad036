// Instantiations of kprod_direct_kernel: Gaussian product form, normalize_rows=0.
#include "kprod_direct.cuh"
KMB_DIRECT_TABLE(kDirect_gaussprod_n0, 0, false, 1)
