// Instantiations of kprod_direct_kernel: Gaussian product form, normalize_rows=1.
#include "kprod_direct.cuh"
KMB_DIRECT_TABLE(kDirect_gaussprod_n1, 0, true, 1)
