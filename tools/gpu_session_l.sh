#!/bin/bash
# round 2, session l: per-phase clock64 timing (KMB_PV16_TIMING) of the NG = 2 / SST = 3 kernel and of builds with parts removed
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
P=kernel_matrix_benchmarks_b200
for name in t0 t7 t24 t31; do
  echo "== $name" | tee -a $O/r2_l_timing.txt
  KMB_B200_LIB=$PWD/$P/libkmb_b200_$name.so timeout 200 python tools/pv16_timing.py 65536 2>&1 | tail -2 | tee -a $O/r2_l_timing.txt
done
