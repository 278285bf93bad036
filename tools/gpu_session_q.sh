#!/bin/bash
# round 2, session q: phase timing of the 8-warp build with the column groups skewed
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
P=kernel_matrix_benchmarks_b200
for skew in 0 600 -600 3000; do
  echo "== t0w1 skew $skew" | tee -a $O/r2_q_timing.txt
  KMB_PV16_SKEW_NS=$skew KMB_B200_LIB=$PWD/$P/libkmb_b200_t0w1.so timeout 200 python tools/pv16_timing.py 65536 2>&1 | tail -2 | tee -a $O/r2_q_timing.txt
done
