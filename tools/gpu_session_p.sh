#!/bin/bash
# round 2, session p: is the epilogue's MUFU phase shared (lock step) or latency bound per warp?  8-warp build, group 1 idle (X = 32)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
P=kernel_matrix_benchmarks_b200
for name in t0w1 t32; do
  echo "== $name" | tee -a $O/r2_p_timing.txt
  KMB_B200_LIB=$PWD/$P/libkmb_b200_$name.so timeout 200 python tools/pv16_timing.py 65536 2>&1 | tail -2 | tee -a $O/r2_p_timing.txt
done
