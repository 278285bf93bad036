#!/bin/bash
# two-set epilogue of kprod_tensor_pv16 (build-time variant): parity tests with the variant library, A/B at full-size C4, phase timing,
# ncu source-level capture of the production pair kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
P=$PWD/kernel_matrix_benchmarks_b200
KMB_B200_LIB=$P/libkmb_b200_sets.so timeout 600 python -m pytest tests/test_product_gpu.py -m gpu -q -x -k "golden or c4 or wide or attention or degenerate or tensor" > $O/r2_gputest_sets.log 2>&1; echo "sets pytest rc=$?"; tail -5 $O/r2_gputest_sets.log
VARIANTS="default sets r1pv16" bash tools/pv16_ab.sh 2>&1 | grep -v "^=="
for v in timing timing_sets; do echo "== $v"; KMB_B200_LIB=$P/libkmb_b200_$v.so timeout 120 python tools/pv16_timing.py 32768 2>&1 | tail -2 | tee -a $O/r2_pv16_phase_timing_sets.txt; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pv16_pair -s 1 -c 1 -o $O/r2_pv16_pair_n64k python tools/pv16_run.py 65536 > $O/ncu_e.log 2>&1; tail -2 $O/ncu_e.log
KMB_B200_LIB=$P/libkmb_b200_sets.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:pv16_pair -s 1 -c 1 -o $O/r2_pv16_pair_sets_n64k python tools/pv16_run.py 65536 > $O/ncu_f.log 2>&1; tail -2 $O/ncu_f.log
ls -la $O/*.ncu-rep
