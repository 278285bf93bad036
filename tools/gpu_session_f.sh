#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -x > $O/r2_gputest_f.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r2_gputest_f.log
VARIANTS="default fused r1pv16" bash tools/pv16_ab.sh 2>&1 | grep -v "^==" 
KMB_B200_LIB=$PWD/kernel_matrix_benchmarks_b200/libkmb_b200_timing.so timeout 120 python tools/pv16_timing.py 32768 2>&1 | tail -2 | tee -a $O/r2_pv16_phase_timing_countdown.txt
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"kprod_sym_kernel|sym_combine" -c 2 --csv --log-file $O/r2_sym_traffic_stcs.csv python bench.py --steps 1 --warmup 1 --no-configs --no-cpu-baseline --no-e2e > $O/ncu_d.log 2>&1
grep "kprod_sym_kernel\|sym_combine" $O/r2_sym_traffic_stcs.csv | awk -F'","' '{print $5, $(NF-2), $NF}' | cut -c1-200
bash tools/gpu_session_sanitizer.sh
bash tools/gpu_session_harness.sh
