#!/bin/bash
# round 2, session s: SM clock inside the full-size C4 kernel (clock64 / globaltimer of CTA 0), power and clocks from nvidia-smi beside it
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
P=kernel_matrix_benchmarks_b200
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,power.limit,clocks_throttle_reasons.active --format=csv -lms 20 > $O/r2_s_smi.csv 2>/dev/null &
SMI=$!
for name in t0w1 told; do
  echo "== $name" | tee -a $O/r2_s_timing.txt
  KMB_B200_LIB=$PWD/$P/libkmb_b200_$name.so timeout 300 python tools/pv16_timing.py 262144 2>&1 | tail -2 | tee -a $O/r2_s_timing.txt
done
sleep 1
echo "== production library, 20 back-to-back C4 queries" | tee -a $O/r2_s_timing.txt
date +%s.%N >> $O/r2_s_timing.txt
KMB_B200_LIB=$PWD/$P/libkmb_b200_w1.so timeout 300 python tools/bench_configs.py c4 2>>$O/r2_s.err | cut -c1-300 | tee -a $O/r2_s_timing.txt
date +%s.%N >> $O/r2_s_timing.txt
kill $SMI
python - <<PY
import csv
rows=[r for r in csv.reader(open("$O/r2_s_smi.csv"))][1:]
busy=[r for r in rows if float(r[2].split()[0])>600]
print("samples", len(rows), "busy(>600W)", len(busy))
import statistics as st
if busy:
    print("sm MHz under load: median", st.median(float(r[0].split()[0]) for r in busy), "min", min(float(r[0].split()[0]) for r in busy), "power max", max(float(r[2].split()[0]) for r in busy), "reasons", set(r[4].strip() for r in busy))
PY
