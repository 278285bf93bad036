#!/bin/bash
# A/B of build-time variants of kprod_tensor_pv16 on ONE box (clocks under the 1000 W cap differ from box to box):
# full-size C4 (exponential + Gaussian kernel) through tools/bench_configs.py with each library, twice, interleaved.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
P=kernel_matrix_benchmarks_b200
for round in 1 2; do
for v in ${VARIANTS:-default fused skew150 skew300 skew600}; do
  lib=$PWD/$P/libkmb_b200_$v.so; [ $v = default ] && lib=$PWD/$P/libkmb_b200.so
  [ -f $lib ] || continue
  KMB_B200_LIB=$lib timeout 300 python tools/bench_configs.py c4 c4g 2>>$O/r2_pv16_ab.err | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(json.dumps({'variant':'$v','round':$round,'config':d['config'],'ms':d['ms'],'kernel_ms':d['kernel_ms'],'rel_l2':d['parity']['rel_l2']}))" | tee -a $O/r2_pv16_ab.jsonl
done
done
for v in timing timing_skew300; do
  echo "== $v"; KMB_B200_LIB=$PWD/$P/libkmb_b200_$v.so timeout 120 python tools/pv16_timing.py 32768 2>&1 | tail -3 | tee -a $O/r2_pv16_phase_timing.txt
done
