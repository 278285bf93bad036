#!/bin/bash
# round 2, session w: |v|^2 line written a block ahead (off the critical path), P stored in two halves under the exponentials;
# against the previous commit's library (prev); tests first
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
P=kernel_matrix_benchmarks_b200
timeout 600 python -m pytest tests/test_product_gpu.py -m gpu -x -q > $O/r2_w_gputests.log 2>&1; echo "gputests rc=$?"; tail -2 $O/r2_w_gputests.log
for round in 1 2; do
for name in ${VARIANTS:-default prev}; do
  lib=$PWD/$P/libkmb_b200_$name.so; [ $name = default ] && lib=$PWD/$P/libkmb_b200.so
  KMB_B200_LIB=$lib timeout 300 python tools/bench_configs.py c4 c4g 2>>$O/r2_w.err | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(json.dumps({'variant':'$name','round':$round,'config':d['config'],'ms':d['ms'],'kernel_ms':d['kernel_ms'],'rel_l2':d['parity']['rel_l2']}))" | tee -a $O/r2_w_ab.jsonl
done
done
for name in tf; do
  echo "== $name" | tee -a $O/r2_w_timing.txt
  KMB_B200_LIB=$PWD/$P/libkmb_b200_$name.so timeout 200 python tools/pv16_timing.py 65536 2>&1 | tail -2 | tee -a $O/r2_w_timing.txt
done
tail -3 $O/r2_w.err
