#!/bin/bash
# round 2, session t: one fused pass per block (default) against two phases (nf) and the NG = 4 / SST = 2 kernel (old); tests first
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
P=kernel_matrix_benchmarks_b200
timeout 600 python -m pytest tests/test_product_gpu.py -m gpu -x -q > $O/r2_t_gputests.log 2>&1; echo "gputests rc=$?"; tail -2 $O/r2_t_gputests.log
for round in 1 2; do
for name in ${VARIANTS:-default nf old}; do
  lib=$PWD/$P/libkmb_b200_$name.so; [ $name = default ] && lib=$PWD/$P/libkmb_b200.so
  KMB_B200_LIB=$lib timeout 300 python tools/bench_configs.py c4 c4g 2>>$O/r2_t.err | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(json.dumps({'variant':'$name','round':$round,'config':d['config'],'ms':d['ms'],'kernel_ms':d['kernel_ms'],'rel_l2':d['parity']['rel_l2']}))" | tee -a $O/r2_t_ab.jsonl
done
done
for name in tf tnf; do
  echo "== $name" | tee -a $O/r2_t_timing.txt
  KMB_B200_LIB=$PWD/$P/libkmb_b200_$name.so timeout 200 python tools/pv16_timing.py 65536 2>&1 | tail -2 | tee -a $O/r2_t_timing.txt
done
tail -3 $O/r2_t.err
