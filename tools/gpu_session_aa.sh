#!/bin/bash
# round 2, session aa: |u|^2 + |v|^2 through one more (BF16) K step of the S contraction: default = exponential kernel only,
# x2 = both kernels, x0 = neither; tests with each library first
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
P=kernel_matrix_benchmarks_b200
for name in default x2; do
  lib=$PWD/$P/libkmb_b200_$name.so; [ $name = default ] && lib=$PWD/$P/libkmb_b200.so
  KMB_B200_LIB=$lib timeout 600 python -m pytest tests/test_product_gpu.py -m gpu -q -k "tensor or c4 or wide or pv or attention" > $O/r2_aa_gputests_$name.log 2>&1; echo "gputests $name rc=$?"; tail -4 $O/r2_aa_gputests_$name.log | cut -c1-200
done
for round in 1 2; do
for name in ${VARIANTS:-default x0 x2}; do
  lib=$PWD/$P/libkmb_b200_$name.so; [ $name = default ] && lib=$PWD/$P/libkmb_b200.so
  KMB_B200_LIB=$lib timeout 300 python tools/bench_configs.py c4 c4g 2>>$O/r2_aa.err | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(json.dumps({'variant':'$name','round':$round,'config':d['config'],'ms':d['ms'],'kernel_ms':d['kernel_ms'],'rel_l2':d['parity']['rel_l2']}))" | tee -a $O/r2_aa_ab.jsonl
done
done
tail -3 $O/r2_aa.err
