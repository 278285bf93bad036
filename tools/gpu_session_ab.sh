#!/bin/bash
# round 2, session ab: exponential kernel, fused pass on top of the BF16 norm step (f2) against two phases (default)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
P=kernel_matrix_benchmarks_b200
for round in 1 2 3; do
for name in ${VARIANTS:-default f2}; do
  lib=$PWD/$P/libkmb_b200_$name.so; [ $name = default ] && lib=$PWD/$P/libkmb_b200.so
  KMB_B200_LIB=$lib timeout 300 python tools/bench_configs.py c4 2>>$O/r2_ab.err | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(json.dumps({'variant':'$name','round':$round,'config':d['config'],'ms':d['ms'],'kernel_ms':d['kernel_ms'],'rel_l2':d['parity']['rel_l2']}))" | tee -a $O/r2_ab_ab.jsonl
done
done
tail -3 $O/r2_ab.err
