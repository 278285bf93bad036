#!/bin/bash
# round 2, session ag: S of the next block loaded before the wait::st / arrive of this one (default) against no prefetch (pf0)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
P=kernel_matrix_benchmarks_b200
timeout 600 python -m pytest tests/test_product_gpu.py -m gpu -x -q -k "tensor or c4 or wide or pv or attention" > $O/r2_ag_gputests.log 2>&1; echo "gputests rc=$?"; tail -2 $O/r2_ag_gputests.log
for round in 1 2 3; do
for name in ${VARIANTS:-default pf0}; do
  lib=$PWD/$P/libkmb_b200_$name.so; [ $name = default ] && lib=$PWD/$P/libkmb_b200.so
  KMB_B200_LIB=$lib timeout 300 python tools/bench_configs.py c4 c4g 2>>$O/r2_ag.err | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(json.dumps({'variant':'$name','round':$round,'config':d['config'],'ms':d['ms'],'kernel_ms':d['kernel_ms'],'rel_l2':d['parity']['rel_l2']}))" | tee -a $O/r2_ag_ab.jsonl
done
done
tail -3 $O/r2_ag.err
