#!/bin/bash
# round 2, session j: the C4 kernel with NG = 2 column groups / SST = 3 S stages (default build) against the
# NG = 4 / SST = 2 build (libkmb_b200_old.so) on ONE box, skew sweep; pv16 GPU tests with the default build first.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
P=kernel_matrix_benchmarks_b200
timeout 600 python -m pytest tests/test_product_gpu.py -m gpu -x -q > $O/r2_j_gputests.log 2>&1; echo "gputests rc=$?"; tail -3 $O/r2_j_gputests.log
for round in 1 2; do
for v in ${VARIANTS:-default:0 old:0 default:300 default:600 default:1000 default:1500}; do
  name=${v%%:*}; skew=${v##*:}
  lib=$PWD/$P/libkmb_b200_$name.so; [ $name = default ] && lib=$PWD/$P/libkmb_b200.so
  [ -f $lib ] || continue
  KMB_PV16_SKEW_NS=$skew KMB_B200_LIB=$lib timeout 300 python tools/bench_configs.py c4 c4g 2>>$O/r2_j_ab.err | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(json.dumps({'variant':'$name','skew_ns':$skew,'round':$round,'config':d['config'],'ms':d['ms'],'kernel_ms':d['kernel_ms'],'rel_l2':d['parity']['rel_l2']}))" | tee -a $O/r2_j_ab.jsonl
done
done
tail -5 $O/r2_j_ab.err
