#!/bin/bash
# round 2, session k: where the C4 block time goes -- timing-only builds of kprod_tensor_pv16 with parts removed (KMB_PV16_X bits:
# 1 no log2 k phase, 2 no MUFU.EX2, 4 no hi/lo split, 8 no P.B MMAs, 16 no S MMAs); results of those builds are wrong by design.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
P=kernel_matrix_benchmarks_b200
for round in 1 2; do
for name in ${VARIANTS:-default x3 x7 x8 x16 x24 x31}; do
  lib=$PWD/$P/libkmb_b200_$name.so; [ $name = default ] && lib=$PWD/$P/libkmb_b200.so
  [ -f $lib ] || continue
  KMB_B200_LIB=$lib timeout 300 python tools/bench_configs.py c4 c4g 2>>$O/r2_k.err | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(json.dumps({'variant':'$name','round':$round,'config':d['config'],'ms':d['ms'],'kernel_ms':d['kernel_ms'],'rel_l2':d['parity']['rel_l2']}))" | tee -a $O/r2_k_parts.jsonl
done
done
tail -5 $O/r2_k.err
