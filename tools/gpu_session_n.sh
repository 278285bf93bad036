#!/bin/bash
# round 2, session n: ring counters instead of run-time modulo, explicit shared-memory line loads; A/B, timing, one ncu capture
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
P=kernel_matrix_benchmarks_b200
timeout 600 python -m pytest tests/test_product_gpu.py -m gpu -x -q -k "tensor or c4 or wide or pv or attention" > $O/r2_n_gputests.log 2>&1; echo "gputests rc=$?"; tail -2 $O/r2_n_gputests.log
for round in 1 2; do
for v in default:128 old:128 default:64; do
  name=${v%%:*}; fl=${v##*:}
  lib=$PWD/$P/libkmb_b200_$name.so; [ $name = default ] && lib=$PWD/$P/libkmb_b200.so
  KMB_PV16_FLUSH_BLOCKS=$fl KMB_B200_LIB=$lib timeout 300 python tools/bench_configs.py c4 c4g 2>>$O/r2_n.err | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(json.dumps({'variant':'$name','flush':$fl,'round':$round,'config':d['config'],'ms':d['ms'],'kernel_ms':d['kernel_ms'],'rel_l2':d['parity']['rel_l2']}))" | tee -a $O/r2_n_ab.jsonl
done
done
for name in t0 t7; do
  echo "== $name" | tee -a $O/r2_n_timing.txt
  KMB_B200_LIB=$PWD/$P/libkmb_b200_$name.so timeout 200 python tools/pv16_timing.py 65536 2>&1 | tail -2 | tee -a $O/r2_n_timing.txt
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pv16_pair -s 1 -c 1 -f -o $O/r2_pv16_ng2 python tools/pv16_run.py 65536 > $O/r2_n_ncu.log 2>&1; echo "ncu rc=$?"
tail -3 $O/r2_n.err
