#!/bin/bash
# round 2, session ai: flush interval of the C4 kernel (final library): 128 (default) against 64 blocks
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
for round in 1 2 3; do
for fl in 128 64; do
  KMB_PV16_FLUSH_BLOCKS=$fl timeout 300 python tools/bench_configs.py c4 c4g 2>>$O/r2_ai.err | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(json.dumps({'flush':$fl,'round':$round,'config':d['config'],'ms':d['ms'],'kernel_ms':d['kernel_ms'],'rel_l2':d['parity']['rel_l2']}))" | tee -a $O/r2_ai_ab.jsonl
done
done
