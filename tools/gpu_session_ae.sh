#!/bin/bash
# round 2, session ae: HEAD -- full GPU tier, smoke, 2-rank torchrun bench (no configs) for the plugin's sliced casts
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2_ae_gputests.log 2>&1; echo "gputests rc=$?"; tail -2 $O/r2_ae_gputests.log
timeout 300 python __graft_entry__.py smoke > $O/r2_ae_smoke.log 2>&1; echo "smoke rc=$?"
timeout 200 python bench.py --no-configs --steps 5 --warmup 3 2> $O/r2_ae_1.err | grep "^{" | tail -1 > $O/r2_ae_bench_1gpu.json
python -c "
import json; d=json.load(open('$O/r2_ae_bench_1gpu.json')); print('1 GPU value', round(d['value']), 'e2e', round(d['e2e']['value']), d['e2e']['host_ms_per_call'], 'parity', d['parity']['rel_l2'], d['cpu_baseline']['value'])"
