#!/bin/bash
# round 2, session u (gpurun --gpus 8): 1-GPU bench line on the same box, then tools/gpu_session_multi.sh 8 harness
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 200 python bench.py --no-configs --steps 5 --warmup 3 > $O/r2_bench_1gpu_on_8gpu_box.json 2> $O/r2_bench_1gpu_on_8gpu_box.err; echo "bench 1 rc=$?"
python -c "
import json; d=json.load(open('$O/r2_bench_1gpu_on_8gpu_box.json')); print('1 GPU value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])"
bash tools/gpu_session_multi.sh 8 harness
