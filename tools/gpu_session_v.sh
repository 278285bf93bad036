#!/bin/bash
# round 2, session v (gpurun --gpus 2): where the end-to-end overhead of a step goes (host ms per plugin call), 1 and 2 GPUs
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 200 python bench.py --no-configs --no-cpu-baseline --steps 10 --warmup 3 2> $O/r2_v_1.err | grep "^{" | tail -1 > $O/r2_v_bench_1gpu.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --no-configs --no-cpu-baseline --steps 10 --warmup 3 2> $O/r2_v_2.err | grep "^{" | tail -1 > $O/r2_v_bench_2gpu.json
python - <<PY
import json
for g in (1, 2):
    d = json.load(open(f"$O/r2_v_bench_{g}gpu.json"))
    print(g, "GPU value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"], 3), d["e2e"]["host_ms_per_call"])
PY
