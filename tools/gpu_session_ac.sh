#!/bin/bash
# round 2, session ac (gpurun --gpus 8): final library, headline bench at 1 and 8 GPUs on one box (no configs block), e2e breakdown
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 200 python bench.py --no-configs --no-cpu-baseline --steps 10 --warmup 3 2> $O/r2_ac_1.err | grep "^{" | tail -1 > $O/r2_ac_bench_1gpu.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --no-configs --no-cpu-baseline --steps 10 --warmup 3 2> $O/r2_ac_8.err | grep "^{" | tail -1 > $O/r2_ac_bench_8gpu.json
python - <<PY
import json
r = {}
for g in (1, 8):
    d = json.load(open(f"$O/r2_ac_bench_{g}gpu.json")); r[g] = d
    print(g, "GPU value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"], 3), d["e2e"]["host_ms_per_call"], "parity", d["parity"]["rel_l2"])
print("efficiency at 8: value", r[8]["value"] / r[1]["value"] / 8, "e2e", r[8]["e2e"]["value"] / r[1]["e2e"]["value"] / 8)
PY
