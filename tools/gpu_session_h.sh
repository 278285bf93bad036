#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
VARIANTS="default vnldg" bash tools/pv16_ab.sh 2>&1 | grep -v "^=="
timeout 400 python bench.py > $O/r2_bench_h.json 2> $O/r2_bench_h.err; echo "bench rc=$?"
python - <<PY
import json
d = json.load(open("$O/r2_bench_h.json"))
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "frac", d["roofline"]["frac"], "parity", d["parity"]["rel_l2"])
print("general", d["general_kernel"]["value"], d["general_kernel"]["kernel_ms"])
for k, v in d["configs"].items():
    print(k, "ms", v.get("ms"), "kernel_ms", v.get("kernel_ms"), "fit", v.get("fit_ms"), v.get("first_fit_ms"), "it", v.get("iterations"), "e2e", v.get("e2e_ms"), "parity", (v.get("parity") or {}).get("rel_l2"), (v.get("parity") or {}).get("rel_residual_oracle"))
print(d["cpu_baseline"])
PY
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 | cut -c1-400
