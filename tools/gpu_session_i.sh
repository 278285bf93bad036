#!/bin/bash
# round 2, session i: after a container restore — full GPU test tier, smoke, the default bench line and the reference arm
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2_i_gputests.log 2>&1; echo "gputests rc=$?"; tail -3 $O/r2_i_gputests.log
timeout 300 python __graft_entry__.py smoke > $O/r2_i_smoke.log 2>&1; echo "smoke rc=$?"; tail -5 $O/r2_i_smoke.log | cut -c1-300
timeout 500 python bench.py > $O/r2_bench_i.json 2> $O/r2_bench_i.err; echo "bench rc=$?"
python - <<PY
import json
d = json.load(open("$O/r2_bench_i.json"))
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "frac", d["roofline"]["frac"], "parity", d["parity"]["rel_l2"])
print("general", d["general_kernel"]["value"], d["general_kernel"]["kernel_ms"])
for k, v in d["configs"].items():
    print(k, "ms", v.get("ms"), "kernel_ms", v.get("kernel_ms"), "fit", v.get("fit_ms"), v.get("first_fit_ms"), "it", v.get("iterations"), "e2e", v.get("e2e_ms"), "parity", (v.get("parity") or {}).get("rel_l2"), (v.get("parity") or {}).get("rel_residual_oracle"))
print(d["cpu_baseline"]); print(d.get("clocks"))
PY
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 | cut -c1-400
