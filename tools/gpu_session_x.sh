#!/bin/bash
# round 2, session x: final library -- full GPU test tier, smoke, default bench line, reference arm, ncu captures of the C4 kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2_x_gputests.log 2>&1; echo "gputests rc=$?"; tail -3 $O/r2_x_gputests.log
timeout 300 python __graft_entry__.py smoke > $O/r2_x_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 $O/r2_x_smoke.log | cut -c1-200
timeout 500 python bench.py > $O/r2_bench_x.json 2> $O/r2_bench_x.err; echo "bench rc=$?"
python - <<PY
import json
d = [json.loads(l) for l in open("$O/r2_bench_x.json") if l.startswith("{")][-1]
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "frac", d["roofline"]["frac"], "parity", d["parity"]["rel_l2"])
for k, v in d["configs"].items():
    print(k, "ms", v.get("ms"), "kernel_ms", v.get("kernel_ms"), "fit", v.get("fit_ms"), "it", v.get("iterations"), "parity", (v.get("parity") or {}).get("rel_l2"), (v.get("parity") or {}).get("rel_residual_oracle"), "frac", (v.get("roofline") or {}).get("frac"))
print(d.get("clocks"))
PY
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 | cut -c1-300
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pv16_pair -s 1 -c 1 -f -o $O/r2_pv16_final_exp python tools/pv16_run.py 65536 absolute-exponential > $O/r2_x_ncu1.log 2>&1; echo "ncu exp rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pv16_pair -s 1 -c 1 -f -o $O/r2_pv16_final_gauss python tools/pv16_run.py 65536 gaussian > $O/r2_x_ncu2.log 2>&1; echo "ncu gauss rc=$?"
