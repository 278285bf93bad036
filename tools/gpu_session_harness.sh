#!/bin/bash
# The BASELINE-size datasets (C2, C3, C4, C5) through the reference's own harness on one B200:
# dataset files with float64 GPU ground truth (spot-verified by the reference's GroundTruth), run.py --local --hardware GPU
# with this repo's algos.yaml, scores by the reference's plotting.metrics.  Appends to gpurun_out/r2_harness_scores.jsonl.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
NAMES=${NAMES:-"product-ucube-D3-E1-M1000000-N1000000-gaussian solver-ucubelam1-D3-E1-M1000000-N1000000-gaussian attention-ucube-D64-E64-M262144-N262144-absolute-exponential product-ucube-D784-E1-M60000-N10000-gaussian"}
for name in $NAMES; do
  echo "=== $name"
  ( time timeout 1200 python tools/run_harness.py --prepare $name ) > $O/r2_harness_prepare_$name.log 2>&1 || { echo "prepare failed"; tail -5 $O/r2_harness_prepare_$name.log; continue; }
  grep "wrote\|exists" $O/r2_harness_prepare_$name.log | cut -c1-400
  ( time timeout 1200 python tools/run_harness.py --dataset $name --hardware GPU --runs 2 ) > $O/r2_harness_run_$name.log 2>&1 || { echo "run failed"; tail -8 $O/r2_harness_run_$name.log; }
  timeout 900 python tools/run_harness.py --score $name --json $O/r2_harness_scores.jsonl > $O/r2_harness_score_$name.log 2>&1 || { echo "score failed"; tail -8 $O/r2_harness_score_$name.log; }
  python - <<PY
import json
for l in open("$O/r2_harness_score_$name.log"):
    try: d = json.loads(l)
    except Exception: continue
    print({k: d.get(k) for k in ("name", "total-time", "query-time", "build-time", "rel-l2-error", "rel-residual", "gpairs_per_s", "n_gpus", "cg_iterations")})
PY
done
