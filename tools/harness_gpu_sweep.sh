#!/bin/bash
# Run on the GPU box (gpurun): the reference harness drives the B200 plugin on every prepared dataset,
# then the reference's own CPU brute force on the C1-sized ones, then everything is scored with the
# reference's metrics.  Output: gpurun_out/harness_scores.jsonl + one log per run.
DS="product-ucube-D3-E1-M10000-N10000-gaussian product-ucube-D3-E1-M10000-N10000-absolute-exponential product-ucube-D3-E1-M10000-N10000-inverse-distance product-ucube-D784-E1-M4000-N1000-gaussian attention-ucube-D64-E64-M4096-N4096-absolute-exponential attention-ucube-D64-E64-M4096-N4096-gaussian solver-ucubelam1-D3-E1-M2000-N2000-gaussian solver-ucubelam1-D3-E1-M10000-N10000-gaussian"
mkdir -p gpurun_out
for d in $DS; do
  python tools/run_harness.py --dataset $d --hardware GPU > gpurun_out/harness_$d.log 2>&1 || tail -20 gpurun_out/harness_$d.log
done
for d in product-ucube-D3-E1-M10000-N10000-gaussian product-ucube-D784-E1-M4000-N1000-gaussian attention-ucube-D64-E64-M4096-N4096-gaussian; do
  python tools/run_harness.py --dataset $d --hardware CPU --algorithm bruteforce-product-blas > gpurun_out/harness_cpu_$d.log 2>&1 || tail -20 gpurun_out/harness_cpu_$d.log
done
python tools/run_harness.py --dataset solver-ucubelam1-D3-E1-M2000-N2000-gaussian --hardware CPU --algorithm bruteforce-solver-blas > gpurun_out/harness_cpu_solver.log 2>&1
rm -f gpurun_out/harness_scores.jsonl
for d in $DS; do
  python tools/run_harness.py --score $d --json gpurun_out/harness_scores.jsonl > /dev/null 2>&1
done
python - <<'PY'
import json
for l in open('gpurun_out/harness_scores.jsonl'):
    r = json.loads(l)
    print(r['dataset'][:48], '|', r['name'][:50], '| total %.4g s | rel-l2 %.2e | max-err %.2e' % (r['total-time'], r['rel-l2-error'], r['max-error']))
PY
