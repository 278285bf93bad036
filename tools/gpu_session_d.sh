#!/bin/bash
# one gpurun call: tests, pv16 flush tuning, wide-signal tensor route, ncu of the symmetric kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/r2_gputest_d.log 2>&1; echo "pytest rc=$?"; tail -12 $O/r2_gputest_d.log
for f in 100000 128 64 32; do
  echo "flush=$f"; KMB_PV16_FLUSH_BLOCKS=$f timeout 300 python tools/bench_configs.py c4 c4g 2>>$O/r2_pv16_flush.err | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(json.dumps({'flush_blocks':$f,'config':d['config'],'ms':d['ms'],'kernel_ms':d['kernel_ms'],'rel_l2':d['parity']['rel_l2']}))" | tee -a $O/r2_pv16_flush.jsonl
done
python - <<'PY' 2>&1 | tee $O/r2_wide_tensor.jsonl
import json, numpy as np, torch, sys
sys.path.insert(0, '.')
from kernel_matrix_benchmarks_b200 import product
from oracle import c_oracle
rng = np.random.RandomState(3)
n = 131072
for D in (3, 16):
    y, x = rng.rand(n, D) * (3.0 / D) ** 0.5, rng.rand(n, D) * (3.0 / D) ** 0.5
    rows = np.sort(rng.choice(n, 64, replace=False))
    for E in (16, 64):
        b = rng.randn(n, E)
        ty, tx, tb = (torch.tensor(a, dtype=torch.float32, device="cuda") for a in (y, x, b))
        for kernel in ("gaussian", "absolute-exponential"):
            for norm in (False, True):
                for path in ("direct", "tensor_f16"):
                    try:
                        out = product.kernel_product(tx, ty, tb, kernel=kernel, normalize_rows=norm, path=path)
                        torch.cuda.synchronize()
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record()
                        for _ in range(3):
                            product.kernel_product(tx, ty, tb, kernel=kernel, normalize_rows=norm, path=path, out=out)
                        e1.record(); torch.cuda.synchronize()
                        ms = e0.elapsed_time(e1) / 3
                        want = c_oracle.kernel_product(kernel, y, x[rows], b, normalize_rows=norm)
                        got = out[torch.as_tensor(rows, device="cuda")].cpu().numpy().astype(np.float64)
                        print(json.dumps({"D": D, "E": E, "kernel": kernel, "norm": norm, "path": path, "ms": ms, "gpairs_per_s": n * n / ms / 1e6,
                                          "rel_l2": float(np.linalg.norm(got - want) / np.linalg.norm(want))}))
                    except Exception as ex:
                        print(json.dumps({"D": D, "E": E, "kernel": kernel, "norm": norm, "path": path, "error": str(ex)[:200]}))
PY
# launch list + full capture of the symmetric kernel (after the plain run above exited 0)
timeout 300 python bench.py --steps 2 --warmup 1 --no-configs --no-cpu-baseline > $O/r2_bench_short.json 2> $O/r2_bench_short.err && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r2_launches_bench_sym_n1m.csv python bench.py --steps 2 --warmup 1 --no-configs --no-cpu-baseline > $O/ncu_a.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kprod_sym_kernel -c 1 -o $O/r2_sym_kernel_n1m python bench.py --steps 1 --warmup 1 --no-configs --no-cpu-baseline --no-e2e > $O/ncu_b.log 2>&1
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"kprod_sym_kernel|sym_combine" -c 4 --csv --log-file $O/r2_sym_traffic.csv python bench.py --steps 1 --warmup 1 --no-configs --no-cpu-baseline --no-e2e > $O/ncu_c.log 2>&1
tail -6 $O/r2_sym_traffic.csv
ls -la $O | tail -12
