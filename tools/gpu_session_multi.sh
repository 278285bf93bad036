#!/bin/bash
# usage (inside gpurun --gpus G): bash tools/gpu_session_multi.sh G [harness]
# bench.py under torchrun on G GPUs, the multi-GPU tests, and (with "harness") the n_gpus sweep of algos.yaml through the
# reference's own run.py on the C2 / C5 datasets.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
G=${1:-2}
O=gpurun_out
mkdir -p $O
nvidia-smi -L | head -8
timeout 600 python -m pytest tests/test_multigpu_gpu.py -m gpu -q > $O/r2_gputest_multi_${G}gpu.log 2>&1; echo "multigpu pytest rc=$?"; tail -4 $O/r2_gputest_multi_${G}gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $G --steps 5 --warmup 3 > $O/r2_bench_${G}gpu.json 2> $O/r2_bench_${G}gpu.err; echo "bench rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $G --workload solve > $O/r2_bench_solve_${G}gpu.json 2> $O/r2_bench_solve_${G}gpu.err; echo "solve rc=$?"; cut -c1-600 $O/r2_bench_solve_${G}gpu.json
python - <<PY
import json
try:
    d = json.load(open("$O/r2_bench_${G}gpu.json"))
    print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "parity", d["parity"]["rel_l2"], d["e2e"].get("parity", {}).get("rel_l2"))
    for k, v in d["configs"].items():
        print(k, "ms", v.get("ms"), "kernel_ms", v.get("kernel_ms"), "fit", v.get("fit_ms"), v.get("first_fit_ms"), "it", v.get("iterations"), "parity", v.get("parity"))
except Exception as e:
    print("bench line unreadable:", e)
    print(open("$O/r2_bench_${G}gpu.err").read()[-3000:])
PY
if [ "$2" = "harness" ]; then
  NAMES="product-ucube-D3-E1-M1000000-N1000000-gaussian solver-ucubelam1-D3-E1-M1000000-N1000000-gaussian" bash tools/gpu_session_harness.sh
fi
