#!/bin/bash
# round 2, session ah: end-to-end arm with the pooled pinned float64 result buffers (host ms per plugin call), 1 GPU
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 200 python bench.py --no-configs --no-cpu-baseline --steps 10 --warmup 3 2> $O/r2_ah_1.err | grep "^{" | tail -1 > $O/r2_ah_bench_1gpu.json
python -c "
import json; d=json.load(open('$O/r2_ah_bench_1gpu.json')); print('1 GPU value', round(d['value']), 'e2e', round(d['e2e']['value']), d['e2e']['ms_per_step'], d['e2e']['host_ms_per_call'])"
timeout 300 python -m pytest tests/test_product_gpu.py tests/test_solver_gpu.py tests/test_harness_gpu.py -m gpu -x -q 2>&1 | tail -2
