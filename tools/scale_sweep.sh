#!/bin/bash
# Run on an 8-GPU box (gpurun --gpus 8): the headline product at 8/4/2 GPUs (strong scaling: symmetric unit list
# split across ranks + one all-reduce), the general kernel with sharded rows at 8 GPUs, and the C5 solve
# (N = M = 1M, lambda = 1) at 8 GPUs in both matvec modes.  Output lines go to gpurun_out/scale_*.json.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for g in 8 4 2; do
  $TR --nproc-per-node $g --master-port $((29500+g)) bench.py --gpus $g --steps 5 --warmup 3 2> gpurun_out/scale_bench_${g}gpu.err | tail -1 > gpurun_out/scale_bench_${g}gpu.json
  python -c "import json;r=json.load(open('gpurun_out/scale_bench_${g}gpu.json'));print(r['n_gpus'],'GPUs',round(r['value']),'Gpairs/s e2e',round(r['e2e']['value']),'frac',round(r['roofline']['frac'],3),'general',round(r['general_kernel']['value']))"
done
$TR --nproc-per-node 8 --master-port 29520 bench.py --gpus 8 --steps 5 --warmup 3 --path direct 2> gpurun_out/scale_bench_8gpu_direct.err | tail -1 > gpurun_out/scale_bench_8gpu_direct.json
python -c "import json;r=json.load(open('gpurun_out/scale_bench_8gpu_direct.json'));print('general kernel, 8 GPUs',round(r['value']),'Gpairs/s e2e',round(r['e2e']['value']),'frac',round(r['roofline']['frac'],3))"
$TR --nproc-per-node 8 --master-port 29600 tools/run_cg_distributed.py 1000000 1.0 symmetric 2> gpurun_out/scale_cg_8gpu.err | tail -1 | tee gpurun_out/scale_cg_sym_8gpu.json
$TR --nproc-per-node 8 --master-port 29601 tools/run_cg_distributed.py 1000000 1.0 rows 2>> gpurun_out/scale_cg_8gpu.err | tail -1 | tee gpurun_out/scale_cg_rows_8gpu.json
