#!/bin/bash
# On an 8-GPU box (gpurun --gpus 8): the row-sharded tensor-path configs (C3, C4) and the preconditioned C5 solve in
# both matvec modes at 8 GPUs.  Output lines go to gpurun_out/.
mkdir -p gpurun_out
G=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $G"
$TR --master-port 29631 tools/bench_configs.py c3 c4 2> gpurun_out/mg8_cfg_err.log | grep '^{' | tee gpurun_out/configs_tensor_${G}gpu.jsonl
$TR --master-port 29632 tools/run_cg_distributed.py 1000000 1.0 symmetric nystrom 2> gpurun_out/mg8_cg1_err.log | tail -n 1 | tee gpurun_out/cg_pcg_sym_${G}gpu.json
$TR --master-port 29633 tools/run_cg_distributed.py 1000000 1.0 rows nystrom 2> gpurun_out/mg8_cg2_err.log | tail -n 1 | tee gpurun_out/cg_pcg_rows_${G}gpu.json
tail -n 2 gpurun_out/mg8_cfg_err.log gpurun_out/mg8_cg1_err.log gpurun_out/mg8_cg2_err.log 2>/dev/null | tail -n 12
