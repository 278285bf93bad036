TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 2 --master-port 29611 tools/run_cg_distributed.py 1000000 1.0 symmetric nystrom 2> gpurun_out/mg_err1.log | tail -1 | tee gpurun_out/cg_pcg_sym_2gpu.json
$TR --nproc-per-node 2 --master-port 29612 tools/run_cg_distributed.py 1000000 1.0 rows nystrom 2> gpurun_out/mg_err2.log | tail -1 | tee gpurun_out/cg_pcg_rows_2gpu.json
tail -n 3 gpurun_out/mg_err1.log; tail -n 3 gpurun_out/mg_err2.log
