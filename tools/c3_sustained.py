#!/usr/bin/env python
"""BASELINE config 3 back to back for a few seconds with the clocks sampled beside it: is the tensor kernel paced by
the power cap (as the driver's cuBLAS reference is: MEASURED_PEAKS.json bf16_tflops 1667 burst / 1405 sustained)?

    python tools/c3_sustained.py [seconds]      (KMB_TENSOR_PAIR=0: single-CTA kernel)
"""
import json, os, subprocess, sys, time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kernel_matrix_benchmarks_b200 import datasets, product  # noqa: E402
from kernel_matrix_benchmarks_b200.algorithms.b200 import B200Product  # noqa: E402

seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 3.0
ds = datasets.config_c3()
algo = B200Product(kernel=ds.kernel, dimension=ds.D, precision="float32")
algo.prepare_data(source_points=ds.source_points, target_points=ds.target_points)
algo.fit()
algo.prepare_query(source_signal=ds.source_signal)
product.set_profiling(True)
for _ in range(3):
    algo.query()
smi = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown",
                        "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
time.sleep(0.3)
t0, main, total = time.perf_counter(), [], []
while time.perf_counter() - t0 < seconds:
    algo.query()
    main.append(product.last_main_kernel_ms())
    total.append(algo.get_additional()["gpu_query_ms"])
smi.terminate()
rows = [r.split(", ") for r in smi.stdout.read().strip().splitlines()][3:]
clk = sorted(float(r[0]) for r in rows if float(r[1]) > 400) or [0.0]
flops = 3 * 2.0 * ds.N * ds.M * ds.D
half = main[len(main) // 2:]
print(json.dumps({"config": "C3 sustained", "pair_kernel": os.environ.get("KMB_TENSOR_PAIR", "1") != "0", "queries": len(main),
                  "main_kernel_ms_first10": float(np.mean(main[:10])), "main_kernel_ms_second_half": float(np.mean(half)),
                  "query_ms_second_half": float(np.mean(total[len(total) // 2:])),
                  "executed_f16_tflops_second_half": flops / (np.mean(half) * 1e-3) / 1e12,
                  "sm_mhz_median_under_load": clk[len(clk) // 2], "power_w_max": max(float(r[1]) for r in rows),
                  "sw_power_cap_active_samples": sum(r[2].strip() == "Active" for r in rows), "samples": len(rows)}))
algo.done()
