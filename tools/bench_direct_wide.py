#!/usr/bin/env python
"""Direct FP32 path with wide signals (D <= 16, E > 4): Gpairs/s and oracle parity.

    KMB_DIRECT_MAX_EP=4|8|16 python tools/bench_direct_wide.py [--n 131072] [--d 3]

The direct kernel re-evaluates the kernel once per pass of e_chunk signal columns (kmb_api.cu: plan_direct);
KMB_DIRECT_MAX_EP caps e_chunk (4 = the round-1 behaviour).  One JSON line per E."""
import argparse
import json
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=131072)
    ap.add_argument("--d", type=int, default=3)
    ap.add_argument("--es", default="4,8,16,64")
    ap.add_argument("--kernel", default="gaussian")
    args = ap.parse_args()
    import torch

    from kernel_matrix_benchmarks_b200 import product
    from oracle import c_oracle

    rng = np.random.RandomState(3)
    y, x = rng.rand(args.n, args.d), rng.rand(args.n, args.d)
    rows = np.sort(rng.choice(args.n, 64, replace=False))
    for E in [int(e) for e in args.es.split(",")]:
        b = rng.randn(args.n, E)
        ty, tx, tb = (torch.tensor(a, dtype=torch.float32, device="cuda") for a in (y, x, b))
        for norm in (False, True):
            out = product.kernel_product(tx, ty, tb, kernel=args.kernel, normalize_rows=norm, path="direct")
            launches = product.last_launch_count()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                product.kernel_product(tx, ty, tb, kernel=args.kernel, normalize_rows=norm, path="direct", out=out)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            want = c_oracle.kernel_product(args.kernel, y, x[rows], b, normalize_rows=norm)
            got = out[torch.as_tensor(rows, device="cuda")].cpu().numpy().astype(np.float64)
            print(json.dumps({"N": args.n, "D": args.d, "E": E, "normalize_rows": norm, "kernel": args.kernel,
                              "max_ep": os.environ.get("KMB_DIRECT_MAX_EP", "16"), "ms": ms, "launches": launches,
                              "gpairs_per_s": args.n * args.n / ms / 1e6,
                              "rel_l2": float(np.linalg.norm(got - want) / np.linalg.norm(want))}))


if __name__ == "__main__":
    main()
