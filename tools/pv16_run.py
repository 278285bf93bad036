#!/usr/bin/env python
"""One exponential-kernel attention product of shape C4 (D = E = 64, row-normalised) at N = M = argv[1] (default 65536)
through the tensor-core P.B kernel -- the target of `ncu --set full --import-source on -k regex:pv16` captures."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from kernel_matrix_benchmarks_b200 import product  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
kernel = sys.argv[2] if len(sys.argv) > 2 else "absolute-exponential"
rng = np.random.RandomState(0)
r = (3.0 / 64) ** 0.5
x = torch.tensor(r * rng.rand(n, 64), dtype=torch.float32, device="cuda")
y = torch.tensor(r * rng.rand(n, 64), dtype=torch.float32, device="cuda")
b = torch.tensor(rng.randn(n, 64), dtype=torch.float32, device="cuda")
for _ in range(2):
    out = product.kernel_product(x, y, b, kernel=kernel, normalize_rows=True)
torch.cuda.synchronize()
print("ok", float(out.abs().mean()))
