import sys, numpy as np, torch
sys.path.insert(0, '.')
from kernel_matrix_benchmarks_b200 import product
rng = np.random.RandomState(0)
D, E = 32, 8
N, M = 128, 128
r = (3.0 / D) ** 0.5
xh, yh = r * rng.rand(N, D), r * rng.rand(M, D)
for mode in ("ones", "onehot"):
    for src in range(0, 128, 9) if mode == "onehot" else [0]:
        b = np.zeros((M, E)); 
        if mode == "ones": b[:] = 1.0
        else: b[src, :] = 1.0
        x = torch.tensor(xh, dtype=torch.float32, device='cuda'); y = torch.tensor(yh, dtype=torch.float32, device='cuda')
        bt = torch.tensor(b, dtype=torch.float32, device='cuda')
        out = product.kernel_product(x, y, bt, kernel='gaussian').cpu().numpy().astype(np.float64)
        K = np.exp(-((xh[:, None, :] - yh[None, :, :]) ** 2).sum(-1))
        want = K @ b
        ratio = out[:, 0] / want[:, 0]
        print(mode, src, 'ratio min/max', ratio.min(), ratio.max(), 'rows0-3', ratio[:4])
