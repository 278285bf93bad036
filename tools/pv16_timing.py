"""Per-phase clock64 averages of the pv16 kernel (instrumented build: see csrc/Makefile, KMB_PV16_TIMING)."""
import os, sys, numpy as np, torch
sys.path.insert(0, '.')
from kernel_matrix_benchmarks_b200 import _lib
if not os.environ.get('KMB_B200_LIB'):
    _lib.LIB_PATH = os.path.join(os.path.dirname(_lib.LIB_PATH), 'libkmb_b200_timing.so')
from kernel_matrix_benchmarks_b200 import product
rng = np.random.RandomState(0)
D, E = 64, 64
N = M = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
r = (3.0 / D) ** 0.5
x = torch.tensor(r * rng.rand(N, D), dtype=torch.float32, device='cuda'); y = torch.tensor(r * rng.rand(M, D), dtype=torch.float32, device='cuda')
b = torch.tensor(rng.randn(M, E), dtype=torch.float32, device='cuda')
names = ['loop top', 'wait S', 'LDTM', 'log2 k', 'rescale', 'P', 'st+arrive', 'blocks', 'MMA: other', 'wait P', 'wait signal', 'MMA: issue', 'wait V', 'cycles/block', 'MHz']
for kernel in ('gaussian', 'absolute-exponential'):
    for rep in range(2):
        out = product.kernel_product(x, y, b, kernel=kernel, normalize_rows=True)
    torch.cuda.synchronize()
    v = out[0, :15].cpu().numpy()
    print(kernel, ' '.join(f'{n}={a:.0f}' for n, a in zip(names, v)), 'epilogue total', v[:7].sum())
