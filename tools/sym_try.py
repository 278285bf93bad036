import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from kernel_matrix_benchmarks_b200 import product, datasets
from oracle import bruteforce_oracle as orc
torch.manual_seed(0)
dev = torch.device('cuda')
def rel(a, b): return float(np.linalg.norm(a - b) / np.linalg.norm(b))
for n in [300, 2048, 5000, 16384, 20001, 65536+17]:
    for D in [3, 2, 1]:
        ds = datasets.uniform_cube(n, D, 1.0, 'gaussian')
        y = torch.tensor(ds.source_points, dtype=torch.float32, device=dev)
        b = torch.tensor(ds.source_signal, dtype=torch.float32, device=dev)
        ref = product.kernel_product(y, y, b, path='direct').cpu().numpy().astype(np.float64)
        got = product.kernel_product(y, y, b, path='direct_sym').cpu().numpy().astype(np.float64)
        rows = np.arange(0, n, max(1, n // 256))
        want = orc.kernel_product('gaussian', ds.source_points, None, ds.source_signal, rows=rows)
        line = f"n={n} D={D} sym-vs-direct {rel(got, ref):.2e} sym-vs-oracle {rel(got[rows], want):.2e} direct-vs-oracle {rel(ref[rows], want):.2e}"
        for parts in (2, 3, 5):
            tot = sum(product.kernel_product_sym_part(y, b, p, parts).clone() for p in range(parts)).cpu().numpy().astype(np.float64)
            line += f" | {parts} parts {rel(tot[rows], want):.2e}"
        print(line, flush=True)
# spread-out data -> difference form fallback
n = 20000
ds = datasets.uniform_cube(n, 3, 6.0, 'gaussian')
y = torch.tensor(ds.source_points, dtype=torch.float32, device=dev); b = torch.tensor(ds.source_signal, dtype=torch.float32, device=dev)
rows = np.arange(0, n, 100); want = orc.kernel_product('gaussian', ds.source_points, None, ds.source_signal, rows=rows)
got = product.kernel_product(y, y, b, path='direct_sym').cpu().numpy().astype(np.float64)
tot = sum(product.kernel_product_sym_part(y, b, p, 3).clone() for p in range(3)).cpu().numpy().astype(np.float64)
print('fallback (radius 6):', product.direct_stats()['form'], rel(got[rows], want), rel(tot[rows], want))
# determinism
a1 = product.kernel_product(y, y, b, path='direct_sym').clone(); a2 = product.kernel_product(y, y, b, path='direct_sym').clone()
print('deterministic:', bool((a1 == a2).all()))
for n in [262144, 1000000]:
    ds = datasets.uniform_cube(n, 3, 1.0, 'gaussian')
    y = torch.tensor(ds.source_points, dtype=torch.float32, device=dev); b = torch.tensor(ds.source_signal, dtype=torch.float32, device=dev)
    for path in ['direct', 'direct_sym']:
        for _ in range(2): product.kernel_product(y, y, b, path=path)
        torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): out = product.kernel_product(y, y, b, path=path)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print(f"n={n} {path}: {ms:.3f} ms  {n*n/ms/1e6:.0f} Gpairs/s", flush=True)
    rows = np.sort(np.random.RandomState(0).choice(n, 256, replace=False))
    want = orc.kernel_product('gaussian', ds.source_points, None, ds.source_signal, rows=rows)
    got = product.kernel_product(y, y, b, path='direct_sym').cpu().numpy().astype(np.float64)
    print(f"n={n} sym vs oracle {rel(got[rows], want):.2e}")
