"""Conjugate gradients on (K + lambda I) b = a with a row-sharded kernel matvec.

The reference solves the dense system with LAPACK
(/root/reference/kernel_matrix_benchmarks/algorithms/bruteforce.py:205-207),
which cannot exist at N = 10^6 (the matrix would be 8 TB).  Here the matrix is
never formed: every iteration applies it through the on-the-fly product.

Sharding (BASELINE.json north_star): rows of x, r, p, Ap are split by target
across ranks; each rank needs the *whole* search direction as the source signal
of its matvec, so the one real exchange per iteration is an all-gather of p
(plus two E-float all-reduces for the dot products).

The loop is written against two small interfaces so the same host logic runs
on the GPU (``CudaShardOps`` + NCCL) and in the CPU gloo tests (tests/ inject an
oracle-backed ops object):

    ops.init(a)                  -> x, r, p, rs          (local rows; rs: E partial sums)
    ops.matvec(p_full)           -> Ap                   (local rows of K @ p_full)
    ops.shift_dot(Ap, p, lam)    -> pAp                  (Ap += lam p; partial p.Ap)
    ops.update(x, r, p, Ap, rs, pAp, rs_new)             (x += a p; r -= a Ap; partial r.r)
    ops.direction(p, r, rs_new, rs)                      (p = r + (rs_new/rs) p)
    comm.all_gather(p_local)     -> p_full (N, E)
    comm.all_reduce(t)           sums the E partials in place
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass

import torch

from . import _lib
from .product import (SYM_MIN_POINTS, Workspace, _ptr, _stream, kernel_block_f64, kernel_product, kernel_product_f64,
                      kernel_product_sym_part, symmetric_applies)


def shard_bounds(n, rank, world):
    """Contiguous row block of ``rank``: equal ceil(n / world) blocks, the last ones short or empty."""
    per = -(-n // world)
    lo = min(n, rank * per)
    return lo, min(n, lo + per), per


class LocalComm:
    """world_size == 1: nothing to exchange."""

    rank, world = 0, 1

    def all_gather(self, p_local, n_total):
        return p_local

    def all_reduce(self, t):
        return t


class ReplicatedComm(LocalComm):
    """Every rank holds ALL rows of the CG vectors (symmetric matvec split over the ranks: the exchange happens inside
    ``ops.matvec``), so nothing is gathered or reduced in the loop -- but the ranks must leave it at the same iteration,
    or the ones that stay would wait forever in the next matvec's all-reduce.  ``agree`` makes the stop decision
    collective: the loop ends only when every rank's test says so (one byte per check)."""

    def __init__(self, dist_comm):
        self.dist_comm = dist_comm

    def agree(self, flag):
        if self.dist_comm.world == 1:
            return flag
        t = flag.to(torch.int32).reshape(1)
        self.dist_comm.all_reduce_min(t)
        return t[0] != 0


class TorchDistComm:
    """One process per GPU; NCCL on the GPU box, gloo in the CPU tests."""

    def __init__(self, group=None):
        import torch.distributed as dist

        self.dist, self.group = dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self._buf = None

    def all_gather(self, p_local, n_total):
        per = -(-n_total // self.world)
        E = p_local.shape[1]
        if self._buf is None or self._buf.shape != (self.world * per, E) or self._buf.device != p_local.device:
            self._buf = torch.zeros((self.world * per, E), dtype=p_local.dtype, device=p_local.device)
            self._pad = torch.zeros((per, E), dtype=p_local.dtype, device=p_local.device)
        send = p_local
        if p_local.shape[0] != per:  # short (or empty) last shard: pad to the common size
            self._pad[: p_local.shape[0]] = p_local
            send = self._pad
        self.dist.all_gather_into_tensor(self._buf, send, group=self.group)
        return self._buf[:n_total]

    def all_reduce(self, t):
        self.dist.all_reduce(t, group=self.group)
        return t

    def all_reduce_min(self, t):
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN, group=self.group)
        return t

    def warm(self, device):
        """The first collective of a process group builds the NCCL communicator (~0.1-0.3 s): do it outside the timers."""
        t = torch.zeros(1, dtype=torch.float32, device=device)
        self.all_reduce(t)
        if t.is_cuda:
            torch.cuda.synchronize(device)


class CudaShardOps:
    """The CUDA side of one rank: product kernel for the matvec + fused CG vector kernels."""

    def __init__(self, points, kernel, row_lo, row_hi, path="auto"):
        """``points`` float32: the production path.  ``points`` float64: the `precision: float64` variant of the solver
        (the reference sweeps float16/32/64, algos.yaml:164-181) -- kmb_product_f64 as the matvec, kmb_cg_*_f64 steps."""
        self.lib = _lib.load()
        self.y = points  # (N, D) all source points, replicated
        self.x = points[row_lo:row_hi]  # this rank's target rows (a view: same memory)
        self.kernel, self.path, self.row_lo = kernel, path, row_lo
        self.n_local = row_hi - row_lo
        self.dtype = points.dtype
        self.f64 = points.dtype == torch.float64
        sfx = "f64" if self.f64 else "f32"
        self._init, self._shift_dot = getattr(self.lib, f"kmb_cg_init_{sfx}"), getattr(self.lib, f"kmb_cg_shift_dot_{sfx}")
        self._update, self._direction = getattr(self.lib, f"kmb_cg_update_{sfx}"), getattr(self.lib, f"kmb_cg_direction_{sfx}")
        self._scalar = ctypes.c_double if self.f64 else ctypes.c_float
        self.ws = Workspace()
        self.scratch = torch.zeros(int(self.lib.kmb_cg_scratch_bytes()), dtype=torch.uint8, device=points.device)
        self.launches = 0

    def _new(self, E):
        return torch.empty((self.n_local, E), dtype=self.dtype, device=self.y.device)

    def init(self, a):
        E = a.shape[1]
        x, r, p = self._new(E), self._new(E), self._new(E)
        rs = torch.empty(E, dtype=self.dtype, device=a.device)
        self.Ap = self._new(E)
        _lib.check(self._init(_ptr(a), _ptr(x), _ptr(r), _ptr(p), _ptr(rs), self.n_local, E,
                              _ptr(self.scratch), _stream()))
        self.launches += 1
        return x, r, p, rs

    def _ap(self, E):
        """The matvec's output buffer (ops.init allocates it for cg_solve; pcg_solve goes straight to matvec)."""
        if getattr(self, "Ap", None) is None or self.Ap.shape != (self.n_local, E):
            self.Ap = self._new(E)
        return self.Ap

    def matvec(self, p_full):
        self._ap(p_full.shape[1])
        if self.f64:
            kernel_product_f64(self.x, self.y, p_full.contiguous(), kernel=self.kernel, row_offset=self.row_lo, out=self.Ap)
        else:
            kernel_product(self.x, self.y, p_full, kernel=self.kernel, path=self.path, row_offset=self.row_lo,
                           out=self.Ap, workspace=self.ws)
        self.launches += int(self.lib.kmb_last_launch_count())
        return self.Ap

    def shift_dot(self, Ap, p, lam, out):
        _lib.check(self._shift_dot(_ptr(Ap), _ptr(p), self._scalar(lam), _ptr(out), self.n_local,
                                   p.shape[1], _ptr(self.scratch), _stream()))
        self.launches += 1
        return out

    def update(self, x, r, p, Ap, rs, pAp, rs_new):
        _lib.check(self._update(_ptr(x), _ptr(r), _ptr(p), _ptr(Ap), _ptr(rs), _ptr(pAp), _ptr(rs_new),
                                self.n_local, p.shape[1], _ptr(self.scratch), _stream()))
        self.launches += 1
        return rs_new

    def direction(self, p, r, rs_new, rs):
        _lib.check(self._direction(_ptr(p), _ptr(r), _ptr(rs_new), _ptr(rs), self.n_local, p.shape[1], _stream()))
        self.launches += 1


class CudaSymmetricOps(CudaShardOps):
    """CG on the symmetric product (kprod_sym): the kernel solve always has targets == sources, so
    every kernel value can serve its row and its column.  The work of one matvec is the triangular
    unit list, cut into ``world`` equal ranges; rank r evaluates range r and the ranks' partial
    results are summed with ONE all-reduce of N floats -- that is the exchange step of this mode
    (instead of the all-gather of p of the row-sharded mode).  All CG vectors are replicated (N floats
    each: their updates cost microseconds), so no other collective is needed and every rank holds the
    whole solution.  Use with ``LocalComm`` in ``cg_solve``: ``dist_comm`` is only used inside matvec.
    """

    def __init__(self, points, kernel, dist_comm=None):
        n = points.shape[0]
        super().__init__(points, kernel, 0, n)
        if not self.applies(points, kernel):
            raise NotImplementedError("the symmetric matvec needs float32 points with D <= 3")
        self.dist_comm = dist_comm if dist_comm is not None else LocalComm()

    @staticmethod
    def applies(points, kernel, E=1):
        return (points.dtype == torch.float32 and points.shape[0] >= SYM_MIN_POINTS
                and symmetric_applies(points, points, kernel, E=E))

    def matvec(self, p_full):
        if p_full.shape[1] != 1:
            raise NotImplementedError("the symmetric matvec takes one right-hand side")
        self._ap(1)
        kernel_product_sym_part(self.y, p_full, self.dist_comm.rank, self.dist_comm.world, kernel=self.kernel, out=self.Ap,
                                workspace=self.ws)
        self.launches += int(self.lib.kmb_last_launch_count())
        self.dist_comm.all_reduce(self.Ap)
        return self.Ap


@dataclass
class CgResult:
    x: torch.Tensor  # local rows of the solution
    iterations: int
    rel_residual: float  # max over right-hand sides of |r| / |a| (recurrence residual)
    converged: bool


class _LaggedFlag:
    """Convergence test without stalling the device: the flag of iteration k is copied to pinned host memory
    asynchronously and read while iteration k+1 is already enqueued, so the host stays one iteration ahead of
    the GPU (the solve then runs at most one iteration past convergence).  CPU tensors: read at once."""

    def __init__(self, device):
        self.cuda = device.type == "cuda"
        if self.cuda:
            self.host = [torch.zeros(1, dtype=torch.bool).pin_memory() for _ in range(2)]
            self.event = [torch.cuda.Event() for _ in range(2)]
        self.pending = None
        self.k = 0

    def push(self, flag):
        """Returns the value of the PREVIOUS flag (None if there is none yet)."""
        previous = self.wait()
        if not self.cuda:
            self.pending = bool(flag)
            return previous
        slot = self.k & 1
        self.host[slot].copy_(flag.reshape(1), non_blocking=True)
        self.event[slot].record()
        self.pending = slot
        self.k += 1
        return previous

    def wait(self):
        if self.pending is None:
            return None
        if not self.cuda:
            value, self.pending = self.pending, None
            return value
        slot, self.pending = self.pending, None
        self.event[slot].synchronize()
        return bool(self.host[slot].item())


def cg_solve(ops, comm, a_local, n_total, *, lam=0.0, rtol=1e-6, max_iter=500, check_every=1, lagged_check=True):
    """Solve (K + lam I) x = a for the rows this rank owns.  Every rank runs the same loop.

    ``lagged_check`` (CUDA only, check_every == 1): test convergence on the flag of the previous iteration so
    that enqueuing iteration k+1 overlaps the device work of iteration k (see _LaggedFlag); the iterate
    returned is the one the blocking check would have returned."""
    if check_every != 1:
        lagged_check = False
    x, r, p, rs = ops.init(a_local)
    comm.all_reduce(rs)
    rs0 = rs.clone()
    pAp, rs_new = torch.empty_like(rs), torch.empty_like(rs)
    tol2 = float(rtol) ** 2
    it, done = 0, bool((rs0 <= 0).all())
    lag = _LaggedFlag(rs.device) if (lagged_check and rs.device.type == "cuda") else None
    x_prev = torch.empty_like(x) if lag is not None else None
    while not done and it < max_iter:
        p_full = comm.all_gather(p, n_total)
        Ap = ops.matvec(p_full)
        comm.all_reduce(ops.shift_dot(Ap, p, float(lam), pAp))
        if lag is not None:
            x_prev.copy_(x)   # the iterate the previous flag speaks about (N floats: microseconds)
        comm.all_reduce(ops.update(x, r, p, Ap, rs, pAp, rs_new))
        ops.direction(p, r, rs_new, rs)
        rs, rs_new = rs_new, rs   # rs: residual norms after this iteration, rs_new: those before it
        it += 1
        if it % check_every == 0 or it == max_iter:
            flag = (rs <= tol2 * rs0).all()
            if hasattr(comm, "agree"):
                flag = comm.agree(flag)
            if lag is None:
                done = bool(flag)  # the only host synchronisation in the loop
            elif lag.push(flag):
                # the PREVIOUS iterate had already converged: return exactly that one (same result as the
                # blocking check; the extra iteration only cost device time that overlapped the host)
                x, rs, it, done = x_prev, rs_new, it - 1, True
    if lag is not None and not done and lag.wait():
        done = True   # the last pushed flag (iteration max_iter) was true
    safe = torch.where(rs0 > 0, rs0, torch.ones_like(rs0))
    rel = float(torch.sqrt(rs / safe).max()) if rs0.numel() else 0.0
    return CgResult(x=x, iterations=it, rel_residual=rel, converged=bool((rs <= tol2 * rs0).all()))


# ---------------------------------------------------------------------------------------------------
# Preconditioned CG (solver quality)
# ---------------------------------------------------------------------------------------------------

def landmark_indices(n, m, seed=0):
    """m distinct row indices of 0..n-1, the same on every rank (CPU generator with a fixed seed)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    return torch.randperm(int(n), generator=g)[: min(int(m), int(n))].sort().values


class NystromPreconditioner:
    """M^-1 for M = K_hat + mu I, where K_hat = C W^-1 C^T is the Nystrom approximation of the kernel matrix on m
    landmark points (C = K[:, landmarks], W = K[landmarks, landmarks]) and mu = max(lam, smallest kept eigenvalue).

    Kernel matrices of smooth kernels have a few hundred eigenvalues above lam ~ 1 even at N = 10^6 (the
    spectrum of the Gaussian kernel decays exponentially), and those are what makes plain CG take ~130
    iterations at BASELINE config 5; K_hat captures them, so the preconditioned system is close to the identity
    and CG converges in a handful of iterations (measured: tests/test_solver_gpu.py, profiles/).

    Construction (float64, once per data set -- the plugin does it in fit()):
        C_loc = k(rows of this rank, landmarks)            own CUDA kernel, kmb_kernel_block_f64
        W + nu I = L L^T, B_loc = C_loc L^-T                m x m Cholesky + triangular solve (torch / cuSOLVER)
        G = sum over ranks of B_loc^T B_loc = V S V^T       one all-reduce of m^2 doubles, m x m eigh
        U_loc = B_loc V_r S_r^-1/2 (orthonormal columns over all rows), eigenvalues S_r of K_hat
    Application per CG iteration (float32, two tall-skinny products, one all-reduce of r x E floats):
        M^-1 v = v / mu + U ((1 / (S + mu) - 1 / mu) (U^T v))

    ``block_fn(points_local, landmarks) -> (n_loc, m) float64`` is injectable so the CPU (gloo) tests can run the
    same host logic with an oracle-backed block.
    """

    def __init__(self, points_local, landmarks, kernel, lam, comm=None, *, rank_tol=1e-10, block_fn=None, dtype=torch.float32):
        comm = comm if comm is not None else LocalComm()
        self.comm = comm
        if block_fn is None:
            def block_fn(p, q):
                return kernel_block_f64(p, q, kernel=kernel)
        lm = landmarks.to(torch.float64).contiguous()
        m = lm.shape[0]
        C = block_fn(points_local.to(torch.float64).contiguous(), lm)          # (n_loc, m)
        W = block_fn(lm, lm)                                                    # (m, m), the same on every rank
        W = 0.5 * (W + W.T)
        nu = 1e-12 * float(torch.trace(W))   # jitter: W is numerically singular for smooth kernels
        L = torch.linalg.cholesky(W + nu * torch.eye(m, dtype=W.dtype, device=W.device))
        B = torch.linalg.solve_triangular(L, C.T, upper=False).T               # C L^-T  (n_loc, m)
        del C
        G = B.T @ B
        comm.all_reduce(G)
        S, V = torch.linalg.eigh(G)
        keep = S > rank_tol * S.max()
        S, V = S[keep], V[:, keep]
        self.U = ((B @ V) / torch.sqrt(S)).to(dtype).contiguous()              # (n_loc, r)
        del B
        self.eigenvalues = S
        self.rank = int(S.numel())
        self.set_shift(lam)

    def replicate(self, n_total):
        """Row-sharded build, replicated application (the symmetric matvec keeps every CG vector whole on every rank): one
        all-gather of U (n x r floats over NVLink) after a build that cost each rank 1 / world of the work; from then on
        ``apply`` needs no collective."""
        if self.comm.world > 1:
            self.U = self.comm.all_gather(self.U, n_total).clone()
            self.comm = LocalComm()
        return self

    def set_shift(self, lam):
        S = self.eigenvalues
        self.mu = max(float(lam), float(S.min())) if S.numel() else max(float(lam), 1.0)
        self.coef = (1.0 / (S + self.mu) - 1.0 / self.mu).to(self.U.dtype).unsqueeze(1)    # (r, 1)

    def apply(self, v_local):
        t = self.U.T @ v_local                                                  # (r, E) partial sums over this rank's rows
        self.comm.all_reduce(t)
        return v_local / self.mu + self.U @ (self.coef * t)


def pcg_solve(ops, comm, a_local, n_total, precond, *, lam=0.0, rtol=1e-6, max_iter=500):
    """Preconditioned CG on (K + lam I) x = a for the rows this rank owns; ``precond.apply`` is M^-1 on local rows.
    Same ops / comm interfaces as cg_solve and the same fused vector kernels: with z = M^-1 r the step length is
    (r.z) / (p.Ap) and the new direction z + ((r.z)_new / (r.z)) p, so ``ops.update`` and ``ops.direction`` are handed
    r.z where plain CG hands them r.r (the update still returns the new r.r for the convergence test).  Convergence
    is tested on the recurrence residual |r| / |a| like cg_solve, with a blocking check every iteration (a
    preconditioned solve takes a handful of iterations)."""
    x, r, p, rs = ops.init(a_local)          # x = 0, r = a, p = a (overwritten below), rs = partial r.r
    comm.all_reduce(rs)
    rs0 = rs.clone()
    pAp, rs_new = torch.empty_like(rs), torch.empty_like(rs)
    tol2 = float(rtol) ** 2
    it, done = 0, bool((rs0 <= 0).all())
    if not done:
        z = precond.apply(r)
        p.copy_(z)
        rz = (r * z).sum(0)
        comm.all_reduce(rz)
    while not done and it < max_iter:
        p_full = comm.all_gather(p, n_total)
        Ap = ops.matvec(p_full)
        comm.all_reduce(ops.shift_dot(Ap, p, float(lam), pAp))           # Ap += lam p; p.Ap
        comm.all_reduce(ops.update(x, r, p, Ap, rz, pAp, rs_new))        # alpha = r.z / p.Ap; x, r; new r.r
        rs, rs_new = rs_new, rs
        it += 1
        flag = (rs <= tol2 * rs0).all()
        done = bool(comm.agree(flag) if hasattr(comm, "agree") else flag)
        if done:
            break
        z = precond.apply(r)
        rz_new = (r * z).sum(0)
        comm.all_reduce(rz_new)
        ops.direction(p, z, rz_new, rz)                                  # p = z + (r.z_new / r.z) p
        rz = rz_new
    safe = torch.where(rs0 > 0, rs0, torch.ones_like(rs0))
    rel = float(torch.sqrt(rs / safe).max()) if rs0.numel() else 0.0
    return CgResult(x=x, iterations=it, rel_residual=rel, converged=bool((rs <= tol2 * rs0).all()))
