"""
2-D block-cyclic distributed Cholesky of K + eta*I for matrices that do not fit one GPU (BASELINE.json configs[4]:
dense Matern n = 100 000 on 2 / 4 / 8 B200), with the panel exchanged by NCCL broadcasts.

Layout: global padded size npad = NB * nb; block (i, j) lives on process (i mod P_r, j mod P_c); each rank stores its
blocks as one dense local matrix (local block rows x local block cols, row-major). K is generated directly in this
layout (every rank evaluates only its own tiles, csrc/gp_matern.cu `gp_matern_cross`).

Right-looking factorisation, per block column k:
  1. the owner of (k, k) factors the nb x nb diagonal block and inverts its factor (gp_potrf_f64 + gp_trtri_f64),
  2. inv(L_kk) is broadcast down process column k mod P_c; those ranks form their panel blocks L_ik = A_ik inv(L_kk)^T
     (DMMA GEMM),
  3. each process row's stack of panel blocks is broadcast to every rank (P_r broadcasts of (n-k nb)/P_r x nb),
  4. every rank updates its local trailing blocks A_ij -= L_ik L_jk^T, i >= j > k (one DMMA GEMM per local block column).
log det = 2 sum log diag(L_kk) (all-reduce); solves use the stored inv(L_kk) blocks with one small all-reduce and one
broadcast per block step (right-hand sides are replicated).

The compute primitives are injected through `ops` (GpuOps = libgpgp kernels). tests/ inject a NumPy implementation to
exercise the distributed algorithm on CPU with the gloo backend; the product has no CPU path.
"""

import ctypes

import numpy

from . import _device as dev
from ._device import lib, check

__all__ = ['process_grid', 'GpuOps', 'BlockCyclicCholesky']


def process_grid(world):
    """P_r x P_c with P_r <= P_c, as square as possible (1x2, 2x2, 2x4, ...)."""
    pr = int(numpy.floor(numpy.sqrt(world)))
    while world % pr:
        pr -= 1
    return pr, world // pr


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


class GpuOps(object):
    """Primitives on torch float64 CUDA tensors, all through the libgpgp C ABI (views: unit column stride)."""

    def __init__(self):
        self.torch = dev.require_cuda()
        self.device = self.torch.device('cuda', self.torch.cuda.current_device())
        self._ws = {}

    def empty(self, shape):
        return self.torch.empty(shape, dtype=self.torch.float64, device=self.device)

    def zeros(self, shape):
        return self.torch.zeros(shape, dtype=self.torch.float64, device=self.device)

    def from_host(self, a, dtype=None):
        t = self.torch.from_numpy(numpy.ascontiguousarray(a))
        return t.to(self.device) if dtype is None else t.to(self.device, dtype=dtype)

    def to_host(self, t):
        return t.cpu().numpy()

    def generate(self, prow, pcol, rg, cg, n, scale, nu, eta, out):
        check(lib.gp_matern_cross(_p(prow), _p(pcol), _p(rg), _p(cg), out.shape[0], out.shape[1], n, prow.shape[1],
                                  dev.host_ptr(scale), float(nu), float(eta), _p(out), out.stride(0), dev.stream_ptr()),
              'gp_matern_cross')

    def potrf_inv(self, D, nvalid):
        """D (nb x nb contiguous): lower Cholesky in place; returns inv(L) (nb x nb, zero above the diagonal)."""
        torch = self.torch
        nb = D.shape[0]
        if nb not in self._ws:
            self._ws[nb] = (torch.empty(lib.gp_potrf_workspace_bytes(nb) // 8, dtype=torch.float64, device=self.device),
                            torch.empty(lib.gp_potri_workspace_bytes(nb) // 8 + 8, dtype=torch.float64, device=self.device),
                            torch.zeros(1, dtype=torch.int32, device=self.device))
        pws, tws, info = self._ws[nb]
        s = dev.stream_ptr()
        check(lib.gp_potrf_f64(_p(D), int(nvalid), nb, _p(info), _p(pws), s), 'gp_potrf_f64')
        W = torch.zeros((nb, nb), dtype=torch.float64, device=self.device)
        check(lib.gp_trtri_f64(_p(D), _p(W), nb, _p(pws), _p(tws), s), 'gp_trtri_f64')
        bad = int(info.item())
        return W, bad

    def gemm_nt(self, C, A, B, alpha, beta):
        """C = beta C + alpha A B^T on (row-stride) views."""
        check(lib.gp_dgemm_f64(0, 0, _p(C), C.stride(0), _p(A), A.stride(0), _p(B), B.stride(0), C.shape[0], C.shape[1],
                               A.shape[1], float(alpha), float(beta), 0, 0, dev.stream_ptr()), 'gp_dgemm_f64')

    def logdet_chol(self, D):
        out = self.empty(1)
        check(lib.gp_logdet_from_chol(_p(D), D.shape[0], D.shape[0], _p(out), dev.stream_ptr()), 'gp_logdet_from_chol')
        return float(out.item())

    # skinny (n x p) products of the distributed substitution: library GEMV-class calls, O(n^2 p) flop in total
    def matmul(self, A, X):
        return self.torch.matmul(A, X)

    def matmul_t(self, A, X):
        return self.torch.matmul(A.transpose(0, 1), X)


class BlockCyclicCholesky(object):

    def __init__(self, points, correlation_scale, nu, nb=1024, ops=None, dist=None):
        if nb % 128:
            raise ValueError('nb should be a multiple of 128')
        if dist is None:
            import torch.distributed as dist
        self.dist = dist
        self.distributed = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank() if self.distributed else 0
        self.world = dist.get_world_size() if self.distributed else 1
        self.P_r, self.P_c = process_grid(self.world)
        self.r, self.c = divmod(self.rank, self.P_c)
        self.ops = ops if ops is not None else GpuOps()
        self.points = numpy.ascontiguousarray(points, dtype=numpy.float64)
        self.n, self.d = self.points.shape
        if numpy.isscalar(correlation_scale):
            correlation_scale = numpy.repeat(float(correlation_scale), self.d)
        self.scale = dev.host_f64(correlation_scale)
        self.nu = float(nu)
        self.nb = int(nb)
        self.NB = (self.n + nb - 1) // nb
        self.npad = self.NB * nb
        self.I_loc = [i for i in range(self.NB) if i % self.P_r == self.r]
        self.J_loc = [j for j in range(self.NB) if j % self.P_c == self.c]
        self.col_groups = None
        if self.distributed and self.world > 1:
            # one group per process column (for inv(L_kk)); every rank must create every group
            self.col_groups = [dist.new_group([rr * self.P_c + cc for rr in range(self.P_r)]) for cc in range(self.P_c)]
        self.Aloc = None
        self.Linv = {}
        self.bytes_received = 0

    # ---- helpers -------------------------------------------------------------------------------------------------
    def _gidx(self, blocks):
        nb = self.nb
        return numpy.concatenate([numpy.arange(b * nb, (b + 1) * nb) for b in blocks]).astype(numpy.int32) \
            if blocks else numpy.zeros(0, dtype=numpy.int32)

    def _bcast(self, t, src, group=None):
        if self.world > 1:
            self.dist.broadcast(t, src=src, group=group)
            if self.rank != src:
                self.bytes_received += t.numel() * 8

    def _allreduce(self, t):
        if self.world > 1:
            self.dist.all_reduce(t)

    def rows_after(self, k, P, me):
        """number of block indices i > k with i mod P == me"""
        return len([i for i in range(k + 1, self.NB) if i % P == me])

    # ---- generation in block-cyclic layout ---------------------------------------------------------------------------
    def generate(self, eta):
        ops, nb = self.ops, self.nb
        rg, cg = self._gidx(self.I_loc), self._gidx(self.J_loc)
        pad = numpy.zeros((self.npad, self.d))
        pad[:self.n] = self.points
        self.Aloc = ops.empty((len(rg), len(cg)))
        if len(rg) and len(cg):
            ops.generate(ops.from_host(pad[rg]), ops.from_host(pad[cg]), ops.from_host(rg), ops.from_host(cg), self.n,
                         self.scale, self.nu, eta, self.Aloc)

    # ---- factorisation ---------------------------------------------------------------------------------------------
    def factor(self, eta):
        """Generates K + eta I in place and factors it. Raises numpy.linalg.LinAlgError (on every rank) if not PD."""
        self.generate(eta)
        ops, nb, NB, P_r, P_c = self.ops, self.nb, self.NB, self.P_r, self.P_c
        A = self.Aloc
        self.Linv = {}
        self.bytes_received = 0
        bad = ops.zeros(1)
        bufs = [None] * P_r
        for k in range(NB):
            pr, pc = k % P_r, k % P_c
            owner = pr * P_c + pc
            Linv = None
            if self.c == pc:
                if self.rank == owner:
                    li, lj = self.I_loc.index(k), self.J_loc.index(k)
                    blk = A[li * nb:(li + 1) * nb, lj * nb:(lj + 1) * nb]
                    D = blk.contiguous()
                    nvalid = max(0, min(nb, self.n - k * nb))
                    Linv, info = ops.potrf_inv(D, nvalid)
                    blk.copy_(D)
                    self.Linv[k] = (Linv, D)
                    if info:
                        bad += float(k * nb + info)
                else:
                    Linv = ops.empty((nb, nb))
                self._bcast(Linv, owner, self.col_groups[pc] if self.col_groups else None)
            # panel blocks below the diagonal, stacked per process row
            for rr in range(P_r):
                cnt = self.rows_after(k, P_r, rr)
                if cnt == 0:
                    bufs[rr] = None
                    continue
                root = rr * P_c + pc
                buf = ops.empty((cnt * nb, nb))
                if self.rank == root:
                    l0 = len(self.I_loc) - cnt
                    lj = self.J_loc.index(k)
                    panel = A[l0 * nb:, lj * nb:(lj + 1) * nb]
                    ops.gemm_nt(buf, panel, Linv, 1.0, 0.0)
                    panel.copy_(buf)
                self._bcast(buf, root)
                bufs[rr] = buf
            # trailing update of the local blocks (i >= j > k)
            mine = bufs[self.r]
            if mine is not None:
                my_rows = [i for i in self.I_loc if i > k]
                for lj, j in enumerate(self.J_loc):
                    if j <= k:
                        continue
                    rows = [i for i in my_rows if i >= j]
                    if not rows:
                        continue
                    skip = len(my_rows) - len(rows)
                    src = bufs[j % P_r]
                    pos = len([i for i in range(k + 1, j) if i % P_r == j % P_r])
                    Lj = src[pos * nb:(pos + 1) * nb]
                    l0 = len(self.I_loc) - len(rows)
                    C = A[l0 * nb:, lj * nb:(lj + 1) * nb]
                    ops.gemm_nt(C, mine[skip * nb:], Lj, -1.0, 1.0)
        self._allreduce(bad)
        info = float(ops.to_host(bad)[0])
        if info != 0.0:
            raise numpy.linalg.LinAlgError('K + eta*I (eta=%g) is not positive definite (block-cyclic potrf).' % eta)
        self.eta = float(eta)

    def logdet(self):
        s = self.ops.zeros(1)
        for k, (Linv, D) in self.Linv.items():
            nvalid = max(0, min(self.nb, self.n - k * self.nb))
            if nvalid == self.nb:
                s += self.ops.logdet_chol(D)
            elif nvalid > 0:
                s += 2.0 * float(numpy.sum(numpy.log(numpy.diag(self.ops.to_host(D))[:nvalid])))
        self._allreduce(s)
        return float(self.ops.to_host(s)[0])

    # ---- solves with replicated right-hand sides -------------------------------------------------------------------
    def solve(self, R):
        """(K + eta I)^-1 R for a host array R (n,) or (n, p); returns a host array (identical on every rank)."""
        ops, nb, NB, P_r, P_c = self.ops, self.nb, self.NB, self.P_r, self.P_c
        R = numpy.asarray(R, dtype=numpy.float64)
        vec = (R.ndim == 1)
        R2 = R.reshape(self.n, -1)
        p = R2.shape[1]
        Bp = numpy.zeros((self.npad, p))
        Bp[:self.n] = R2
        b = ops.from_host(Bp)
        A = self.Aloc
        # forward: L y = b
        acc = ops.zeros((self.npad, p))
        y = ops.zeros((self.npad, p))
        for k in range(NB):
            pr, pc = k % P_r, k % P_c
            owner = pr * P_c + pc
            t = acc[k * nb:(k + 1) * nb].clone()
            self._allreduce(t)
            yk = ops.empty((nb, p))
            if self.rank == owner:
                yk.copy_(ops.matmul(self.Linv[k][0], b[k * nb:(k + 1) * nb] - t))
            self._bcast(yk, owner)
            y[k * nb:(k + 1) * nb] = yk
            if self.c == pc:
                rows = [i for i in self.I_loc if i > k]
                if rows:
                    l0 = len(self.I_loc) - len(rows)
                    lj = self.J_loc.index(k)
                    upd = ops.matmul(A[l0 * nb:, lj * nb:(lj + 1) * nb], yk)
                    for q, i in enumerate(rows):
                        acc[i * nb:(i + 1) * nb] += upd[q * nb:(q + 1) * nb]
        # backward: L^T x = y
        x = ops.zeros((self.npad, p))
        for k in range(NB - 1, -1, -1):
            pr, pc = k % P_r, k % P_c
            owner = pr * P_c + pc
            t = ops.zeros((nb, p))
            if self.c == pc:
                rows = [i for i in self.I_loc if i > k]
                if rows:
                    l0 = len(self.I_loc) - len(rows)
                    lj = self.J_loc.index(k)
                    xs = ops.empty((len(rows) * nb, p))
                    for q, i in enumerate(rows):
                        xs[q * nb:(q + 1) * nb] = x[i * nb:(i + 1) * nb]
                    t = ops.matmul_t(A[l0 * nb:, lj * nb:(lj + 1) * nb], xs)
            t = t.contiguous()
            self._allreduce(t)
            xk = ops.empty((nb, p))
            if self.rank == owner:
                xk.copy_(ops.matmul_t(self.Linv[k][0], y[k * nb:(k + 1) * nb] - t))
            self._bcast(xk, owner)
            x[k * nb:(k + 1) * nb] = xk
        out = ops.to_host(x)[:self.n]
        return out[:, 0] if vec else out

    # ---- profile log-likelihood at (sigma_hat(eta), eta) from the distributed factor ------------------------------------
    def profile_log_likelihood(self, z, X, eta):
        """l^(sigma_hat, eta) exactly as ProfileLikelihood.log_likelihood (reference _profile_likelihood.py:38-85)
        evaluated at sigma_hat^2 = z^T M z / (n - m); returns (l^, sigma_hat)."""
        self.factor(eta)
        n, m = X.shape
        R = numpy.c_[X, z]
        S = self.solve(R)
        G = R.T @ S
        B = G[:m, :m]
        beta = numpy.linalg.solve(B, G[:m, m])
        zMz = G[m, m] - G[:m, m] @ beta
        sigma2 = zMz / (n - m)
        lp = -0.5 * (n - m) * numpy.log(sigma2) - 0.5 * self.logdet() - 0.5 * numpy.log(numpy.linalg.det(B)) - 0.5 * (n - m)
        return float(lp), float(numpy.sqrt(sigma2))
