"""
Distributed dense Cholesky + log-likelihood GRADIENT for matrices that do not fit one GPU (BASELINE.json configs[4]:
dense Matern n = 100 000 on 2 / 4 / 8 B200), panels exchanged by NCCL broadcasts over NVLink. The reference
is single-node LAPACK (gaussian_proc/_mixed_correlation/_linear_solver.py:71, mixed_correlation.py:183-191,250-261); the
quantities produced here are those of _likelihood/_profile_likelihood.py:38-132 plus the d/d rho extension.

Layout (B200-first: NVSwitch moves a 400 MB panel in about a millisecond and every GPU has 180 GB, so the factor is
REPLICATED while the work is partitioned):
  * block size nb, NB = ceil(n / nb) block rows / columns, identity padding to npad = NB nb;
  * block-cyclic ownership on a 1 x P process grid (the P_r = 1 case of the 2-D layout) with boustrophedon ("snake")
    order: block column j belongs to rank j mod P on even rounds and P - 1 - (j mod P) on odd ones, so that every pair of
    rounds hands each rank the same amount of trailing-update work; K + eta I is generated directly in this layout;
  * every finished panel k (the diagonal block L_kk, the block column below it and inv(L_kk)) is broadcast ONCE to all
    ranks and KEPT: after the factorisation each rank holds the whole factor L (n^2/2 doubles, 40 GB at n = 100 000).

Factorisation (right-looking, look-ahead 1): the owner of block column k+1 applies panel k to that column first, factors
its diagonal block (gp_potrf_f64 + gp_trtri_f64), forms the panel with one DMMA GEMM and broadcasts it on a high-priority
side stream while every rank - the owner included - still applies panel k to its other columns on the main stream. The
broadcast of panel k+1 therefore overlaps trailing update k; nothing on the main stream ever waits for NCCL except the
first use of a panel.

Gradient (no further panel traffic): with L replicated, the rows of W = inv(L) are independent (W_i L = e_i^T), so block
ROW i of W belongs to rank snake(i) and is obtained by a block back-substitution from the right made of two DMMA GEMMs
per block column (the staircase of zeros is skipped with per-row-tile k ranges, gp_dgemm_ktab_f64). Then
    tr Kn^-1      = ||W||_F^2                       (local sums, one all-reduce)
    tr Kn^-1 dK   = <W^T W, dK> = sum_ranks <X_r^T X_r, dK>,  X_r = this rank's rows of W:
                    block column by block column, X_r^T X_r[:, J] is one TN DMMA GEMM and dK/drho[:, J] is regenerated
                    from the points (never stored), reduced at once by a weighted Frobenius inner product;
    S = Kn^-1 [X z] by block substitution on the replicated factor (local, identical on every rank),
    G = R^T S, H = S^T S, Q = S^T dK S (dK S regenerated on the fly) -> the host algebra of _likelihood/_fused.py.
Flops per rank: (n^3/3 + n^3/3 + n^3/3) / P, the same n^3 as the single-GPU evaluator.

The compute primitives are injected through `ops` (GpuOps = libgpgp kernels through the C ABI). tests/ inject a NumPy
implementation to exercise the distributed algorithm on CPU with the gloo backend; the product has no CPU path.
"""

import contextlib
import ctypes
import time

import numpy

from . import _device as dev
from ._device import lib, check

__all__ = ['process_grid', 'snake_owner', 'GpuOps', 'BlockCyclicCholesky']


def process_grid(world):
    """(P_r, P_c) of the block-cyclic layout: 1 x world. The factor is replicated by the panel broadcasts (every rank needs
    every panel for the gradient phase), so a taller grid would not save traffic; the snake order balances the columns."""
    return 1, int(world)


def snake_owner(j, P):
    """Owner of block index j on P ranks in boustrophedon order."""
    r, pos = divmod(int(j), int(P))
    return pos if r % 2 == 0 else P - 1 - pos


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


class GpuOps(object):
    """Primitives on torch float64 CUDA tensors (row-major views with unit column stride), all through the libgpgp C ABI.
    torch owns buffers, streams and events only."""

    def __init__(self):
        self.torch = dev.require_cuda()
        self.device = self.torch.device('cuda', self.torch.cuda.current_device())
        self._ws = {}
        self._rect_ws = None

    # ---- buffers ------------------------------------------------------------------------------------------------
    def empty(self, shape):
        return self.torch.empty(shape, dtype=self.torch.float64, device=self.device)

    def zeros(self, shape):
        return self.torch.zeros(shape, dtype=self.torch.float64, device=self.device)

    def from_host(self, a, dtype=None):
        t = self.torch.from_numpy(numpy.ascontiguousarray(a))
        return t.to(self.device) if dtype is None else t.to(self.device, dtype=dtype)

    def int_tensor(self, values):
        return self.torch.tensor(list(values), dtype=self.torch.int32, device=self.device)

    def to_host(self, t):
        return t.cpu().numpy()

    # ---- streams / events ------------------------------------------------------------------------------------------
    def new_stream(self):
        return self.torch.cuda.Stream(priority=-1)

    def use(self, stream):
        return self.torch.cuda.stream(stream)

    def event(self, stream=None, timing=False):
        e = self.torch.cuda.Event(enable_timing=timing)
        e.record(stream if stream is not None else self.torch.cuda.current_stream())
        return e

    def wait(self, stream, event):
        (stream if stream is not None else self.torch.cuda.current_stream()).wait_event(event)

    def elapsed_s(self, e0, e1):
        return e0.elapsed_time(e1) * 1e-3

    def synchronize(self):
        self.torch.cuda.synchronize()

    # ---- kernels ---------------------------------------------------------------------------------------------------
    def generate(self, prow, pcol, rg, cg, n, scale, nu, eta, out):
        check(lib.gp_matern_cross(_p(prow), _p(pcol), _p(rg), _p(cg), out.shape[0], out.shape[1], n, prow.shape[1],
                                  dev.host_ptr(scale), float(nu), float(eta), _p(out), out.stride(0), dev.stream_ptr()),
              'gp_matern_cross')

    def generate_dk(self, prow, pcol, rg, cg, n, scale, nu, out):
        check(lib.gp_matern_cross_dk(_p(prow), _p(pcol), _p(rg), _p(cg), out.shape[0], out.shape[1], n, prow.shape[1],
                                     dev.host_ptr(scale), float(nu), _p(out), out.stride(0), dev.stream_ptr()),
              'gp_matern_cross_dk')

    def potrf_inv(self, D, nvalid, Linv):
        """D (nb x nb contiguous): lower Cholesky in place; Linv (nb x nb contiguous) <- inv(L), zero above the diagonal.
        Returns a device int tensor with the LAPACK-style info (no host synchronisation)."""
        torch = self.torch
        nb = D.shape[0]
        if nb not in self._ws:
            self._ws[nb] = (torch.empty(lib.gp_potrf_workspace_bytes(nb) // 8, dtype=torch.float64, device=self.device),
                            torch.empty(lib.gp_potri_workspace_bytes(nb) // 8 + 8, dtype=torch.float64, device=self.device))
        pws, tws = self._ws[nb]
        info = torch.zeros(1, dtype=torch.int32, device=self.device)
        s = dev.stream_ptr()
        check(lib.gp_potrf_f64(_p(D), int(nvalid), nb, _p(info), _p(pws), s), 'gp_potrf_f64')
        Linv.zero_()
        check(lib.gp_trtri_f64(_p(D), _p(Linv), nb, _p(pws), _p(tws), s), 'gp_trtri_f64')
        return info

    def gemm(self, C, A, B, alpha, beta, at=0, bt=0, kbeg=None, kend=None):
        """C = beta C + alpha op(A) op(B); at = 0: A is the (M x K) view, at = 1: A is the (K x M) view; bt = 0: B is the
        (N x K) view, bt = 1: B is the (K x N) view. kbeg / kend: per-128-row-tile k ranges (device int32)."""
        M, N = C.shape
        K = A.shape[1] if at == 0 else A.shape[0]
        if kbeg is None and kend is None:
            check(lib.gp_dgemm_f64(at, bt, _p(C), C.stride(0), _p(A), A.stride(0), _p(B), B.stride(0), M, N, K, float(alpha),
                                   float(beta), 0, 0, dev.stream_ptr()), 'gp_dgemm_f64')
        else:
            check(lib.gp_dgemm_ktab_f64(at, bt, _p(C), C.stride(0), _p(A), A.stride(0), _p(B), B.stride(0), M, N, K,
                                        float(alpha), float(beta), _p(kbeg) if kbeg is not None else None,
                                        _p(kend) if kend is not None else None, dev.stream_ptr()), 'gp_dgemm_ktab_f64')

    def logdet_chol(self, D, nvalid, out):
        """out (device, 1 double) <- 2 sum_{i < nvalid} log D_ii for the lower factor D (row stride D.stride(0))"""
        check(lib.gp_logdet_from_chol(_p(D), int(nvalid), D.stride(0), _p(out), dev.stream_ptr()), 'gp_logdet_from_chol')

    def _rect_workspace(self, M, N, p):
        need = int(lib.gp_rect_workspace_bytes(M, N, p)) // 8 + 8
        if self._rect_ws is None or self._rect_ws.numel() < need:
            self._rect_ws = self.empty(need)
        return self._rect_ws

    def rect_apply(self, X, R, Y, alpha=1.0, beta=0.0):
        """Y = alpha X R + beta Y  (X: M x N slab, R: N x p, p <= 16)"""
        check(lib.gp_rect_apply(_p(X), X.shape[0], X.shape[1], X.stride(0), _p(R), R.shape[1], R.stride(0), _p(Y), Y.stride(0),
                                float(alpha), float(beta), dev.stream_ptr()), 'gp_rect_apply')

    def rect_apply_t(self, X, Y, S, alpha=1.0, beta=0.0):
        """S = alpha X^T Y + beta S  (X: M x N slab, Y: M x p, S: N x p)"""
        ws = self._rect_workspace(X.shape[0], X.shape[1], Y.shape[1])
        check(lib.gp_rect_apply_t(_p(X), X.shape[0], X.shape[1], X.stride(0), _p(Y), Y.shape[1], Y.stride(0), _p(S), S.stride(0),
                                  float(alpha), float(beta), _p(ws), dev.stream_ptr()), 'gp_rect_apply_t')

    def pair_dot(self, A, B, rows_w1, w_rest, accum):
        """accum[0] += sum_{r < rows_w1} <A_r, B_r> + w_rest sum_{r >= rows_w1} <A_r, B_r>"""
        ws = self._rect_workspace(1, 1, 1)
        check(lib.gp_pair_dot(_p(A), A.stride(0), _p(B), B.stride(0), A.shape[0], A.shape[1], int(rows_w1), float(w_rest),
                              _p(accum), _p(ws), dev.stream_ptr()), 'gp_pair_dot')

    def dk_apply(self, points, n, scale, nu, S, V):
        """V = dK/drho S (dK regenerated from the points)"""
        check(lib.gp_dk_apply(_p(points), n, points.shape[1], dev.host_ptr(scale), float(nu), _p(S), S.shape[1], S.stride(0),
                              _p(V), dev.stream_ptr()), 'gp_dk_apply')


class BlockCyclicCholesky(object):

    def __init__(self, points, correlation_scale, nu, nb=512, ops=None, dist=None):
        if nb % 128:
            raise ValueError('nb should be a multiple of 128')
        if dist is None:
            import torch.distributed as dist
        self.dist = dist
        self.distributed = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank() if self.distributed else 0
        self.world = dist.get_world_size() if self.distributed else 1
        self.P_r, self.P_c = process_grid(self.world)
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
        P = self.world
        self.J_loc = [j for j in range(self.NB) if snake_owner(j, P) == self.rank]     # my block columns of K / rows of W
        self.Aloc = None
        self.panels = {}         # k -> ((NB - k + 1) nb x nb): L[k nb:, k-th block column] followed by inv(L_kk)
        self.X = None            # my block rows of W = inv(L): (len(J_loc) nb) x npad, zero right of the staircase
        self.eta = None
        self.stats = {}
        self.bytes_received = 0
        self._pts_dev = None
        self._Abuf = self._Lflat = self._Xbuf = self._Tbuf = None

    # ---- helpers -------------------------------------------------------------------------------------------------
    def _gidx(self, blocks):
        nb = self.nb
        return numpy.concatenate([numpy.arange(b * nb, (b + 1) * nb) for b in blocks]).astype(numpy.int32) \
            if len(blocks) else numpy.zeros(0, dtype=numpy.int32)

    def _padded_points(self):
        if self._pts_dev is None:
            pad = numpy.zeros((self.npad, self.d))
            pad[:self.n] = self.points
            self._pts_dev = self.ops.from_host(pad)
        return self._pts_dev

    def _bcast(self, t, src):
        if self.world > 1:
            self.dist.broadcast(t, src=src)
            if self.rank != src:
                self.bytes_received += t.numel() * 8

    def _allreduce(self, t):
        if self.world > 1:
            self.dist.all_reduce(t)

    def _nvalid(self, k):
        return max(0, min(self.nb, self.n - k * self.nb))

    # ---- generation in block-cyclic layout ---------------------------------------------------------------------------
    def generate(self, eta):
        """Aloc (npad x len(J_loc) nb) <- my block columns of K + eta I (identity padding)."""
        ops = self.ops
        pts = self._padded_points()
        cg = self._gidx(self.J_loc)
        rg = numpy.arange(self.npad, dtype=numpy.int32)
        if self._Abuf is None:
            self._Abuf = ops.empty((self.npad, len(cg)))      # kept across factorisations (no allocator traffic per call)
        self.Aloc = self._Abuf
        if len(cg):
            ops.generate(pts, ops.from_host(self._padded_host()[cg]), ops.from_host(rg), ops.from_host(cg), self.n, self.scale,
                         self.nu, eta, self.Aloc)

    def _padded_host(self):
        pad = numpy.zeros((self.npad, self.d))
        pad[:self.n] = self.points
        return pad

    # ---- factorisation ---------------------------------------------------------------------------------------------
    def factor(self, eta):
        """Generates K + eta I in place and factors it; the panels stay replicated on every rank. Raises
        numpy.linalg.LinAlgError (on every rank) if the matrix is not positive definite."""
        ops, nb, NB, P = self.ops, self.nb, self.NB, self.world
        ops.synchronize()
        t0 = time.perf_counter()
        self.generate(eta)
        A = self.Aloc
        self.panels, self.X = {}, None
        self.bytes_received = 0
        bad = ops.zeros(1)
        infos = []
        lcol = {j: q for q, j in enumerate(self.J_loc)}
        side = ops.new_stream()
        # ONE buffer for the whole replicated factor (panel k: (NB - k + 1) nb rows), allocated once: a per-panel
        # allocation inside the loop would go through cudaMalloc - a device-wide synchronisation - and stall the look-ahead
        offs = numpy.concatenate([[0], numpy.cumsum([(NB - k + 1) * nb for k in range(NB)])])
        if self._Lflat is None:
            self._Lflat = ops.empty((int(offs[-1]), nb))
        Lflat = self._Lflat
        comm = []                    # (event before, event after) of every broadcast, on the side stream

        def apply_panel(k, j):
            """block column j (mine) -= panel_k rows >= j times (the j-th block of panel_k)^T"""
            msg = self.panels[k]
            q = lcol[j]
            C = A[j * nb:, q * nb:(q + 1) * nb]
            ops.gemm(C, msg[(j - k) * nb:(NB - k) * nb], msg[(j - k) * nb:(j - k + 1) * nb], -1.0, 1.0)

        def prepare_panel(k):
            """on the side stream: the owner factors the diagonal block and forms the panel; everybody joins the broadcast"""
            owner = snake_owner(k, P)
            msg = Lflat[int(offs[k]):int(offs[k + 1])]
            if self.rank == owner:
                q = lcol[k]
                col = A[k * nb:, q * nb:(q + 1) * nb]
                D = msg[:nb]
                D.copy_(col[:nb])
                Linv = msg[(NB - k) * nb:]
                infos.append((k, ops.potrf_inv(D, self._nvalid(k), Linv)))
                if NB - k > 1:
                    ops.gemm(msg[nb:(NB - k) * nb], col[nb:], Linv, 1.0, 0.0)      # L_ik = A_ik inv(L_kk)^T
            e0 = ops.event(timing=True) if self.world > 1 else None
            self._bcast(msg, owner)
            if e0 is not None:
                comm.append((e0, ops.event(timing=True)))
            self.panels[k] = msg

        main_done = None           # event: trailing update of the previous step enqueued on the main stream
        generated = ops.event()      # recorded on the MAIN stream (before the side stream becomes current)
        with ops.use(side):
            ops.wait(side, generated)        # the side stream starts after the generation
            prepare_panel(0)
            ready = ops.event()
        for k in range(NB):
            ops.wait(None, ready)                                   # main stream: panel k has arrived
            if k + 1 < NB:
                with ops.use(side):
                    if main_done is not None:
                        ops.wait(side, main_done)                   # column k+1 has all its updates from panels < k
                    if (k + 1) in lcol:
                        apply_panel(k, k + 1)
                    prepare_panel(k + 1)
                    ready = ops.event()
            for j in self.J_loc:
                if j > k + 1:
                    apply_panel(k, j)
            main_done = ops.event()
        for k, info in infos:
            bad += (info.to(bad.dtype) > 0) * float(1 + k) * 1e-6 + (info.to(bad.dtype) > 0)    # count + sum of (1 + block index) / 1e6
        self._allreduce(bad)
        ops.synchronize()
        self.stats = {'factor_s': time.perf_counter() - t0, 'bytes_received': self.bytes_received,
                      'comm_s': sum(ops.elapsed_s(a, b) for a, b in comm) if comm else 0.0}
        nbad = float(ops.to_host(bad)[0])
        if nbad != 0.0:
            raise numpy.linalg.LinAlgError('K + eta*I (eta=%g) is not positive definite (distributed potrf: %d diagonal blocks '
                                           'with a non-positive pivot, mean block index %.1f of %d).'
                                           % (eta, int(nbad), (nbad - int(nbad)) * 1e6 / max(int(nbad), 1) - 1, NB))
        self.Aloc = None                 # (the buffer itself stays in self._Abuf for the next factorisation)
        self.eta = float(eta)

    def logdet(self):
        """log det (K + eta I) = 2 sum log diag(L) from the replicated diagonal blocks (local, identical on every rank)."""
        ops = self.ops
        acc = ops.zeros(self.NB)
        for k in range(self.NB):
            nv = self._nvalid(k)
            if nv > 0:
                ops.logdet_chol(self.panels[k][:self.nb], nv, acc[k:k + 1])
        return float(ops.to_host(acc).sum())

    # ---- solves on the replicated factor (local; identical on every rank) ----------------------------------------------
    def solve_dev(self, Bdev):
        """(K + eta I)^-1 B for a device block (npad x p, p <= 16), in place: block substitution with the kept panels."""
        ops, nb, NB = self.ops, self.nb, self.NB
        p = Bdev.shape[1]
        t = ops.empty((nb, p))
        for k in range(NB):                                           # L y = b
            msg = self.panels[k]
            Linv = msg[(NB - k) * nb:]
            bk = Bdev[k * nb:(k + 1) * nb]
            ops.rect_apply(Linv, bk, t)
            bk.copy_(t)
            if k + 1 < NB:
                ops.rect_apply(msg[nb:(NB - k) * nb], bk, Bdev[(k + 1) * nb:], alpha=-1.0, beta=1.0)
        for k in range(NB - 1, -1, -1):                               # L^T x = y
            msg = self.panels[k]
            Linv = msg[(NB - k) * nb:]
            bk = Bdev[k * nb:(k + 1) * nb]
            if k + 1 < NB:
                ops.rect_apply_t(msg[nb:(NB - k) * nb], Bdev[(k + 1) * nb:], bk, alpha=-1.0, beta=1.0)
            ops.rect_apply_t(Linv, bk, t)
            bk.copy_(t)
        return Bdev

    def solve(self, R):
        """(K + eta I)^-1 R for a host array R (n,) or (n, p); returns a host array (identical on every rank)."""
        R = numpy.asarray(R, dtype=numpy.float64)
        vec = (R.ndim == 1)
        R2 = R.reshape(self.n, -1)
        out = numpy.empty_like(R2)
        for c0 in range(0, R2.shape[1], 16):
            blk = R2[:, c0:c0 + 16]
            Bp = numpy.zeros((self.npad, blk.shape[1]))
            Bp[:self.n] = blk
            out[:, c0:c0 + 16] = self.ops.to_host(self.solve_dev(self.ops.from_host(Bp)))[:self.n]
        return out[:, 0] if vec else out

    def solve_with_rows(self, Rdev):
        """(K + eta I)^-1 R = W^T (W R) from the rows of W = inv(L) this rank holds: two passes over the local slab and one
        all-reduce of the (npad x p) result - instead of 2 NB latency-bound substitution steps on the replicated panels."""
        ops, nb = self.ops, self.nb
        X = self.inverse_rows()
        nq = len(self.J_loc)
        p = Rdev.shape[1]
        S = ops.zeros((self.npad, p))
        if nq:
            Y = ops.empty((nq * nb, p))
            ops.rect_apply(X[:nq * nb], Rdev, Y)
            ops.rect_apply_t(X[:nq * nb], Y, S)
        self._allreduce(S)
        return S

    # ---- rows of inv(L) ----------------------------------------------------------------------------------------------
    def inverse_rows(self):
        """X <- my block rows of W = inv(L) (block back-substitution from the right on the replicated factor)."""
        if self.X is not None:
            return self.X
        ops, nb, NB = self.ops, self.nb, self.NB
        I = self.J_loc
        nq = len(I)
        if self._Tbuf is None:
            self._Tbuf = ops.empty((max(nq, 1) * nb, nb))
        if nq and self._Abuf is not None:
            # my block columns of K are dead once the factor sits in the panels: the rows of W take over that storage
            # (npad x nq nb and nq nb x npad have the same number of entries)
            X = self._Abuf.reshape(-1)[:nq * nb * self.npad].view(nq * nb, self.npad)
        else:
            if self._Xbuf is None:
                self._Xbuf = ops.empty((max(nq, 1) * nb, self.npad))
            X = self._Xbuf
        T = self._Tbuf
        X.zero_()
        tiles = nb // 128
        for k in range(NB - 1, -1, -1):
            msg = self.panels[k]
            Linv = msg[(NB - k) * nb:]
            q0 = next((q for q, i in enumerate(I) if i >= k), nq)        # first local row with i_q >= k
            if q0 < nq and I[q0] == k:
                X[q0 * nb:(q0 + 1) * nb, k * nb:(k + 1) * nb].copy_(Linv)
                q0 += 1
            if q0 >= nq:
                continue
            rows = slice(q0 * nb, nq * nb)
            # T = X[rows, columns > k] L[rows > k, k]; row block q only has columns <= i_q
            kend = ops.int_tensor([(I[q] - k) * nb for q in range(q0, nq) for _ in range(tiles)])
            ops.gemm(T[rows], X[rows, (k + 1) * nb:], msg[nb:(NB - k) * nb], 1.0, 0.0, at=0, bt=1, kend=kend)
            ops.gemm(X[rows, k * nb:(k + 1) * nb], T[rows], Linv, -1.0, 0.0, at=0, bt=1)
        self.X = X
        return X

    def inverse_traces(self, with_dk=True):
        """(tr Kn^-1, tr Kn^-1 dK/drho) - all-reduced, identical on every rank."""
        ops, nb, NB, npad = self.ops, self.nb, self.NB, self.npad
        X = self.inverse_rows()
        I = self.J_loc
        nq = len(I)
        acc = ops.zeros(2)
        if nq:
            ops.pair_dot(X[:nq * nb], X[:nq * nb], nq * nb, 1.0, acc[0:1])
        if with_dk and nq:
            pts = self._padded_points()
            host = self._padded_host()
            C = ops.empty((npad, nb))
            D = ops.empty((npad, nb))
            tiles = nb // 128
            for j in range(NB):
                q0 = next((q for q, i in enumerate(I) if i >= j), nq)
                if q0 >= nq:
                    break
                rows = slice(q0 * nb, nq * nb)
                M = npad - j * nb
                # C[J', J] = sum over my rows i >= j' of W[i, J']^T W[i, J]   (J' >= J): TN product with a staircase k start
                kbeg = ops.int_tensor([(next((q for q in range(q0, nq) if I[q] >= jp), nq) - q0) * nb
                                       for jp in range(j, NB) for _ in range(tiles)])
                ops.gemm(C[:M], X[rows, j * nb:], X[rows, j * nb:(j + 1) * nb], 1.0, 0.0, at=1, bt=1, kbeg=kbeg)
                cg = self._gidx([j])
                rg = numpy.arange(j * nb, npad, dtype=numpy.int32)
                ops.generate_dk(pts[j * nb:], ops.from_host(host[cg]), ops.from_host(rg), ops.from_host(cg), self.n, self.scale,
                                self.nu, D[:M])
                ops.pair_dot(C[:M], D[:M], nb, 2.0, acc[1:2])          # the diagonal block once, the blocks below twice
        self._allreduce(acc)
        out = ops.to_host(acc)
        # the identity padding block of L inverts to itself: its npad - n unit entries are not part of tr Kn^-1
        return float(out[0]) - (npad - self.n), float(out[1])

    # ---- likelihood ----------------------------------------------------------------------------------------------------
    def _fused_out(self, z, X, with_gradient):
        """out[] in the layout of gp_loglik_dense (include/gpgp.h) for the host algebra of _likelihood/_fused.py"""
        ops = self.ops
        n, m = X.shape
        p = m + 1
        R = numpy.zeros((self.npad, p))
        R[:n, :m] = X
        R[:n, m] = z
        Rd = ops.from_host(R)
        ops.synchronize()
        if with_gradient:
            t0 = time.perf_counter()
            self.inverse_rows()
            ops.synchronize()
            self.stats['inverse_rows_s'] = time.perf_counter() - t0
        t0 = time.perf_counter()
        S = self.solve_with_rows(Rd) if with_gradient else self.solve_dev(Rd.clone())
        Sh, Rh = ops.to_host(S), R
        out = numpy.zeros(8 + 4 * p * p)
        out[0] = self.logdet()
        out[8:8 + p * p] = (Rh.T @ Sh).ravel()
        out[8 + p * p:8 + 2 * p * p] = (Sh.T @ Sh).ravel()
        if with_gradient:
            V = ops.zeros((self.npad, p))
            ops.dk_apply(ops.from_host(self.points), n, self.scale, self.nu, S, V)
            out[8 + 2 * p * p:8 + 3 * p * p] = (Sh.T @ ops.to_host(V)).ravel()
        ops.synchronize()
        self.stats['solve_s'] = time.perf_counter() - t0
        if with_gradient:
            t0 = time.perf_counter()
            out[1], out[3] = self.inverse_traces(with_dk=True)
            ops.synchronize()
            self.stats['traces_s'] = time.perf_counter() - t0
        return out

    def profile_log_likelihood(self, z, X, eta):
        """l^(sigma_hat, eta) exactly as ProfileLikelihood.log_likelihood (reference _profile_likelihood.py:38-85)
        evaluated at sigma_hat^2 = z^T M z / (n - m); returns (l^, sigma_hat)."""
        from ._likelihood._fused import FusedQuantities
        self.factor(eta)
        n, m = X.shape
        q = FusedQuantities(self._fused_out(z, X, False), n, m, float(eta), 0)
        sigma2 = q.zMz / (n - m)
        lp = -0.5 * (n - m) * numpy.log(sigma2) - 0.5 * q.logdet_Kn - 0.5 * numpy.log(numpy.linalg.det(q.B)) - 0.5 * (n - m)
        return float(lp), float(numpy.sqrt(sigma2))

    def profile_log_likelihood_and_gradient(self, z, X, eta):
        """(l^, d l^/d eta, d l^/d rho) at one (rho, eta) - the unit of BASELINE.json's metric - for n beyond one GPU
        (reference formulas: _profile_likelihood.py:38-132; d/d rho: SURVEY 8a A10)."""
        from ._likelihood._fused import FusedQuantities
        from ._likelihood._profile_likelihood import ProfileLikelihood
        self.factor(eta)
        n, m = X.shape
        q = FusedQuantities(self._fused_out(z, X, True), n, m, float(eta), 7)
        return ProfileLikelihood._gradient_from(q, X.shape, True)


@contextlib.contextmanager
def _null():
    yield
