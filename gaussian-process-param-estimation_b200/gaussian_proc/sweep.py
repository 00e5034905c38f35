"""
Hyper-parameter grid sweep (BASELINE.json configs[2]; the workload of the reference's legacy
examples/FindOptimalCovarianceParameters.py:632-702, a multiprocessing.Pool over independent cells): profile
log-likelihood and its derivatives with respect to eta and the correlation scale rho on a (rho x eta) grid.

Cells are grouped by rho; each rank owns a contiguous range of rho values, generates K(rho) once on its GPU and
loops over eta with one Cholesky each. No data-path collective; the per-cell results are all-gathered at the end.
"""

import numpy

from . import _distributed as gpd

__all__ = ['likelihood_grid']


class _GpuRowEvaluator(object):
    """All eta cells of one rho: K(rho) is generated once; `concurrency` operators (each with its own scratch, sharing
    the read-only K) evaluate cells on their own CUDA streams so that the latency-bound phases of one evaluation
    (diagonal blocks, small recursion levels) are filled by the GEMMs of another."""

    def __init__(self, points, z, X, nu, concurrency):
        from . import _device as dev
        self.torch = dev.require_cuda()
        self.points = numpy.ascontiguousarray(points, dtype=float)
        self.z, self.X, self.nu = z, X, float(nu)
        self.concurrency = max(1, int(concurrency))
        self.streams = [self.torch.cuda.Stream() for _ in range(self.concurrency)]
        self.ops = None
        self.n_ops = -1

    def row(self, rho, etas):
        from .generate_correlation.generate_correlation import generate_dense_correlation
        from ._mixed_correlation import MixedCorrelation
        from ._likelihood import ProfileLikelihood
        torch = self.torch
        K = generate_dense_correlation(self.points, numpy.repeat(float(rho), self.points.shape[1]), self.nu)
        if self.ops is None or self.n_ops != K.n:
            self.ops = [MixedCorrelation(K) for _ in range(self.concurrency)]
            self.n_ops = K.n
        else:
            for op in self.ops:      # reuse the scratch buffers, swap the matrix
                op.K = K
                op.engine.K = K
                op.engine.invalidate()
        cur = torch.cuda.current_stream()
        for s in self.streams:
            s.wait_stream(cur)
        out = numpy.empty((len(etas), 3))
        pending = []
        for j, eta in enumerate(etas):
            slot = j % self.concurrency
            if len(pending) >= self.concurrency:
                jj, fin = pending.pop(0)
                out[jj] = fin()
            with torch.cuda.stream(self.streams[slot]):
                fin = ProfileLikelihood.log_likelihood_and_gradient_async(self.z, self.X, self.ops[slot], eta)
            pending.append((j, fin))
        for jj, fin in pending:
            out[jj] = fin()
        for s in self.streams:
            cur.wait_stream(s)
        return out


class _GpuEigenRowEvaluator(object):
    """All eta cells of one rho through ONE eigendecomposition of K(rho) (imate_method='eigenvalue', the reference's
    default, likelihood.py:41; _dense.EigenEngine): every cell after the eigensolve is O(n^2 p). Pays off for long eta
    rows (about 30 cells at n = 8000); the eigensolver itself is the cuSOLVER library routine."""

    def __init__(self, points, z, X, nu):
        from . import _device as dev
        dev.require_cuda()
        self.points = numpy.ascontiguousarray(points, dtype=float)
        self.z, self.X, self.nu = z, X, float(nu)

    def row(self, rho, etas):
        from .generate_correlation.generate_correlation import generate_dense_correlation
        from ._mixed_correlation import MixedCorrelation
        from ._likelihood import ProfileLikelihood
        K = generate_dense_correlation(self.points, numpy.repeat(float(rho), self.points.shape[1]), self.nu)
        Km = MixedCorrelation(K, imate_method='eigenvalue')
        return numpy.array([ProfileLikelihood.log_likelihood_and_gradient(self.z, self.X, Km, eta) for eta in etas])


class _GpuSparseRowEvaluator(object):
    """All eta cells of one rho for the kernel-threshold sparse correlation: the CSR matrix and its row-blocked
    operator are generated once per rho; every eta is one stochastic evaluation (batched SLQ + CG, _sparse.py)."""

    def __init__(self, points, z, X, nu, density, imate_options):
        from . import _device as dev
        dev.require_cuda()
        self.points = numpy.ascontiguousarray(points, dtype=float)
        self.z, self.X, self.nu, self.density = z, X, float(nu), float(density)
        self.options = dict(imate_options or {})
        self.options.setdefault('eager_rhs_basis', True)      # every row asks several eta of one operator

    def row(self, rho, etas):
        from ._sparse import generate_sparse_correlation
        from ._mixed_correlation import MixedCorrelation
        from ._likelihood import ProfileLikelihood
        K = generate_sparse_correlation(self.points, numpy.repeat(float(rho), self.points.shape[1]), self.nu, self.density,
                                        device=True, with_derivative=True)
        Km = MixedCorrelation(K, imate_method='slq', imate_options=self.options)
        return numpy.array([ProfileLikelihood.log_likelihood_and_gradient(self.z, self.X, Km, eta) for eta in etas])


def likelihood_grid(points, z, X, nu, rhos, etas, evaluate=None, concurrency=4, sparse=False, density=1e-3,
                    imate_options=None, method='cholesky'):
    """Returns an array (len(rhos), len(etas), 3) with [l^(sigma_hat, eta), d l^/d eta, d l^/d rho] per cell, identical
    on every rank. `evaluate(rho, eta)` may be injected (tests); by default it is the fused GPU evaluator with
    `concurrency` cells in flight per GPU (dense), or, with ``sparse``, the stochastic evaluator on the kernel-threshold
    correlation of the given ``density`` (``imate_options``: estimator settings, see _sparse.DEFAULTS).
    ``method='eigenvalue'`` (dense): one eigendecomposition per rho instead of one Cholesky per cell."""
    rhos = numpy.asarray(rhos, dtype=float)
    etas = numpy.asarray(etas, dtype=float)
    rank, world = gpd.rank_world()
    begin, end = gpd.partition_cells(len(rhos), world, rank)
    if evaluate is not None:
        rows = None
    elif sparse:
        rows = _GpuSparseRowEvaluator(points, z, X, nu, density, imate_options)
    elif method == 'eigenvalue':
        rows = _GpuEigenRowEvaluator(points, z, X, nu)
    else:
        rows = _GpuRowEvaluator(points, z, X, nu, concurrency)
    local = numpy.empty(((end - begin) * len(etas), 5))
    k = 0
    for i in range(begin, end):
        vals = rows.row(rhos[i], etas) if rows is not None else [evaluate(rhos[i], eta) for eta in etas]
        for j in range(len(etas)):
            local[k, :2] = (i, j)
            local[k, 2:] = vals[j]
            k += 1
    gathered = gpd.allgather_rows(local)
    out = numpy.full((len(rhos), len(etas), 3), numpy.nan)
    out[gathered[:, 0].astype(int), gathered[:, 1].astype(int)] = gathered[:, 2:]
    return out
