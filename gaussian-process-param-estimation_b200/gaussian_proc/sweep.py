"""
Hyper-parameter grid sweep (BASELINE.json configs[2]; the workload of the reference's legacy
examples/FindOptimalCovarianceParameters.py:632-702, a multiprocessing.Pool over independent cells): profile
log-likelihood and its derivatives with respect to eta and the correlation scale rho on a (rho x eta) grid.

Cells are grouped by rho; each rank owns a contiguous range of rho values, generates K(rho) once on its GPU and
loops over eta with one Cholesky each. No data-path collective; the per-cell results are all-gathered at the end.
"""

import numpy

from . import _distributed as gpd

__all__ = ['likelihood_grid', 'profile_likelihood_surface']


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
    """All eta cells of one rho through ONE tridiagonalisation of K(rho) (imate_method='eigenvalue', the reference's
    default, likelihood.py:41; _dense.EigenEngine, csrc/gp_eig.cu): every cell after it is O(n p) for l^ and d l^/d eta.
    Pays off for long eta rows / root finds. d l^/d rho is not a function of the spectrum: with ``with_rho`` those cells
    add one Cholesky evaluation each (then the plain Cholesky rows are the faster choice)."""

    def __init__(self, points, z, X, nu, with_rho=True):
        from . import _device as dev
        dev.require_cuda()
        self.points = numpy.ascontiguousarray(points, dtype=float)
        self.z, self.X, self.nu, self.with_rho = z, X, float(nu), bool(with_rho)

    def row(self, rho, etas):
        from .generate_correlation.generate_correlation import generate_dense_correlation
        from ._mixed_correlation import MixedCorrelation
        from ._likelihood import ProfileLikelihood
        K = generate_dense_correlation(self.points, numpy.repeat(float(rho), self.points.shape[1]), self.nu)
        Km = MixedCorrelation(K, imate_method='eigenvalue')
        rows = [ProfileLikelihood.log_likelihood_and_gradient(self.z, self.X, Km, eta, with_rho=self.with_rho) for eta in etas]
        return numpy.array([[r[0], r[1], numpy.nan if r[2] is None else r[2]] for r in rows], dtype=float)


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
        from ._sparse import generate_sparse_operator
        from ._mixed_correlation import MixedCorrelation
        from ._likelihood import ProfileLikelihood
        K = generate_sparse_operator(self.points, numpy.repeat(float(rho), self.points.shape[1]), self.nu, self.density,
                                     with_derivative=True)
        Km = MixedCorrelation(K, imate_method='slq', imate_options=self.options)
        return numpy.array([ProfileLikelihood.log_likelihood_and_gradient(self.z, self.X, Km, eta) for eta in etas])


class _Checkpoint(object):
    """Restartable sweeps: every finished row (one correlation matrix, all its cells) is written to
    `<dir>/row_<key>.npy` by the rank that computed it (write to a temporary name, then rename: a killed job never leaves
    a torn file). A later call with the same directory - any number of ranks - loads the finished rows and computes only
    the missing ones. The directory must be visible to every rank (one node: any local path)."""

    def __init__(self, directory, signature):
        import os
        self.dir = directory
        self.os = os
        if directory is not None:
            os.makedirs(directory, exist_ok=True)
            sig = os.path.join(directory, 'signature.txt')
            if os.path.exists(sig):
                if open(sig).read() != signature:
                    raise ValueError('checkpoint directory %s belongs to a different sweep' % directory)
            else:
                tmp = sig + '.%d.tmp' % os.getpid()
                open(tmp, 'w').write(signature)
                os.replace(tmp, sig)

    def _path(self, key):
        return self.os.path.join(self.dir, 'row_%s.npy' % key)

    def load(self, key):
        if self.dir is None or not self.os.path.exists(self._path(key)):
            return None
        return numpy.load(self._path(key))

    def save(self, key, row):
        if self.dir is None:
            return
        tmp = self._path(key) + '.%d.tmp.npy' % self.os.getpid()
        numpy.save(tmp, numpy.asarray(row))
        self.os.replace(tmp, self._path(key))


def _signature(tag, *arrays):
    import hashlib
    h = hashlib.blake2b(digest_size=16)
    h.update(tag.encode())
    for a in arrays:
        h.update(numpy.ascontiguousarray(numpy.asarray(a, dtype=float)).tobytes())
    return h.hexdigest()


def _frozen(a):
    """a read-only copy: its device copy is then keyed once instead of being re-fingerprinted for every cell"""
    if a is None:
        return None
    b = numpy.array(a, dtype=float, copy=True)
    b.setflags(write=False)
    return b


def likelihood_grid(points, z, X, nu, rhos, etas, evaluate=None, concurrency=4, sparse=False, density=1e-3,
                    imate_options=None, method='cholesky', checkpoint=None, with_rho=True):
    """Returns an array (len(rhos), len(etas), 3) with [l^(sigma_hat, eta), d l^/d eta, d l^/d rho] per cell, identical
    on every rank. `evaluate(rho, eta)` may be injected (tests); by default it is the fused GPU evaluator with
    `concurrency` cells in flight per GPU (dense), or, with ``sparse``, the stochastic evaluator on the kernel-threshold
    correlation of the given ``density`` (``imate_options``: estimator settings, see _sparse.DEFAULTS).
    ``method='eigenvalue'`` (dense): one tridiagonalisation per rho instead of one Cholesky per cell (``with_rho=False``
    skips d l^/d rho, NaN in the result, which the spectrum alone cannot give).
    ``nu`` may be a sequence: the smoothness becomes a third sweep axis (the legacy (rho, nu) workload,
    examples/FindOptimalCovarianceParameters.py:665-690) and the result has shape (len(nu), len(rhos), len(etas), 3); the
    (nu, rho) rows are what the ranks share out.
    ``checkpoint``: a directory; finished rows are kept there and a restarted sweep computes only the missing ones."""
    rhos = numpy.asarray(rhos, dtype=float)
    etas = numpy.asarray(etas, dtype=float)
    nus = numpy.atleast_1d(numpy.asarray(nu, dtype=float))
    z, X = _frozen(z), _frozen(X)
    rows_all = [(a, i) for a in range(len(nus)) for i in range(len(rhos))]           # one correlation matrix per row
    rank, world = gpd.rank_world()
    ck = _Checkpoint(checkpoint, _signature('grid|%s|%s|%g' % (method, sparse, density), nus, rhos, etas,
                                            numpy.asarray(points).shape if points is not None else [0]))
    done = {}
    if checkpoint is not None:
        for (a, i) in rows_all:
            row = ck.load('%d_%d' % (a, i))
            if row is not None and row.shape == (len(etas), 3):
                done[(a, i)] = row
    todo = [r for r in rows_all if r not in done]
    begin, end = gpd.partition_cells(len(todo), world, rank)
    evaluators = {}

    def evaluator(a):
        if a not in evaluators:
            if sparse:
                evaluators[a] = _GpuSparseRowEvaluator(points, z, X, nus[a], density, imate_options)
            elif method == 'eigenvalue':
                evaluators[a] = _GpuEigenRowEvaluator(points, z, X, nus[a], with_rho)
            else:
                evaluators[a] = _GpuRowEvaluator(points, z, X, nus[a], concurrency)
        return evaluators[a]

    local = numpy.empty(((end - begin) * len(etas), 6))
    k = 0
    for (a, i) in todo[begin:end]:
        if evaluate is not None:
            vals = numpy.array([evaluate(rhos[i], eta) if len(nus) == 1 and numpy.ndim(nu) == 0 else evaluate(nus[a], rhos[i], eta)
                                for eta in etas], dtype=float)
        else:
            vals = numpy.asarray(evaluator(a).row(rhos[i], etas), dtype=float)
        ck.save('%d_%d' % (a, i), vals)
        for j in range(len(etas)):
            local[k, :3] = (a, i, j)
            local[k, 3:] = vals[j]
            k += 1
    gathered = gpd.allgather_rows(local)
    out = numpy.full((len(nus), len(rhos), len(etas), 3), numpy.nan)
    for (a, i), row in done.items():
        out[a, i] = row
    if gathered.shape[0]:
        out[gathered[:, 0].astype(int), gathered[:, 1].astype(int), gathered[:, 2].astype(int)] = gathered[:, 3:]
    return out[0] if numpy.ndim(nu) == 0 else out


def profile_likelihood_surface(points, z, X, rhos, nus, interval_eta=(1e-3, 1e3), checkpoint=None, evaluate=None):
    """The reference's legacy sweep (examples/FindOptimalCovarianceParameters.py:632-702, a multiprocessing.Pool over the
    61 x 60 (rho, nu) grid of data/OptimalCovariance_WithoutPrior.pickle): for every (rho, nu) the profile likelihood
    maximised over eta by the root of d l^/d eta (ProfileLikelihood.find_log_likelihood_der1_zeros), i.e.
    Lp[i, j] = l^(sigma_hat, eta_hat; rho_i, nu_j). The (rho, nu) cells are shared out over the ranks (no data-path
    collective, one all-gather of the results); returns (Lp, eta_hat), each (len(rhos), len(nus)), identical on every rank.
    ``checkpoint``: restartable, one file per rho row."""
    import contextlib
    import io
    rhos = numpy.asarray(rhos, dtype=float)
    nus = numpy.asarray(nus, dtype=float)
    z, X = _frozen(z), _frozen(X)
    rank, world = gpd.rank_world()
    ck = _Checkpoint(checkpoint, _signature('surface', rhos, nus, interval_eta))
    done = {}
    if checkpoint is not None:
        for i in range(len(rhos)):
            row = ck.load(str(i))
            if row is not None and row.shape == (len(nus), 2):
                done[i] = row
    todo = [i for i in range(len(rhos)) if i not in done]
    begin, end = gpd.partition_cells(len(todo), world, rank)

    def cell(rho, nu_):
        if evaluate is not None:
            return evaluate(rho, nu_)
        from .generate_correlation.generate_correlation import generate_dense_correlation
        from ._mixed_correlation import MixedCorrelation
        from ._likelihood import ProfileLikelihood
        K = generate_dense_correlation(numpy.ascontiguousarray(points, dtype=float), numpy.repeat(float(rho), points.shape[1]),
                                       float(nu_))
        Km = MixedCorrelation(K)
        with contextlib.redirect_stdout(io.StringIO()):
            res = ProfileLikelihood.find_log_likelihood_der1_zeros(z, X, Km, list(interval_eta))
        return ProfileLikelihood.log_likelihood(z, X, Km, False, [res['sigma'], res['eta']]), res['eta']

    local = numpy.empty(((end - begin) * len(nus), 4))
    k = 0
    for i in todo[begin:end]:
        row = numpy.array([cell(rhos[i], nu_) for nu_ in nus], dtype=float)
        ck.save(str(i), row)
        for j in range(len(nus)):
            local[k] = (i, j, row[j, 0], row[j, 1])
            k += 1
    gathered = gpd.allgather_rows(local)
    Lp = numpy.full((len(rhos), len(nus)), numpy.nan)
    eta_hat = numpy.full((len(rhos), len(nus)), numpy.nan)
    for i, row in done.items():
        Lp[i], eta_hat[i] = row[:, 0], row[:, 1]
    if gathered.shape[0]:
        ii, jj = gathered[:, 0].astype(int), gathered[:, 1].astype(int)
        Lp[ii, jj], eta_hat[ii, jj] = gathered[:, 2], gathered[:, 3]
    return Lp, eta_hat
