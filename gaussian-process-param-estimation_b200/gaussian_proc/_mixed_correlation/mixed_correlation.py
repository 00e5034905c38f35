"""
MixedCorrelation -- the operator K + eta*I behind the six duck-typed methods the likelihood code uses
(reference: gaussian_proc/_mixed_correlation/mixed_correlation.py:34-335 and _linear_solver.py:24-73).

Dense K: one blocked FP64 Cholesky per eta on the GPU (cached), reused by logdet / solve / traceinv, instead of the
reference's fresh dposv per solve. Sparse K (CSR): SpMM + stochastic Lanczos quadrature / CG (see _sparse.py).
"""

import numpy
import scipy.sparse

from .._dense import DeviceCorrelation, DenseEngine

__all__ = ['MixedCorrelation']

_DENSE_METHODS = ('cholesky', 'eigenvalue')
_SPARSE_METHODS = ('slq', 'hutchinson')


class MixedCorrelation(object):
    """
    Same constructor and methods as the reference class (mixed_correlation.py:34-35). Differences, all documented
    in DESIGN.md: ``imate_method`` 'cholesky' (dense) and 'slq' / 'hutchinson' (sparse) are implemented natively;
    'eigenvalue' (the method the reference's Likelihood hard-codes, likelihood.py:41) computes all eigenvalues of K once
    in __init__ like mixed_correlation.py:76-79 -- with this library's own Householder tridiagonalisation + bisection
    (csrc/gp_eig.cu, no cuSOLVER) -- after which logdet / traceinv / trace are O(n) reductions over lambda + eta;
    ``solve`` uses the native Cholesky engine, and the fused likelihood evaluation runs on the tridiagonal form
    (_dense.EigenEngine): one reduction per matrix, O(n p) per eta for l^ and d l^/d eta.
    ``interpolate=True``: tr (K + eta I)^-1 is interpolated in eta from evaluations at ``interpolant_points``
    (_interpolate_traceinv.py, the role of imate.InterpolateTraceInv; rational polynomial scheme, parity unpinned).
    """

    def __init__(self, K, interpolate=False, interpolant_points=None, imate_method='cholesky', imate_options={}):
        if interpolate:
            if interpolant_points is None:
                raise TypeError('When "interpolate" is set to "True", the "interpolant_points" cannot be None.')
        self.interpolate = interpolate
        self.interpolant_points = interpolant_points
        self.imate_method = imate_method
        self.imate_options = dict(imate_options)
        self.interpolate_traceinv = None      # built on first use (evaluates traceinv at the interpolant points)
        self.sparse = False

        if scipy.sparse.issparse(K) or type(K).__name__ in ('DeviceCSR', 'DeviceRowBlocks'):
            from .._sparse import SparseEngine
            if imate_method not in _SPARSE_METHODS:
                raise ValueError('For a sparse K, existing methods are "slq" and "hutchinson".')
            self.sparse = True
            # imate_options['probe_split']=True: the probes of every stochastic estimate are split over the ranks of the
            # initialised torch.distributed group; the estimates are then identical on every rank
            probe_range = self.imate_options.pop('probe_range', None)
            if self.imate_options.pop('probe_split', False):
                from .._distributed import rank_world
                probe_range = rank_world()
            if self.imate_options.pop('row_slabs', False):
                # ONE evaluation on all GPUs of the group: this rank holds a slab of the operator's rows; halo rows and
                # reductions cross NVLink inside the kernels (gaussian_proc/_slab.py)
                from .._slab import SlabSparseEngine
                self.engine = SlabSparseEngine(K, imate_method, self.imate_options)
            else:
                self.engine = SparseEngine(K, imate_method, self.imate_options, probe_range=probe_range)
            self.K = self.engine.K
        else:
            if imate_method not in _DENSE_METHODS:
                raise ValueError('Existing methods are "eigenvalue", "cholesky", "hutchinson", and "slq".')
            if not isinstance(K, DeviceCorrelation):
                K = DeviceCorrelation.from_numpy(K)
            self.K = K
            self.engine = DenseEngine(K)
            self.K_eigenvalues = None
            self._eigen = None
            if imate_method == 'eigenvalue':
                # ONE tridiagonalisation + all eigenvalues (the reference's eigh(K, eigvals_only=True), :78-79)
                from .._dense import EigenEngine
                self._eigen = EigenEngine(K)
                self.K_eigenvalues = self._eigen.lam

    def eigen_engine(self):
        """The eigendecomposition of K behind imate_method='eigenvalue' (_dense.EigenEngine): the fused likelihood
        evaluation costs O(n^2 p) per eta on it."""
        if self._eigen is None:
            from .._dense import EigenEngine
            self._eigen = EigenEngine(self.K)
        return self._eigen

    # -- extension: generator parameters for d/d(correlation_scale) when K came in as a plain array
    def set_kernel(self, points, correlation_scale, nu):
        from .. import _device as dev
        torch = dev.require_cuda()
        points = numpy.ascontiguousarray(points, dtype=numpy.float64)
        if numpy.isscalar(correlation_scale):
            correlation_scale = numpy.repeat(float(correlation_scale), points.shape[1])
        self.K.points = torch.from_numpy(points).cuda()
        self.K.correlation_scale = dev.host_f64(correlation_scale)
        self.K.nu = float(nu)

    def get_matrix_size(self):
        """mixed_correlation.py:85-90"""
        return self.K.shape[0]

    def trace(self, eta, exponent=1):
        """tr (K + eta I)^exponent for exponent 0, 1, 2 (the exact branches of mixed_correlation.py:105-123)."""
        n = self.K.shape[0]
        if exponent == 0:
            return float(n)
        trK, trK2 = self.engine.trace_K()
        if exponent == 1:
            return trK + eta * n if eta != 0 else trK
        if exponent == 2:
            return trK2 if eta == 0 else trK2 + 2.0 * eta * trK + eta ** 2 * n
        if not self.sparse and self.imate_method == 'eigenvalue':
            return float(((self.K_eigenvalues + eta) ** exponent).sum().item())           # :127-136
        raise ValueError('Existing methods are "exact", "eigenvalue", and "slq".')

    def traceinv(self, eta, exponent=1):
        """mixed_correlation.py:155-215"""
        if self.interpolate and exponent == 1:                                          # :167-170
            if self.interpolate_traceinv is None:
                from ._interpolate_traceinv import InterpolateTraceInv
                self.interpolate_traceinv = InterpolateTraceInv(
                    lambda t: self._traceinv_direct(t, 1), self.K.shape[0], self.interpolant_points,
                    self.imate_options.get('interpolation_method', 'RPF'))
            return self.interpolate_traceinv.interpolate(eta)
        return self._traceinv_direct(eta, exponent)

    def _traceinv_direct(self, eta, exponent=1):
        if not self.sparse and self.imate_method == 'eigenvalue':
            if exponent in (1, 2):
                return self._eigen.reductions(eta)[exponent]                              # :172-181
            lam = self.K_eigenvalues.cpu().numpy()                  # (other exponents: a host sum over the n eigenvalues)
            return float(numpy.sum((lam + eta) ** (-float(exponent))))
        return self.engine.traceinv(eta, exponent)

    def logdet(self, eta, exponent=1):
        """mixed_correlation.py:221-274; logdet(Kn^p) = p logdet(Kn)."""
        if not self.sparse and self.imate_method == 'eigenvalue':
            logdet, _, _, bad = self._eigen.reductions(eta)                               # :239-248
            if bad:
                raise numpy.linalg.LinAlgError('K + eta*I (eta=%g) is not positive definite.' % eta)
            return exponent * logdet
        return exponent * self.engine.logdet(eta)

    def solve(self, eta, Y):
        """mixed_correlation.py:280-299 -> linear_solver(Kn, Y, 'sym_pos')."""
        return self.engine.solve(eta, Y)

    def dot(self, eta, x, exponent=1):
        """mixed_correlation.py:305-335. exponent must be a non-negative int. For exponent 2 the reference returns
        2*(K x + eta x); this build returns the intended (K + eta I)^2 x."""
        if not isinstance(exponent, int):
            raise ValueError('"exponent" should be an integer.')
        elif exponent < 0:
            raise ValueError('"exponent" should be a non-negative integer.')
        x = numpy.asarray(x, dtype=float)
        if exponent == 0:
            return numpy.zeros_like(x)
        y = x
        for _ in range(exponent):
            y = self.engine.matmul(y) + (eta * y if eta != 0 else 0.0)
        return y
