"""
Device-resident dense correlation matrix and the FP64 factorisation engine built on libgpgp.

``DeviceCorrelation`` is what ``generate_correlation(..., device=True)`` returns and what ``MixedCorrelation`` keeps
internally: the padded (npad x npad, identity padding) matrix in HBM plus, when known, the generator parameters
(points, correlation_scale, nu) that let the d/d(correlation_scale) reductions re-evaluate dK on the fly.

``DenseEngine`` owns the scratch buffers for one matrix size and exposes the stream-ordered building blocks
(factor, logdet, solve, inverse traces, fused evaluation). One factorisation is cached per eta.
"""

import ctypes

import numpy

from . import _device as dev
from ._device import lib, check

__all__ = ['DeviceCorrelation', 'DenseEngine', 'EigenEngine', 'FLAG_TRACEINV', 'FLAG_INVERSE', 'FLAG_DRHO', 'FLAG_CUBIC']

FLAG_TRACEINV = 1
FLAG_INVERSE = 2
FLAG_DRHO = 4
FLAG_CUBIC = 8      # third moments T3 = R^T Kn^-3 R (Hessian, second eta-derivative)
MAX_RHS = 16


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


class DeviceCorrelation(object):
    """Padded dense correlation matrix in device memory (float64, row-major, leading dimension npad)."""

    def __init__(self, n, data, points=None, correlation_scale=None, nu=None):
        self.n = int(n)
        self.npad = dev.padded_size(n)
        self.data = data              # torch.float64 CUDA tensor (npad, npad)
        self.points = points          # torch.float64 CUDA tensor (n, d) or None
        self.correlation_scale = None if correlation_scale is None else dev.host_f64(correlation_scale)
        self.nu = nu

    @property
    def shape(self):
        return (self.n, self.n)

    @classmethod
    def from_numpy(cls, K):
        torch = dev.require_cuda()
        K = numpy.asarray(K, dtype=numpy.float64)
        if K.ndim != 2 or K.shape[0] != K.shape[1]:
            raise ValueError('K should be a square matrix.')
        n = K.shape[0]
        npad = dev.padded_size(n)
        data = torch.zeros((npad, npad), dtype=torch.float64, device='cuda')
        data[:n, :n].copy_(torch.from_numpy(numpy.ascontiguousarray(K)))
        if npad > n:
            idx = torch.arange(n, npad, device='cuda')
            data[idx, idx] = 1.0
        return cls(n, data)

    def to_numpy(self):
        return self.data[:self.n, :self.n].cpu().numpy()

    def has_kernel(self):
        return self.points is not None and self.correlation_scale is not None and self.nu is not None

    def isotropic(self):
        s = self.correlation_scale
        return s is not None and bool(numpy.all(s == s[0]))


class DenseEngine(object):
    """Scratch buffers + cached Cholesky state for one DeviceCorrelation."""

    def __init__(self, K):
        torch = dev.require_cuda()
        self.K = K
        self.n, self.npad = K.n, K.npad
        npad = self.npad
        f64 = torch.float64
        self.A = torch.empty((npad, npad), dtype=f64, device='cuda')       # K + eta I -> L (-> inverse in fused mode)
        self.W = None                                                       # inv(L), allocated on first use
        self.Ainv = None                                                    # explicit inverse (generic path only)
        self.potrf_ws = torch.empty(lib.gp_potrf_workspace_bytes(npad) // 8, dtype=f64, device='cuda')
        self.ws = torch.empty(lib.gp_loglik_workspace_bytes(npad) // 8 + 8, dtype=f64, device='cuda')
        self.info = torch.zeros(1, dtype=torch.int32, device='cuda')
        self.scalar = torch.zeros(8, dtype=f64, device='cuda')
        self._eta = None        # eta whose factor L currently sits in A
        self._have_W = False
        self._have_Ainv = False
        self.launch_count = 0
        self.last_dscale = None  # per-dimension gradient pieces of the last fused call (anisotropic correlation_scale)

    # ---- helpers -------------------------------------------------------------------------------------------
    def _need_W(self):
        if self.W is None:
            self.W = dev.torch.zeros((self.npad, self.npad), dtype=dev.torch.float64, device='cuda')
        return self.W

    def invalidate(self):
        self._eta = None
        self._have_W = False
        self._have_Ainv = False

    def pad_rhs(self, Y):
        """host (n,) or (n, k) array -> device (npad, k) zero-padded; returns (tensor, was_vector)."""
        torch = dev.torch
        Y = numpy.asarray(Y, dtype=numpy.float64)
        vec = (Y.ndim == 1)
        Y2 = Y.reshape(self.n, -1)
        B = torch.zeros((self.npad, Y2.shape[1]), dtype=torch.float64, device='cuda')
        B[:self.n].copy_(torch.from_numpy(numpy.ascontiguousarray(Y2)))
        return B, vec

    # ---- building blocks --------------------------------------------------------------------------------------
    def factor(self, eta):
        """Cholesky of K + eta I (cached per eta). Raises numpy.linalg.LinAlgError when not positive definite,
        as scipy.linalg.solve(assume_a='pos') does in the reference (_linear_solver.py:71)."""
        eta = float(eta)
        if self._eta is not None and self._eta == eta:
            return
        s = dev.stream_ptr()
        self.invalidate()
        check(lib.gp_shift_copy(_p(self.K.data), self.n, self.npad, eta, _p(self.A), s), 'gp_shift_copy')
        check(lib.gp_potrf_f64(_p(self.A), self.n, self.npad, _p(self.info), _p(self.potrf_ws), s), 'gp_potrf_f64')
        info = int(self.info.item())
        if info != 0:
            raise numpy.linalg.LinAlgError(
                '%d-th leading minor of K + eta*I (eta=%g) is not positive definite.' % (info, eta))
        self._eta = eta

    def logdet(self, eta):
        self.factor(eta)
        check(lib.gp_logdet_from_chol(_p(self.A), self.n, self.npad, _p(self.scalar), dev.stream_ptr()),
              'gp_logdet_from_chol')
        return float(self.scalar[0].item())

    def solve(self, eta, Y):
        self.factor(eta)
        Y = numpy.asarray(Y, dtype=numpy.float64)
        vec = (Y.ndim == 1)
        Y2 = Y.reshape(self.n, -1)
        out = numpy.empty_like(Y2)
        for c0 in range(0, Y2.shape[1], MAX_RHS):
            blk = Y2[:, c0:c0 + MAX_RHS]
            B, _ = self.pad_rhs(blk)
            k = B.shape[1]
            check(lib.gp_potrs_f64(_p(self.A), self.npad, _p(self.potrf_ws), _p(B), k, k, dev.stream_ptr()),
                  'gp_potrs_f64')
            out[:, c0:c0 + k] = B[:self.n].cpu().numpy()
        return out[:, 0] if vec else out

    def _trtri(self, eta):
        self.factor(eta)
        if not self._have_W:
            W = self._need_W()
            check(lib.gp_trtri_f64(_p(self.A), _p(W), self.npad, _p(self.potrf_ws), _p(self.ws), dev.stream_ptr()),
                  'gp_trtri_f64')
            self._have_W = True

    def traceinv(self, eta, exponent=1):
        """tr (K + eta I)^-p for p in {1, 2} from the Cholesky factor (imate 'cholesky' method restated:
        tr Kn^-1 = ||inv(L)||_F^2, tr Kn^-2 = ||Kn^-1||_F^2)."""
        torch = dev.torch
        self._trtri(eta)
        s = dev.stream_ptr()
        if exponent == 1:
            check(lib.gp_inverse_traces(_p(self.W), self.n, self.npad, 0, _p(self.scalar), _p(self.ws), s),
                  'gp_inverse_traces')
            return float(self.scalar[1].item())
        if exponent == 2:
            if not self._have_Ainv:
                if self.Ainv is None:
                    self.Ainv = torch.empty((self.npad, self.npad), dtype=torch.float64, device='cuda')
                check(lib.gp_lauum_f64(_p(self.W), _p(self.Ainv), self.npad, s), 'gp_lauum_f64')
                self._have_Ainv = True
            check(lib.gp_inverse_traces(_p(self.Ainv), self.n, self.npad, 1, _p(self.scalar), _p(self.ws), s),
                  'gp_inverse_traces')
            return float(self.scalar[2].item())
        raise ValueError('traceinv on the dense Cholesky engine supports exponent 1 and 2.')

    def trace_K(self):
        """(tr K, tr K^2 = ||K||_F^2) of the unshifted matrix."""
        check(lib.gp_inverse_traces(_p(self.K.data), self.n, self.npad, 1, _p(self.scalar), _p(self.ws),
                                    dev.stream_ptr()), 'gp_inverse_traces')
        v = self.scalar.cpu().numpy()
        return float(v[1]), float(v[2])

    def matmul(self, X):
        """K @ X for host X (n,) or (n, k) -- the reference's K_mixed.dot(0, x) (mixed_correlation.py:305-335)."""
        torch = dev.torch
        X = numpy.asarray(X, dtype=numpy.float64)
        vec = (X.ndim == 1)
        B, _ = self.pad_rhs(X)
        k = B.shape[1]
        out = torch.empty((self.npad, k), dtype=torch.float64, device='cuda')
        check(lib.gp_symm_skinny(_p(self.K.data), self.n, self.npad, _p(B), k, _p(out), dev.stream_ptr()),
              'gp_symm_skinny')
        res = out[:self.n].cpu().numpy()
        return res[:, 0] if vec else res

    # ---- fused evaluation -----------------------------------------------------------------------------------
    def fused(self, eta, R_dev, p, flags):
        """One gp_loglik_dense call. R_dev: device (npad, p) = [X z] zero-padded. Returns host array out[]."""
        torch = dev.torch
        K = self.K
        out = torch.empty(int(lib.gp_loglik_out_len(p)), dtype=torch.float64, device='cuda')
        W = self._need_W() if (flags & 3) else None
        pts = scale = None
        d, nu = 0, 0.0
        aniso = False
        if flags & FLAG_DRHO:
            if not K.has_kernel():
                raise ValueError('d/d(correlation_scale) needs a correlation generated by generate_correlation('
                                 '..., device=True) or MixedCorrelation.set_kernel(points, correlation_scale, nu).')
            pts, scale, d, nu = K.points, K.correlation_scale, K.points.shape[1], float(K.nu)
            aniso = not K.isotropic()
        self.invalidate()  # A / W are overwritten
        kflags = int(flags) & ~FLAG_DRHO if aniso else int(flags)
        rc = lib.gp_loglik_dense(_p(K.data), self.n, self.npad, _p(R_dev), p, float(eta), kflags,
                                 _p(pts) if (pts is not None and not aniso) else None, d if not aniso else 0,
                                 dev.host_ptr(scale) if (scale is not None and not aniso) else None, nu,
                                 _p(self.A), _p(W) if W is not None else None, _p(self.potrf_ws), _p(self.ws), _p(out),
                                 dev.stream_ptr())
        check(rc, 'gp_loglik_dense')
        self.last_dscale = None
        if aniso:
            # one correlation scale per dimension (the reference kernel, _kernels.pyx:107-136): d/d scale[k] for every k
            # from the SAME factorisation - tr(Kn^-1 dK_k) and Q_k = S^T dK_k S, dK_k regenerated from the points
            ds = torch.empty((d, 1 + p * p), dtype=torch.float64, device='cuda')
            for k in range(d):
                check(lib.gp_loglik_dense_dscale(_p(self.A), self.n, self.npad, p, _p(pts), d, dev.host_ptr(scale), nu, k,
                                                 _p(self.ws), _p(ds[k]), dev.stream_ptr()), 'gp_loglik_dense_dscale')
            self.last_dscale = ds
        return out


class EigenEngine(object):
    """``imate_method='eigenvalue'`` -- the method the reference's Likelihood hard-codes (likelihood.py:41,
    mixed_correlation.py:76-79,127-136,172-181,239-248) -- on this library's own kernels (csrc/gp_eig.cu), no cuSOLVER /
    cuBLAS call:

        K = Q T Q^T         gp_sytrd_f64: Householder tridiagonalisation, the rank-2 update fused into the next
                            matrix-vector product (one read + write of the trailing matrix per column, HBM-bound)
        lam = eig(T)        gp_stebz_f64: bisection on Sturm counts, one thread per eigenvalue   (= K_eigenvalues of the
                            reference, mixed_correlation.py:78-79)
        logdet, tr Kn^-1, tr Kn^-2 at any eta: O(n) reductions over lam + eta (gp_eig_reduce)
        a = Q^T [X z]       gp_ormtr_skinny (once per [X z])
        y = (T + eta I)^-1 a   gp_tridiag_solve, O(n p) per eta;  G = a^T y, H = y^T y, T3 = y^T (T + eta I)^-1 y

    so that after ONE reduction per correlation matrix every eta of a sweep row or of a root find costs O(n p) for
    l^ and d l^/d eta. The d/d rho pieces (tr Kn^-1 dK, S^T dK S; an extension the reference does not have) cannot be had
    from the spectrum of K alone: cells that ask for them take the Cholesky evaluator for those two numbers.
    Same out[] layout as DenseEngine.fused / gp_loglik_dense."""

    def __init__(self, K):
        torch = dev.require_cuda()
        self.K = K
        self.n, self.npad = K.n, K.npad
        n = self.n
        f64 = torch.float64
        ldw = n + (n & 1)
        self.ldw = ldw
        self.Q = torch.empty((n, ldw), dtype=f64, device='cuda')        # work copy -> Householder vectors (rows)
        self.Q[:, :n].copy_(K.data[:n, :n])
        self.d = torch.empty(n, dtype=f64, device='cuda')
        self.e = torch.zeros(n, dtype=f64, device='cuda')
        self.tau = torch.zeros(n, dtype=f64, device='cuda')
        self.lam = torch.empty(n, dtype=f64, device='cuda')
        ws = torch.empty(lib.gp_sytrd_workspace_bytes(n) // 8 + 16, dtype=f64, device='cuda')
        s = dev.stream_ptr()
        check(lib.gp_sytrd_f64(_p(self.Q), n, ldw, _p(self.d), _p(self.e), _p(self.tau), _p(ws), s), 'gp_sytrd_f64')
        check(lib.gp_stebz_f64(_p(self.d), _p(self.e), n, _p(self.lam), _p(ws), s), 'gp_stebz_f64')
        self._piv = torch.empty(n, dtype=f64, device='cuda')
        self._red = torch.zeros(8, dtype=f64, device='cuda')
        self._a = None
        self._gram_ws = torch.empty(lib.gp_gram_workspace_bytes(16) // 8, dtype=f64, device='cuda')
        self._chol = None            # Cholesky evaluator for the d/d rho pieces

    # ---- O(n) reductions over the spectrum --------------------------------------------------------------------
    def reductions(self, eta):
        """(logdet Kn, tr Kn^-1, tr Kn^-2, number of eigenvalues with lam + eta <= 0) as host floats"""
        check(lib.gp_eig_reduce(_p(self.lam), self.n, float(eta), _p(self._red), dev.stream_ptr()), 'gp_eig_reduce')
        r = self._red.cpu().numpy()
        return float(r[0]), float(r[1]), float(r[2]), int(r[3])

    def _projected_rhs(self, R_dev, p):
        """a = Q^T R (n x p), cached while the same device block is passed"""
        if self._a is None or self._a[0] is not R_dev or self._a[1].shape[1] != p:
            a = R_dev[:self.n, :p].clone(memory_format=dev.torch.contiguous_format)     # a COPY: ormtr works in place
            check(lib.gp_ormtr_skinny(_p(self.Q), self.n, self.ldw, _p(self.tau), 1, _p(a), p, p, dev.stream_ptr()),
                  'gp_ormtr_skinny')
            self._a = (R_dev, a)
        return self._a[1]

    def _gram(self, X, Y, out_view, p):
        check(lib.gp_gram_skinny(_p(X), _p(Y), self.n, p, _p(out_view), _p(self._gram_ws), dev.stream_ptr()), 'gp_gram_skinny')

    def fused(self, eta, R_dev, p, flags):
        """Returns out[] (device tensor, layout of gp_loglik_dense); out[4] = 1 when K + eta I is not positive definite."""
        torch = dev.torch
        n = self.n
        out = torch.zeros(int(lib.gp_loglik_out_len(p)), dtype=torch.float64, device='cuda')
        s = dev.stream_ptr()
        a = self._projected_rhs(R_dev, p)
        y = torch.empty_like(a)
        info = torch.zeros(2, dtype=torch.float64, device='cuda')
        check(lib.gp_tridiag_solve(_p(self.d), _p(self.e), n, float(eta), _p(a), p, p, _p(y), p, _p(self._piv), _p(info), s),
              'gp_tridiag_solve')
        check(lib.gp_eig_reduce(_p(self.lam), n, float(eta), _p(out), s), 'gp_eig_reduce')     # out[0..2], out[3] = #bad
        out[4] = ((out[3] > 0) | (info[1] > 0)).to(torch.float64)
        out[3] = 0.0
        self._gram(a, y, out[8:8 + p * p], p)
        self._gram(y, y, out[8 + p * p:8 + 2 * p * p], p)
        if flags & FLAG_CUBIC:
            y2 = torch.empty_like(a)
            check(lib.gp_tridiag_solve(_p(self.d), _p(self.e), n, float(eta), _p(y), p, p, _p(y2), p, _p(self._piv), _p(info), s),
                  'gp_tridiag_solve')
            self._gram(y, y2, out[8 + 3 * p * p:8 + 4 * p * p], p)
        if flags & FLAG_DRHO:
            # not a function of the spectrum of K: one Cholesky evaluation supplies tr(Kn^-1 dK) and Q = S^T dK S
            if self._chol is None:
                self._chol = DenseEngine(self.K)
            oc = self._chol.fused(eta, R_dev, p, FLAG_TRACEINV | FLAG_INVERSE | FLAG_DRHO)
            if self._chol.last_dscale is not None:
                raise ValueError("imate_method='eigenvalue': d/d(correlation_scale) per dimension is available on the "
                                 "Cholesky method only.")
            out[3] = oc[3]
            out[8 + 2 * p * p:8 + 3 * p * p] = oc[8 + 2 * p * p:8 + 3 * p * p]
        return out
