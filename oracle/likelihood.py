"""TEST INFRASTRUCTURE. CPU oracle for the mixed-correlation operator and the likelihood functions.

Restates (NumPy/SciPy, same formulas and evaluation order where it matters):
  MixedCorrelation              gaussian_proc/_mixed_correlation/mixed_correlation.py:34-335
  linear_solver                 gaussian_proc/_mixed_correlation/_linear_solver.py:24-73
  DirectLikelihood              gaussian_proc/_likelihood/_direct_likelihood.py:31-340
  ProfileLikelihood             gaussian_proc/_likelihood/_profile_likelihood.py:38-132, 244-415
  root finding                  gaussian_proc/_likelihood/_root_finding.py:21-309
imate (absent third-party dependency, unpinned in requirements.txt:5) is restated for the two deterministic
methods only: 'eigenvalue' (reductions over eigenvalues of K computed by scipy.linalg.eigh, mixed_correlation.py:76-79)
and 'cholesky' (2 sum log L_ii; trace of the explicit inverse).
Extensions that the reference does not have (d/d correlation_scale) are marked EXTENSION.
"""

import numpy
import scipy.linalg
import scipy.sparse
import scipy.sparse.linalg


class MixedCorrelation(object):
    """K + eta I (mixed_correlation.py:34-335); imate_method in {'eigenvalue', 'cholesky'}."""

    def __init__(self, K, imate_method='cholesky'):
        self.K = K
        self.imate_method = imate_method
        self.sparse = scipy.sparse.issparse(K)
        n = K.shape[0]
        self.I = scipy.sparse.eye(n, format='csr') if self.sparse else numpy.eye(n)
        self.K_eigenvalues = None
        if imate_method == 'eigenvalue':
            self.K_eigenvalues = scipy.linalg.eigh(K, eigvals_only=True, check_finite=False)
        elif imate_method != 'cholesky':
            raise ValueError('oracle supports imate_method "eigenvalue" and "cholesky" only')

    def get_matrix_size(self):
        return self.K.shape[0]

    def _dense_Kn(self, eta):
        Kd = self.K.toarray() if self.sparse else self.K
        return Kd + eta * numpy.eye(Kd.shape[0])

    def trace(self, eta, exponent=1):
        """:96-149 (exact branches for exponent 0, 1, 2)."""
        n = self.K.shape[0]
        if exponent == 0:
            return float(n)
        trK = self.K.diagonal().sum()
        if exponent == 1:
            return trK + eta * n
        if exponent == 2:
            trK2 = (self.K.multiply(self.K)).sum() if self.sparse else numpy.sum(self.K * self.K)
            return trK2 + 2.0 * eta * trK + eta ** 2 * n
        return numpy.sum((self.K_eigenvalues + eta) ** exponent)

    def traceinv(self, eta, exponent=1):
        """:155-215"""
        if self.imate_method == 'eigenvalue':
            return numpy.sum(1.0 / (self.K_eigenvalues + eta) ** exponent)
        Kinv = numpy.linalg.inv(self._dense_Kn(eta))
        return numpy.trace(numpy.linalg.matrix_power(Kinv, exponent))

    def logdet(self, eta, exponent=1):
        """:221-274"""
        if self.imate_method == 'eigenvalue':
            return exponent * numpy.sum(numpy.log(self.K_eigenvalues + eta))
        L = numpy.linalg.cholesky(self._dense_Kn(eta))
        return exponent * 2.0 * numpy.sum(numpy.log(numpy.diag(L)))

    def solve(self, eta, Y):
        """:280-299 -> linear_solver(Kn, Y, 'sym_pos') (_linear_solver.py:24-73). Dense: LAPACK posv. Sparse: CG with
        the reference's tol=1e-6 (spelled rtol for SciPy >= 1.14, atol=0)."""
        Kn = self.K + eta * self.I
        if self.sparse:
            if Y.ndim == 1:
                return scipy.sparse.linalg.cg(Kn, Y, rtol=1e-6, atol=0)[0]
            X = numpy.zeros(Y.shape, order='F')
            for i in range(Y.shape[1]):
                X[:, i] = scipy.sparse.linalg.cg(Kn, Y[:, i], rtol=1e-6, atol=0)[0]
            return X
        return scipy.linalg.solve(Kn, Y, assume_a='pos')

    def dot(self, eta, x, exponent=1):
        """:305-335 (exponent 1; the reference's exponent-2 branch returns 2 K x, SURVEY Q5 -- restated as is)."""
        y = numpy.zeros_like(x)
        for _ in range(exponent):
            y = y + self.K.dot(x)
            if eta != 0:
                y = y + eta * x
        return y


# =================
# Direct likelihood  (sigma, sigma0)
# =================

class DirectLikelihood(object):

    @staticmethod
    def M_dot(K_mixed, Binv, Y, sigma, sigma0, z):
        """_direct_likelihood.py:276-340"""
        if numpy.abs(sigma) < 1e-8:
            w = z / sigma0 ** 2
        else:
            w = K_mixed.solve((sigma0 / sigma) ** 2, z) / sigma ** 2
        return w - Y @ (Binv @ (Y.T @ z))

    @staticmethod
    def log_likelihood(z, X, K_mixed, sign_switch, hyperparam):
        """_direct_likelihood.py:31-83"""
        sigma, sigma0 = hyperparam[0], hyperparam[1]
        n, m = X.shape
        if numpy.abs(sigma) < 1e-8:
            logdet_S = n * numpy.log(sigma0 ** 2)
            Y = X / sigma0 ** 2
        else:
            eta = (sigma0 / sigma) ** 2
            logdet_S = n * numpy.log(sigma ** 2) + K_mixed.logdet(eta)
            Y = K_mixed.solve(eta, X) / sigma ** 2
        B = X.T @ Y
        logdet_B = numpy.log(numpy.linalg.det(B))
        Mz = DirectLikelihood.M_dot(K_mixed, numpy.linalg.inv(B), Y, sigma, sigma0, z)
        lp = -0.5 * (n - m) * numpy.log(2.0 * numpy.pi) - 0.5 * logdet_S - 0.5 * logdet_B - 0.5 * numpy.dot(z, Mz)
        return -lp if sign_switch else lp

    @staticmethod
    def log_likelihood_jacobian(z, X, K_mixed, sign_switch, hyperparam):
        """_direct_likelihood.py:89-157; derivatives w.r.t. the VARIANCES sigma^2, sigma0^2."""
        sigma, sigma0 = hyperparam[0], hyperparam[1]
        n, m = X.shape
        small = numpy.abs(sigma) < 1e-8
        if small:
            Y = X / sigma0 ** 2
        else:
            eta = (sigma0 / sigma) ** 2
            Y = K_mixed.solve(eta, X) / sigma ** 2
        Binv = numpy.linalg.inv(X.T @ Y)
        Mz = DirectLikelihood.M_dot(K_mixed, Binv, Y, sigma, sigma0, z)
        KMz = K_mixed.dot(0, Mz)
        zMMz = numpy.dot(Mz, Mz)
        zMKMz = numpy.dot(Mz, KMz)
        if small:
            trace_M = (n - m) / sigma0 ** 2
            trace_KM = K_mixed.trace(0) / sigma0 ** 2 - numpy.trace(Binv @ (Y.T @ K_mixed.dot(0, Y)))
        else:
            trace_M = K_mixed.traceinv(eta) / sigma ** 2 - numpy.trace(Binv @ (Y.T @ Y))
            trace_KM = (n - m) / sigma ** 2 - eta * trace_M
        jac = numpy.array([-0.5 * trace_KM + 0.5 * zMKMz, -0.5 * trace_M + 0.5 * zMMz], dtype=float)
        return -jac if sign_switch else jac

    @staticmethod
    def log_likelihood_hessian(z, X, K_mixed, sign_switch, hyperparam):
        """_direct_likelihood.py:163-270 (regular branch |sigma| >= 1e-16 only)."""
        sigma, sigma0 = hyperparam[0], hyperparam[1]
        n, m = X.shape
        eta = (sigma0 / sigma) ** 2
        Y = K_mixed.solve(eta, X) / sigma ** 2
        V = K_mixed.solve(eta, Y) / sigma ** 2
        Binv = numpy.linalg.inv(X.T @ Y)
        A = Binv @ (Y.T @ Y)
        Mz = DirectLikelihood.M_dot(K_mixed, Binv, Y, sigma, sigma0, z)
        MMz = DirectLikelihood.M_dot(K_mixed, Binv, Y, sigma, sigma0, Mz)
        KMz = K_mixed.dot(0, Mz)
        MKMz = DirectLikelihood.M_dot(K_mixed, Binv, Y, sigma, sigma0, KMz)
        zMMMz = numpy.dot(Mz, MMz)
        zMMKMz = numpy.dot(MMz, KMz)
        zMKMKMz = numpy.dot(KMz, MKMz)
        trace_M = K_mixed.traceinv(eta) / sigma ** 2 - numpy.trace(A)
        trace_S2inv = K_mixed.traceinv(eta, exponent=2) / sigma ** 4
        trace_M2 = trace_S2inv - 2.0 * numpy.trace(Binv @ (Y.T @ V)) + numpy.trace(A @ A)
        trace_KMKM = (n - m) / sigma ** 4 - (2 * eta / sigma ** 2) * trace_M + eta ** 2 * trace_M2
        trace_KMM = trace_M / sigma ** 2 - eta * trace_M2
        d00 = 0.5 * (trace_M2 - 2.0 * zMMMz)
        d11 = 0.5 * (trace_KMKM - 2.0 * zMKMKMz)
        d10 = 0.5 * (trace_KMM - 2.0 * zMMKMz)
        H = numpy.array([[d11, d10], [d10, d00]], dtype=float)
        return -H if sign_switch else H

    @staticmethod
    def log_likelihood_der1_rho(z, X, K_mixed, dK, hyperparam):
        """EXTENSION (not in the reference): d l / d rho = -1/2 sigma^2 tr(M dK) + 1/2 sigma^2 z^T M dK M z
        with M built exactly as the reference's M_dot builds it; tr(M dK) = tr(Sinv dK) - tr(Binv Y^T dK Y)."""
        sigma, sigma0 = hyperparam[0], hyperparam[1]
        eta = (sigma0 / sigma) ** 2
        Y = K_mixed.solve(eta, X) / sigma ** 2
        Binv = numpy.linalg.inv(X.T @ Y)
        Mz = DirectLikelihood.M_dot(K_mixed, Binv, Y, sigma, sigma0, z)
        Sinv_dK = K_mixed.solve(eta, dK) / sigma ** 2
        trace_MdK = numpy.trace(Sinv_dK) - numpy.trace(Binv @ (Y.T @ (dK @ Y)))
        return -0.5 * sigma ** 2 * trace_MdK + 0.5 * sigma ** 2 * numpy.dot(Mz, dK @ Mz)


# ==================
# Profile likelihood  (sigma, eta)
# ==================

class ProfileLikelihood(object):

    @staticmethod
    def _parts(z, X, K_mixed, eta):
        Y = K_mixed.solve(eta, X)
        w = K_mixed.solve(eta, z)
        B = X.T @ Y
        Binv = numpy.linalg.inv(B)
        Mz = w - Y @ (Binv @ (Y.T @ z))
        return Y, w, B, Binv, Mz

    @staticmethod
    def log_likelihood(z, X, K_mixed, sign_switch, hyperparam):
        """_profile_likelihood.py:38-85 (no 2 pi constant)."""
        sigma, eta = hyperparam[0], hyperparam[1]
        n, m = X.shape
        logdet_Kn = K_mixed.logdet(eta)
        Y, w, B, Binv, _ = ProfileLikelihood._parts(z, X, K_mixed, eta)
        YBinvYt = Y @ (Binv @ Y.T)
        lp = -0.5 * (n - m) * numpy.log(sigma ** 2) - 0.5 * logdet_Kn - 0.5 * numpy.log(numpy.linalg.det(B)) \
            - (0.5 / sigma ** 2) * numpy.dot(z, w - YBinvYt @ z)
        return -lp if sign_switch else lp

    @staticmethod
    def log_likelihood_der1_eta(z, X, K_mixed, log_eta):
        """_profile_likelihood.py:91-132: argument log10(eta), value d l / d eta."""
        eta = 0.0 if numpy.isneginf(log_eta) else 10.0 ** log_eta
        n, m = X.shape
        Y, w, B, Binv, Mz = ProfileLikelihood._parts(z, X, K_mixed, eta)
        trace_M = K_mixed.traceinv(eta) - numpy.trace(Binv @ (Y.T @ Y))
        zMz = numpy.dot(z, Mz)
        zM2z = numpy.dot(Mz, Mz)
        sigma02 = zMz / (n - m)
        return -0.5 * (trace_M - zM2z / sigma02)

    @staticmethod
    def find_optimal_sigma(z, X, K_mixed, eta):
        """:267-281"""
        n, m = X.shape
        _, _, _, _, Mz = ProfileLikelihood._parts(z, X, K_mixed, eta)
        return numpy.sqrt(numpy.dot(z, Mz) / (n - m))

    @staticmethod
    def log_likelihood_der1_rho(z, X, K_mixed, dK, eta):
        """EXTENSION: profiled d l^/d rho = -1/2 tr(M_eta dK) + z^T M dK M z / (2 sigma^2_hat)."""
        n, m = X.shape
        Y, w, B, Binv, Mz = ProfileLikelihood._parts(z, X, K_mixed, eta)
        sigma2 = numpy.dot(z, Mz) / (n - m)
        trace_MdK = numpy.trace(K_mixed.solve(eta, dK)) - numpy.trace(Binv @ (Y.T @ (dK @ Y)))
        return -0.5 * trace_MdK + 0.5 * numpy.dot(Mz, dK @ Mz) / sigma2

    @staticmethod
    def find_log_likelihood_der1_zeros(z, X, K_mixed, interval_eta, tol=1e-6, max_iterations=100,
                                       num_bracket_trials=3):
        """:244-415, sign-change branch (the eta -> 0 / inf fallbacks need der2_eta and are not on the measured path)."""
        f = lambda t: ProfileLikelihood.log_likelihood_der1_eta(z, X, K_mixed, t)  # noqa: E731
        bracket = [numpy.log10(interval_eta[0]), numpy.log10(interval_eta[1])]
        found, bracket, values = find_interval_with_sign_change(f, bracket, num_bracket_trials)
        if not found:
            raise ValueError('oracle: no sign change of d l/d eta in the interval')
        res = chandrupatla_method(f, bracket, values, eps_m=tol, eps_a=tol, maxiter=max_iterations)
        eta = 10 ** res['root']
        sigma = ProfileLikelihood.find_optimal_sigma(z, X, K_mixed, eta)
        return {'sigma': sigma, 'sigma0': numpy.sqrt(eta) * sigma, 'eta': eta, 'success': True,
                'iterations': res['iterations']}


# ============
# root finding  (scalar restatement of _root_finding.py)
# ============

def find_interval_with_sign_change(f, bracket, num_bracket_trials):
    """_root_finding.py:21-148: bisect towards the smaller |f|, else extrapolate by half an interval."""
    x0, x1 = bracket
    f0, f1 = f(x0), f(x1)
    trials = 0
    while trials < num_bracket_trials:
        trials += 1
        if numpy.sign(f0) != numpy.sign(f1):
            return True, [x0, x1], [f0, f1]
        xm = x0 * 0.5 + x1 * 0.5
        fm = f(xm)
        if numpy.sign(f0) != numpy.sign(fm):
            if abs(f0) < abs(f1):
                return True, [x0, xm], [f0, fm]
            return True, [xm, x1], [fm, f1]
        if abs(fm) < min(abs(f0), abs(f1)):
            if abs(f0) < abs(f1):
                x1, f1 = xm, fm
            else:
                x0, f0 = xm, fm
            continue
        t = 1.5 if abs(f0) > abs(f1) else -0.5
        xe = x0 * (1 - t) + x1 * t
        fe = f(xe)
        if numpy.sign(f0) != numpy.sign(fe):
            if abs(f0) > abs(f1):
                return True, [xe, x0], [fe, f0]
            return True, [x1, xe], [f1, fe]
        if t > 0:
            x0, f0, x1, f1 = x1, f1, xe, fe
        else:
            x1, f1, x0, f0 = x0, f0, xe, fe
    return False, [x0, x1], [f0, f1]


def chandrupatla_method(f, bracket, bracket_values, eps_m, eps_a, maxiter=50):
    """_root_finding.py:155-309 for scalar f: inverse quadratic interpolation when the IQI validity test
    phi^2 < xi and (1-phi)^2 < 1-xi holds, bisection otherwise; stop when fm == 0 or tlim > 0.5."""
    b, a = bracket[0], bracket[1]
    fb, fa = bracket_values[0], bracket_values[1]
    c, fc = a, fa
    t = 0.5
    iterations = 0
    xm = a
    while maxiter > 0:
        maxiter -= 1
        xt = a + t * (b - a)
        ft = f(xt)
        if numpy.sign(ft) == numpy.sign(fa):
            c, fc = a, fa
        else:
            c, b, fc, fb = b, a, fb, fa
        a, fa = xt, ft
        if abs(fa) < abs(fb):
            xm, fm = a, fa
        else:
            xm, fm = b, fb
        tol = 2 * eps_m * abs(xm) + eps_a
        tlim = tol / abs(b - c)
        if fm == 0 or tlim > 0.5:
            break
        iterations += 1
        xi = (a - b) / (c - b)
        phi = (fa - fb) / (fc - fb)
        if phi ** 2 < xi and (1 - phi) ** 2 < 1 - xi:
            t = fa / (fb - fa) * fc / (fb - fc) + (c - a) / (b - a) * fa / (fc - fa) * fb / (fc - fb)
        else:
            t = 0.5
        t = min(1 - tlim, max(tlim, t))
    return {'root': xm, 'iterations': iterations}
