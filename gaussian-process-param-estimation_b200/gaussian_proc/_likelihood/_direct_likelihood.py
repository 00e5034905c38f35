"""
DirectLikelihood -- log-likelihood in (sigma, sigma0) and its derivatives; same static-method signatures and the
same returned numbers as the reference's gaussian_proc/_likelihood/_direct_likelihood.py (:31-83 log_likelihood,
:89-157 jacobian, :163-270 hessian, :276-340 M_dot, :346-405 maximize_log_likelihood). Plotting is out of scope.

l and its jacobian come from ONE fused device evaluation (one Cholesky) instead of the reference's 4 dposv calls; the
hessian and the |sigma| < tol branches go through the generic MixedCorrelation methods (cached factor).
The jacobian/hessian are, like the reference's, derivatives with respect to the variances sigma^2, sigma0^2.
"""

from functools import partial

import numpy
import scipy.optimize

from . import _fused

__all__ = ['DirectLikelihood']


class DirectLikelihood(object):

    # ---- log likelihood (_direct_likelihood.py:31-83) -------------------------------------------------------
    @staticmethod
    def log_likelihood(z, X, K_mixed, sign_switch, hyperparam):
        sigma, sigma0 = hyperparam[0], hyperparam[1]
        n, m = X.shape
        tol = 1e-8
        if numpy.abs(sigma) < tol:
            # sigma^2 K is ignored against sigma0^2 I (:49-55)
            logdet_S = n * numpy.log(sigma0 ** 2)
            Y = X / sigma0 ** 2
            B = numpy.matmul(X.T, Y)
            logdet_B = numpy.log(numpy.linalg.det(B))
            Mz = DirectLikelihood.M_dot(K_mixed, numpy.linalg.inv(B), Y, sigma, sigma0, z)
            zMz = numpy.dot(z, Mz)
        else:
            eta = (sigma0 / sigma) ** 2
            q = _fused.evaluate(z, X, K_mixed, eta)
            logdet_S = n * numpy.log(sigma ** 2) + q.logdet_Kn
            logdet_B = numpy.log(numpy.linalg.det(q.B / sigma ** 2))
            zMz = q.zMz / sigma ** 2
        lp = -0.5 * (n - m) * numpy.log(2.0 * numpy.pi) - 0.5 * logdet_S - 0.5 * logdet_B - 0.5 * zMz
        return -lp if sign_switch else lp

    # ---- jacobian (_direct_likelihood.py:89-157) ------------------------------------------------------------
    @staticmethod
    def log_likelihood_jacobian(z, X, K_mixed, sign_switch, hyperparam):
        sigma, sigma0 = hyperparam[0], hyperparam[1]
        n, m = X.shape
        tol = 1e-8
        if numpy.abs(sigma) < tol:
            Y = X / sigma0 ** 2
            Binv = numpy.linalg.inv(numpy.matmul(X.T, Y))
            Mz = DirectLikelihood.M_dot(K_mixed, Binv, Y, sigma, sigma0, z)
            KMz = K_mixed.dot(0, Mz)
            zMMz = numpy.dot(Mz, Mz)
            zMKMz = numpy.dot(Mz, KMz)
            trace_M = (n - m) / sigma0 ** 2
            YtKY = numpy.matmul(Y.T, K_mixed.dot(0, Y))
            trace_KM = K_mixed.trace(0) / sigma0 ** 2 - numpy.trace(numpy.matmul(Binv, YtKY))
        else:
            eta = (sigma0 / sigma) ** 2
            q = _fused.evaluate(z, X, K_mixed, eta, traceinv=True)
            zMMz = q.zM2z / sigma ** 4
            zMKMz = q.zMKMz / sigma ** 4
            trace_M = q.trace_M / sigma ** 2
            trace_KM = (n - m) / sigma ** 2 - eta * trace_M
        jacobian = numpy.array([-0.5 * trace_KM + 0.5 * zMKMz, -0.5 * trace_M + 0.5 * zMMz], dtype=float)
        return -jacobian if sign_switch else jacobian

    # ---- extension: l, d l/d(sigma^2), d l/d(sigma0^2), d l/d(rho) from one factorisation ----------------------
    @staticmethod
    def log_likelihood_and_gradient(z, X, K_mixed, hyperparam, with_rho=True):
        """Returns (l, jacobian[2], dl_drho). EXTENSION named by BASELINE.json's north_star: the derivative with
        respect to an isotropic correlation_scale rho, d l/d rho = -1/2 sigma^2 tr(M dK) + 1/2 sigma^2 z^T M dK M z
        (SURVEY 8a A9), evaluated with dK/d rho regenerated on the fly."""
        sigma, sigma0 = hyperparam[0], hyperparam[1]
        n, m = X.shape
        eta = (sigma0 / sigma) ** 2
        q = _fused.evaluate(z, X, K_mixed, eta, traceinv=True, drho=with_rho)
        lp = -0.5 * (n - m) * numpy.log(2.0 * numpy.pi) - 0.5 * (n * numpy.log(sigma ** 2) + q.logdet_Kn) \
            - 0.5 * numpy.log(numpy.linalg.det(q.B / sigma ** 2)) - 0.5 * q.zMz / sigma ** 2
        trace_M = q.trace_M / sigma ** 2
        trace_KM = (n - m) / sigma ** 2 - eta * trace_M
        jac = numpy.array([-0.5 * trace_KM + 0.5 * q.zMKMz / sigma ** 4, -0.5 * trace_M + 0.5 * q.zM2z / sigma ** 4])
        drho = None
        if with_rho:
            drho = -0.5 * q.trace_MdK + 0.5 * q.zMdKMz / sigma ** 2
        return lp, jac, drho

    @staticmethod
    def log_likelihood_der1_rho(z, X, K_mixed, hyperparam):
        return DirectLikelihood.log_likelihood_and_gradient(z, X, K_mixed, hyperparam, with_rho=True)[2]

    # ---- hessian (_direct_likelihood.py:163-270) ------------------------------------------------------------
    @staticmethod
    def log_likelihood_hessian(z, X, K_mixed, sign_switch, hyperparam):
        sigma, sigma0 = hyperparam[0], hyperparam[1]
        n, m = X.shape
        tol = 1e-16
        small = numpy.abs(sigma) < tol
        if small:
            Y = X / sigma0 ** 2
            V = Y / sigma0 ** 2
        else:
            eta = (sigma0 / sigma) ** 2
            Y = K_mixed.solve(eta, X) / sigma ** 2
            V = K_mixed.solve(eta, Y) / sigma ** 2
        Binv = numpy.linalg.inv(numpy.matmul(X.T, Y))
        A = numpy.matmul(Binv, numpy.matmul(Y.T, Y))
        Mz = DirectLikelihood.M_dot(K_mixed, Binv, Y, sigma, sigma0, z)
        MMz = DirectLikelihood.M_dot(K_mixed, Binv, Y, sigma, sigma0, Mz)
        KMz = K_mixed.dot(0, Mz)
        zMMMz = numpy.dot(Mz, MMz)
        MKMz = DirectLikelihood.M_dot(K_mixed, Binv, Y, sigma, sigma0, KMz)
        zMMKMz = numpy.dot(MMz, KMz)
        zMKMKMz = numpy.dot(KMz, MKMz)
        if small:
            trace_M = (n - m) / sigma0 ** 2
            trace_S2inv = n / sigma0 ** 4
        else:
            trace_M = K_mixed.traceinv(eta) / sigma ** 2 - numpy.trace(A)
            trace_S2inv = K_mixed.traceinv(eta, exponent=2) / sigma ** 4
        trace_C = numpy.trace(numpy.matmul(Binv, numpy.matmul(Y.T, V)))
        trace_M2 = trace_S2inv - 2.0 * trace_C + numpy.trace(numpy.matmul(A, A))
        if small:
            trace_K2 = K_mixed.trace(0, exponent=2)
            D = numpy.matmul(X.T, X)
            E = numpy.matmul(K_mixed.dot(0, X, exponent=2), D)
            trace_KMKM = (trace_K2 - 2.0 * numpy.trace(E) + numpy.trace(numpy.matmul(E, E))) / sigma0 ** 4
            YtKY = numpy.matmul(Y.T, K_mixed.dot(0, Y))
            trace_KM = K_mixed.trace(0) / sigma0 ** 2 - numpy.trace(numpy.matmul(Binv, YtKY))
            trace_KMM = trace_KM / sigma0 ** 2
        else:
            trace_KMKM = (n - m) / sigma ** 4 - (2 * eta / sigma ** 2) * trace_M + (eta ** 2) * trace_M2
            trace_KMM = trace_M / sigma ** 2 - eta * trace_M2
        der2_sigma0_sigma0 = 0.5 * (trace_M2 - 2.0 * zMMMz)
        der2_sigma_sigma = 0.5 * (trace_KMKM - 2.0 * zMKMKMz)
        der2_sigma_sigma0 = 0.5 * (trace_KMM - 2.0 * zMMKMz)
        hessian = numpy.array([[der2_sigma_sigma, der2_sigma_sigma0], [der2_sigma_sigma0, der2_sigma0_sigma0]],
                              dtype=float)
        return -hessian if sign_switch else hessian

    # ---- M dot (_direct_likelihood.py:276-340) ----------------------------------------------------------------
    @staticmethod
    def M_dot(K_mixed, Binv, Y, sigma, sigma0, z):
        tol = 1e-8
        if numpy.abs(sigma) < tol:
            w = z / sigma0 ** 2
        else:
            eta = (sigma0 / sigma) ** 2
            w = K_mixed.solve(eta, z) / sigma ** 2
        return w - numpy.matmul(Y, numpy.matmul(Binv, numpy.matmul(Y.T, z)))

    # ---- maximise (_direct_likelihood.py:346-405) -------------------------------------------------------------
    @staticmethod
    def maximize_log_likelihood(z, X, K_mixed, tol=1e-3, hyperparam_guess=[0.2, 0.2], method='Nelder-Mead'):
        """scipy 'trust-exact' with the analytic jacobian and hessian, start (0.2, 0.2), tol 1e-3 -- the optimiser the
        reference hard-wires at :378-384 (its `method` argument is overridden there; kept for signature parity)."""
        print('Maximize log likelihood with sigma sigma0 ...')
        sign_switch = True
        fun = partial(DirectLikelihood.log_likelihood, z, X, K_mixed, sign_switch)
        jac = partial(DirectLikelihood.log_likelihood_jacobian, z, X, K_mixed, sign_switch)
        hess = partial(DirectLikelihood.log_likelihood_hessian, z, X, K_mixed, sign_switch)
        res = scipy.optimize.minimize(fun, hyperparam_guess, method='trust-exact', tol=tol, jac=jac, hess=hess)
        print(res)
        print('Iter: %d, Eval: %d, Success: %s' % (res.nit, res.nfev, res.success))
        sigma, sigma0 = res.x[0], res.x[1]
        return {'sigma': sigma, 'sigma0': sigma0, 'eta': (sigma0 / sigma) ** 2, 'max_lp': -res.fun}
