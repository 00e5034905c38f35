"""
DirectLikelihood -- log-likelihood in (sigma, sigma0) and its derivatives; same static-method signatures and the
same returned numbers as the reference's gaussian_proc/_likelihood/_direct_likelihood.py (:31-83 log_likelihood,
:89-157 jacobian, :163-270 hessian, :276-340 M_dot, :346-405 maximize_log_likelihood). Plotting is out of scope.

l and its jacobian come from ONE fused device evaluation (one Cholesky) instead of the reference's 4 dposv calls; the
hessian from one fused evaluation with third moments (two skinny solve batches); only the |sigma| < tol limits go
through the generic MixedCorrelation.dot / trace methods. `chain_rule=True` gives sigma-space derivatives.
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
        if with_rho and q.trace_MdK_dims is not None:
            drho = -0.5 * q.trace_MdK_dims + 0.5 * q.zMdKMz_dims / sigma ** 2      # one entry per correlation_scale[k]
        elif with_rho:
            drho = -0.5 * q.trace_MdK + 0.5 * q.zMdKMz / sigma ** 2
        return lp, jac, drho

    @staticmethod
    def log_likelihood_der1_rho(z, X, K_mixed, hyperparam):
        return DirectLikelihood.log_likelihood_and_gradient(z, X, K_mixed, hyperparam, with_rho=True)[2]

    # ---- second derivatives (_direct_likelihood.py:163-270) ------------------------------------------------------
    # Everything up to second order is a polynomial in the moments of the projected precision P = M_eta / sigma^2
    # (M_eta = Kn^-1 - Kn^-1 X B^-1 X^T Kn^-1, Kn = K + eta I, eta = sigma0^2 / sigma^2): with a = sigma^2, b = sigma0^2,
    #     t_k = tr P^k,   s_k = z^T P^k z,   P K P = (P - b P^2) / a      (because P (a K + b I) P = P)
    #     tr(P K) = (n - m - b t_1) / a          tr(P K P) = (t_1 - b t_2) / a        tr(P K P K) = (n - m - 2 b t_1 + b^2 t_2) / a^2
    #     z P K P z = (s_1 - b s_2) / a          z P^2 K P z = (s_2 - b s_3) / a      z P K P K P z = (s_1 - 2 b s_2 + b^2 s_3) / a^2
    # The moments come from ONE fused device evaluation (one factorisation, two skinny solve batches, csrc/gp_loglik.cu
    # flag 8) instead of the reference's five dposv calls.
    @staticmethod
    def _variance_space_derivatives(nm, a, b, t, s):
        """gradient (2,) and Hessian (2, 2) of l with respect to (a, b) = (sigma^2, sigma0^2) from the moments
        t[1..2], s[1..3] of P; nm = n - m. Row / column 0 is a, 1 is b (the reference's ordering)."""
        tr_PK = (nm - b * t[1]) / a
        tr_PKP = (t[1] - b * t[2]) / a
        tr_PKPK = (nm - 2.0 * b * t[1] + b * b * t[2]) / (a * a)
        z_PKP = (s[1] - b * s[2]) / a
        z_P2KP = (s[2] - b * s[3]) / a
        z_PKPKP = (s[1] - 2.0 * b * s[2] + b * b * s[3]) / (a * a)
        grad = numpy.array([-0.5 * tr_PK + 0.5 * z_PKP, -0.5 * t[1] + 0.5 * s[2]])
        hess = numpy.array([[0.5 * tr_PKPK - z_PKPKP, 0.5 * tr_PKP - z_P2KP],
                            [0.5 * tr_PKP - z_P2KP, 0.5 * t[2] - s[3]]])
        return grad, hess

    @staticmethod
    def _moments(z, X, K_mixed, sigma, sigma0):
        """(n - m, a, b, t, s) at (sigma, sigma0): the regular case from the fused evaluator; sigma -> 0 (the covariance
        degenerates to sigma0^2 I, P = Pi / b with the projector Pi = I - X (X^T X)^-1 X^T) in closed form."""
        n, m = X.shape
        a, b = sigma ** 2, sigma0 ** 2
        q = _fused.evaluate(z, X, K_mixed, b / a, traceinv=True, cubic=True)
        tq, sq = q.moments()
        t = [None] + [tq[k] / a ** k for k in (1, 2)]
        s = [None] + [sq[k] / a ** k for k in (1, 2, 3)]
        return n - m, a, b, t, s

    @staticmethod
    def _degenerate_derivatives(z, X, K_mixed, sigma0):
        """sigma = 0 limit of the gradient and Hessian in (sigma^2, sigma0^2): P = Pi / b, so the K-weighted terms need
        Pi K Pi explicitly - formed through K_mixed.dot on an orthonormal basis of span(X) (no factorisation at all)."""
        n, m = X.shape
        b = sigma0 ** 2
        Qx = numpy.linalg.qr(numpy.asarray(X, dtype=float))[0]                  # Pi = I - Qx Qx^T
        proj = lambda v: v - Qx @ (Qx.T @ v)                                     # noqa: E731
        u = proj(numpy.asarray(z, dtype=float))                                  # Pi z
        Ku = K_mixed.dot(0, u)
        KQ = K_mixed.dot(0, Qx)
        C1 = Qx.T @ KQ                                                           # Qx^T K Qx
        C2 = KQ.T @ KQ                                                           # Qx^T K^2 Qx
        tr_K, tr_K2 = K_mixed.trace(0), K_mixed.trace(0, exponent=2)
        tr_PiK = tr_K - numpy.trace(C1)
        tr_PiKPiK = tr_K2 - 2.0 * numpy.trace(C2) + numpy.trace(C1 @ C1)
        PKu = proj(Ku)
        grad = numpy.array([-0.5 * tr_PiK / b + 0.5 * (u @ Ku) / b ** 2, -0.5 * (n - m) / b + 0.5 * (u @ u) / b ** 2])
        hess = numpy.array([[0.5 * tr_PiKPiK / b ** 2 - (Ku @ PKu) / b ** 3, 0.5 * tr_PiK / b ** 2 - (u @ Ku) / b ** 3],
                            [0.5 * tr_PiK / b ** 2 - (u @ Ku) / b ** 3, 0.5 * (n - m) / b ** 2 - (u @ u) / b ** 3]])
        return grad, hess

    @staticmethod
    def log_likelihood_hessian(z, X, K_mixed, sign_switch, hyperparam, chain_rule=False):
        """Second derivatives of l. Default (chain_rule=False): with respect to the variances (sigma^2, sigma0^2) - the
        numbers the reference returns. chain_rule=True: with respect to the optimiser's own variables
        (sigma, sigma0), i.e. d2/dsigma2 = 2 dl/da + 4 sigma^2 d2l/da2 etc., which lets trust-exact converge."""
        sigma, sigma0 = hyperparam[0], hyperparam[1]
        if numpy.abs(sigma) < 1e-16:
            grad, hess = DirectLikelihood._degenerate_derivatives(z, X, K_mixed, sigma0)
        else:
            grad, hess = DirectLikelihood._variance_space_derivatives(*DirectLikelihood._moments(z, X, K_mixed, sigma, sigma0))
        if chain_rule:
            J = numpy.array([2.0 * sigma, 2.0 * sigma0])
            hess = hess * numpy.outer(J, J) + numpy.diag(2.0 * grad)
        return -hess if sign_switch else hess

    @staticmethod
    def log_likelihood_jacobian_sigma(z, X, K_mixed, sign_switch, hyperparam):
        """Gradient with respect to (sigma, sigma0) themselves (chain rule applied to log_likelihood_jacobian)."""
        jac = DirectLikelihood.log_likelihood_jacobian(z, X, K_mixed, False, hyperparam)
        jac = jac * numpy.array([2.0 * hyperparam[0], 2.0 * hyperparam[1]])
        return -jac if sign_switch else jac

    # ---- M dot (_direct_likelihood.py:276-340): M z = Sigma^-1 z - Y Binv Y^T z ----------------------------------------
    @staticmethod
    def M_dot(K_mixed, Binv, Y, sigma, sigma0, z):
        if numpy.abs(sigma) < 1e-8:
            lead = z / sigma0 ** 2
        else:
            lead = K_mixed.solve((sigma0 / sigma) ** 2, z) / sigma ** 2
        return lead - Y @ (Binv @ (Y.T @ z))

    # ---- maximise (_direct_likelihood.py:346-405) -------------------------------------------------------------
    @staticmethod
    def maximize_log_likelihood(z, X, K_mixed, tol=1e-3, hyperparam_guess=[0.2, 0.2], method='Nelder-Mead', chain_rule=False):
        """scipy 'trust-exact' with the analytic jacobian and hessian, start (0.2, 0.2), tol 1e-3 -- the optimiser the
        reference hard-wires at :378-384 (its `method` argument is overridden there; kept for signature parity).
        chain_rule=True hands the optimiser derivatives in its own variables (sigma, sigma0) instead of the reference's
        variance-space ones (with those trust-exact stops with success: False)."""
        print('Maximize log likelihood with sigma sigma0 ...')
        sign_switch = True
        fun = partial(DirectLikelihood.log_likelihood, z, X, K_mixed, sign_switch)
        if chain_rule:
            jac = partial(DirectLikelihood.log_likelihood_jacobian_sigma, z, X, K_mixed, sign_switch)
            hess = partial(DirectLikelihood.log_likelihood_hessian, z, X, K_mixed, sign_switch, chain_rule=True)
        else:
            jac = partial(DirectLikelihood.log_likelihood_jacobian, z, X, K_mixed, sign_switch)
            hess = partial(DirectLikelihood.log_likelihood_hessian, z, X, K_mixed, sign_switch)
        res = scipy.optimize.minimize(fun, hyperparam_guess, method='trust-exact', tol=tol, jac=jac, hess=hess)
        print(res)
        print('Iter: %d, Eval: %d, Success: %s' % (res.nit, res.nfev, res.success))
        sigma, sigma0 = res.x[0], res.x[1]
        return {'sigma': sigma, 'sigma0': sigma0, 'eta': (sigma0 / sigma) ** 2, 'max_lp': -res.fun, 'success': bool(res.success)}
