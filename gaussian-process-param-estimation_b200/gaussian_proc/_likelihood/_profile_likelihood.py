"""
ProfileLikelihood -- log-likelihood in (sigma, eta), its eta-derivative profiled over sigma, and the root-finding
driver; same signatures and returned numbers as gaussian_proc/_likelihood/_profile_likelihood.py (:38-85
log_likelihood, :91-132 log_likelihood_der1_eta, :138-192 der2_eta, :198-238 Nelder-Mead maximisation, :244-415
find_log_likelihood_der1_zeros). Bounds / asymptotes / plots (:456-732) are paper figures and out of scope.
"""

from functools import partial

import numpy
from scipy.optimize import minimize

from . import _fused
from ._root_finding import find_interval_with_sign_change, chandrupatla_method

__all__ = ['ProfileLikelihood']


class ProfileLikelihood(object):

    # ---- log likelihood (:38-85); no 2 pi constant ------------------------------------------------
    @staticmethod
    def log_likelihood(z, X, K_mixed, sign_switch, hyperparam):
        sigma, eta = hyperparam[0], hyperparam[1]
        n, m = X.shape
        q = _fused.evaluate(z, X, K_mixed, eta)
        lp = -0.5 * (n - m) * numpy.log(sigma ** 2) - 0.5 * q.logdet_Kn - 0.5 * numpy.log(numpy.linalg.det(q.B)) \
            - (0.5 / (sigma ** 2)) * q.zMz
        return -lp if sign_switch else lp

    # ---- d l/d eta at sigma^2_hat(eta) (:91-132); argument is log10(eta) ---------------------------
    @staticmethod
    def log_likelihood_der1_eta(z, X, K_mixed, log_eta):
        eta = 0.0 if numpy.isneginf(log_eta) else 10.0 ** log_eta
        n, m = X.shape
        q = _fused.evaluate(z, X, K_mixed, eta, traceinv=True)
        sigma02 = q.zMz / (n - m)
        return -0.5 * (q.trace_M - q.zM2z / sigma02)

    # ---- extension: profiled d l/d rho ----------------------------------------------------------
    @staticmethod
    def log_likelihood_der1_rho(z, X, K_mixed, eta):
        """d l^/d rho = -1/2 tr(M dK) + z^T M dK M z / (2 sigma^2_hat(eta)), isotropic correlation_scale."""
        return ProfileLikelihood.log_likelihood_and_gradient(z, X, K_mixed, eta)[2]

    @staticmethod
    def log_likelihood_and_gradient_async(z, X, K_mixed, eta, with_rho=True):
        """Enqueues the evaluation on torch's current stream; returns a zero-argument callable that completes it and
        returns (l^, d l^/d eta, d l^/d rho). Lets a sweep keep several cells in flight on different streams."""
        h = _fused.evaluate_async(z, X, K_mixed, eta, traceinv=True, drho=with_rho)
        return lambda: ProfileLikelihood._gradient_from(_fused.finish(h), X.shape, with_rho)

    @staticmethod
    def log_likelihood_and_gradient(z, X, K_mixed, eta, with_rho=True):
        """(l^(sigma_hat, eta), d l^/d eta, d l^/d rho) from ONE factorisation -- the unit bench.py counts as a
        'loglik+grad evaluation'."""
        q = _fused.evaluate(z, X, K_mixed, eta, traceinv=True, drho=with_rho)
        return ProfileLikelihood._gradient_from(q, X.shape, with_rho)

    @staticmethod
    def _gradient_from(q, shape, with_rho):
        n, m = shape
        sigma2 = q.zMz / (n - m)
        lp = -0.5 * (n - m) * numpy.log(sigma2) - 0.5 * q.logdet_Kn - 0.5 * numpy.log(numpy.linalg.det(q.B)) \
            - 0.5 * (n - m)
        deta = -0.5 * (q.trace_M - q.zM2z / sigma2)
        if with_rho and q.trace_MdK_dims is not None:
            # anisotropic correlation_scale: the gradient with respect to every scale[k] (array of length d)
            drho = -0.5 * q.trace_MdK_dims + 0.5 * q.zMdKMz_dims / sigma2
        else:
            drho = (-0.5 * q.trace_MdK + 0.5 * q.zMdKMz / sigma2) if with_rho else None
        return lp, deta, drho

    # ---- second derivative, valid at the stationary point only (:138-192) --------------------------
    @staticmethod
    def log_likelihood_der2_eta(z, X, K_mixed, eta):
        """d2 l^/d eta2 at a root of d l^/d eta, from the moments t_k = tr M^k, s_k = z^T M^k z of ONE fused evaluation:
        (n - m) / (2 s_1) * ((t_2 / (n - m) + (t_1 / (n - m))^2) s_1 - 2 s_3)."""
        n, m = X.shape
        t, s = _fused.evaluate(z, X, K_mixed, eta, traceinv=True, cubic=True).moments()
        nm = float(n - m)
        return 0.5 * nm / s[1] * ((t[2] / nm + (t[1] / nm) ** 2) * s[1] - 2.0 * s[3])

    # ---- Nelder-Mead over (sigma, eta) (:198-238) --------------------------------------------------------------
    @staticmethod
    def maximize_log_likelihood_with_sigma_eta(z, X, K_mixed, tol=1e-6, hyperparam_guess=[0.1, 0.1],
                                               method='Nelder-Mead'):
        print('Maximize log likelihood with sigma eta ...')
        fun = partial(ProfileLikelihood.log_likelihood, z, X, K_mixed, True)
        res = minimize(fun, hyperparam_guess, method='Nelder-Mead', tol=tol)
        print('Iter: %d, Eval: %d, success: %s' % (res.nit, res.nfev, res.success))
        sigma, eta = res.x[0], res.x[1]
        return {'sigma': sigma, 'sigma0': numpy.sqrt(eta) * sigma, 'eta': eta, 'max_lp': -res.fun}

    # ---- root of d l/d eta (:244-415) ----------------------------------------------------------------------------
    @staticmethod
    def find_log_likelihood_der1_zeros(z, X, K_mixed, interval_eta, tol=1e-6, max_iterations=100,
                                       num_bracket_trials=3):

        def find_optimal_sigma(eta):
            n, m = X.shape
            q = _fused.evaluate(z, X, K_mixed, eta)
            return numpy.sqrt(q.zMz / (n - m))

        def find_optimal_sigma0():
            n, m = X.shape
            Binv = numpy.linalg.inv(numpy.matmul(X.T, X))
            v = numpy.matmul(X, numpy.matmul(Binv, numpy.matmul(X.T, z)))
            return numpy.sqrt(numpy.dot(z, z - v) / (n - m))

        print('Find root of log likelihood derivative ...')
        tune = getattr(K_mixed, 'sparse', False) and 'eager_rhs_basis' not in K_mixed.imate_options
        if tune:
            # the root find asks many eta of this operator: keep the Krylov run of [X z] from the first one on - for the
            # duration of this call only (the caller's operator gets its own setting back)
            saved = K_mixed.engine.opt.get('eager_rhs_basis', False)
            K_mixed.engine.opt['eager_rhs_basis'] = True
        try:
            return ProfileLikelihood._find_der1_zeros(z, X, K_mixed, interval_eta, tol, max_iterations, num_bracket_trials,
                                                      find_optimal_sigma, find_optimal_sigma0)
        finally:
            if tune:
                K_mixed.engine.opt['eager_rhs_basis'] = saved

    @staticmethod
    def _find_der1_zeros(z, X, K_mixed, interval_eta, tol, max_iterations, num_bracket_trials, find_optimal_sigma,
                         find_optimal_sigma0):
        f = partial(ProfileLikelihood.log_likelihood_der1_eta, z, X, K_mixed)
        bracket = [numpy.log10(interval_eta[0]), numpy.log10(interval_eta[1])]
        bracket_found, bracket, bracket_values = find_interval_with_sign_change(f, bracket, num_bracket_trials, args=(), )

        if bracket_found:
            res = chandrupatla_method(f, bracket, bracket_values, verbose=False, eps_m=tol, eps_a=tol,
                                      maxiter=max_iterations)
            print('Iter: %d' % (res['iterations']))
            eta = 10 ** res['root']
            sigma = find_optimal_sigma(eta)
            sigma0 = numpy.sqrt(eta) * sigma
            success = True
        else:
            # No sign change on the interval (:352-405): d l^/d eta keeps one sign, so the maximiser sits at an end of
            # [0, inf). The curvature at eta = 0 tells which: the same sign as the slope means the slope does not turn
            # around towards zero -> eta = 0, otherwise the extremum is at infinity.
            slopes = numpy.sign([bracket_values[0], bracket_values[1]])
            if slopes[0] != slopes[1] or slopes[0] == 0:
                raise ValueError('eta must be zero or inf at this point.')
            curvature = ProfileLikelihood.log_likelihood_der2_eta(z, X, K_mixed, 0.0)
            for label, where, value in (('dL/deta  ', bracket[0], bracket_values[0]), ('dL/deta  ', bracket[1], bracket_values[1])):
                print('%s at eta = %0.2e:\t %0.6g' % (label, where, value))
            print('d2L/deta2 at eta = 0:\t %0.6g' % curvature)
            if numpy.sign(curvature) == slopes[0]:
                eta, sigma0, sigma = 0.0, 0, find_optimal_sigma(0.0)
            else:
                eta, sigma, sigma0 = numpy.inf, 0, find_optimal_sigma0()
            success = True
        return {'sigma': sigma, 'sigma0': sigma0, 'eta': eta, 'success': success}
