"""
Likelihood -- thin dispatcher, same as the reference's gaussian_proc/_likelihood/likelihood.py:23-102.
The reference hard-codes imate_method='eigenvalue' (:40-43), which bars sparse K; here the method is a
pass-through keyword. Default: 'cholesky' for dense K (same values as the reference's 'eigenvalue' to rounding, one
factorisation per evaluation and no O(n^3) eigen-decomposition up front), 'slq' for sparse; 'eigenvalue' is accepted.
"""

import scipy.sparse

from .._mixed_correlation import MixedCorrelation
from ._direct_likelihood import DirectLikelihood
from ._profile_likelihood import ProfileLikelihood

__all__ = ['Likelihood']


class Likelihood(object):

    def __init__(self, X, K, likelihood_method='direct', imate_method=None, imate_options={}):
        self.X = X
        self.K = K
        self.likelihood_method = likelihood_method
        if imate_method is None:
            is_sparse = scipy.sparse.issparse(K) or type(K).__name__ in ('DeviceCSR', 'DeviceRowBlocks')
            imate_method = 'slq' if is_sparse else 'cholesky'
        self.K_mixed = MixedCorrelation(self.K, interpolate=False, imate_method=imate_method,
                                        imate_options=imate_options)

    def likelihood(self, z, hyperparam):
        """likelihood.py:55-61"""
        return DirectLikelihood.log_likelihood(z, self.X, self.K_mixed, False, hyperparam)

    def maximize_log_likelihood(self, z, plot=False, interval_eta=None):
        """likelihood.py:67-102 ('direct' -> trust-exact over (sigma, sigma0); 'profiled' -> root of d l/d eta on
        eta in [1e-4, 1e3], the reference's hard-coded interval :88). ``interval_eta`` overrides it: a hard-thresholded
        sparse K is indefinite, so its search has to start above -lambda_min(K) (the legacy sparse runs
        used [1, 1e3], examples/CompareVariousNumberOfPoints.py:72)."""
        if plot:
            raise NotImplementedError('plotting is out of scope of the B200 build')
        if self.likelihood_method == 'direct':
            results = DirectLikelihood.maximize_log_likelihood(z, self.X, self.K_mixed)
        elif self.likelihood_method == 'profiled':
            results = ProfileLikelihood.find_log_likelihood_der1_zeros(
                z, self.X, self.K_mixed, [1e-4, 1e+3] if interval_eta is None else list(interval_eta))
        else:
            raise ValueError('"likelihood_method" should be "direct" or "profiled".')
        return results
