"""GaussianProcess -- API shell, same as the reference's gaussian_proc/gaussian_process/gaussian_process.py:21-59."""

from .._likelihood import Likelihood

__all__ = ['GaussianProcess']


class GaussianProcess(object):
    """Gaussian process for regression: ``GaussianProcess(X, K, likelihood_method='direct').train(z)``.
    ``K`` may be a NumPy array / SciPy CSR matrix (copied to the GPU) or a device handle returned by
    ``generate_correlation(..., device=True)``."""

    def __init__(self, X, K, likelihood_method='direct', imate_method=None, imate_options={}):
        self.X = X
        self.K = K
        self.likelihood = Likelihood(X, K, likelihood_method=likelihood_method, imate_method=imate_method,
                                     imate_options=imate_options)
        self.results = None

    def train(self, z, plot=False, interval_eta=None):
        """Finds the hyperparameters; prints the result dict like the reference (:52-59) and also returns it.
        ``interval_eta``: search interval of the profiled method (default: the reference's [1e-4, 1e3])."""
        results = self.likelihood.maximize_log_likelihood(z, plot=plot, interval_eta=interval_eta)
        self.results = results
        print(results)
        return results
