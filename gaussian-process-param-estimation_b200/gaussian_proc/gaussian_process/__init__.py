from .gaussian_process import GaussianProcess

__all__ = ['GaussianProcess']
