"""
gaussian_proc -- B200 (sm_100a) build of the Gaussian-process log-likelihood hot path.

Same public surface as the reference package (gaussian_proc/__init__.py:72-75 exports exactly these two names);
the compute runs in hand-written CUDA kernels (libgpgp.so) through ctypes. There is no CPU fallback.
"""

from .generate_correlation import generate_correlation
from .gaussian_process import GaussianProcess

__all__ = ['generate_correlation', 'GaussianProcess']
__version__ = '0.1.0'
