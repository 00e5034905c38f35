"""
generate_correlation -- same signature and dispatch as the reference's
gaussian_proc/generate_correlation/generate_correlation.py:32-222 (plotting, :229-281, is out of scope), with the
dense and sparse generators of _generate_dense_correlation.pyx / _generate_sparse_correlation.pyx replaced by CUDA
kernels (csrc/gp_matern.cu, csrc/gp_sparse.cu).
"""

import ctypes

import numpy

from .. import _device as dev
from .._device import lib, check
from .._dense import DeviceCorrelation

__all__ = ['generate_correlation']


def generate_correlation(
        points,
        correlation_scale=0.1,
        nu=0.5,
        grid=True,
        sparse=False,
        density=0.001,
        plot=False,
        verbose=False,
        device=False):
    """
    Matern correlation matrix of a set of points.

    Arguments are those of the reference (generate_correlation.py:32-40); ``grid`` is accepted and unused, exactly
    like there; ``plot=True`` raises (plotting is out of scope). The extra ``device`` flag keeps the
    result in GPU memory: a :class:`DeviceCorrelation` (dense) or :class:`DeviceCSR` (sparse) handle that
    ``GaussianProcess`` / ``MixedCorrelation`` accept directly. With ``device=False`` the return types are the
    reference's: C-contiguous float64 ``numpy.ndarray`` or canonical ``scipy.sparse.csr_matrix`` (int32 indices).
    """
    if plot:
        raise NotImplementedError('plotting is out of scope of the B200 build (reference: generate_correlation.py:229)')
    points = numpy.ascontiguousarray(points, dtype=numpy.float64)
    if points.ndim != 2:
        raise ValueError('"points" should be a 2D array.')
    dimension = points.shape[1]

    # scalar scale -> one entry per dimension (generate_correlation.py:191-196)
    if numpy.isscalar(correlation_scale):
        correlation_scale = numpy.repeat(numpy.array([correlation_scale], dtype=float), dimension)
    correlation_scale = numpy.ascontiguousarray(correlation_scale, dtype=numpy.float64)
    if correlation_scale.size != dimension:
        raise ValueError('"correlation_scale" should be a scalar or have one entry per dimension.')

    if sparse:
        from .._sparse import generate_sparse_correlation
        return generate_sparse_correlation(points, correlation_scale, float(nu), float(density), verbose, device)
    K = generate_dense_correlation(points, correlation_scale, float(nu), verbose)
    return K if device else K.to_numpy()


def generate_dense_correlation(points, correlation_scale, nu, verbose=False, with_derivative=False):
    """Device generator behind the reference's generate_dense_correlation (_generate_dense_correlation.pyx:98-162).
    Returns a DeviceCorrelation (and dK/d rho as a second padded tensor when ``with_derivative``)."""
    torch = dev.require_cuda()
    n, d = points.shape
    npad = dev.padded_size(n)
    dpts = torch.from_numpy(points).cuda()
    K = torch.empty((npad, npad), dtype=torch.float64, device='cuda')
    dK = torch.empty((npad, npad), dtype=torch.float64, device='cuda') if with_derivative else None
    rc = lib.gp_matern_dense(ctypes.c_void_p(dpts.data_ptr()), n, d, dev.host_ptr(correlation_scale), nu,
                             ctypes.c_void_p(K.data_ptr()), npad,
                             ctypes.c_void_p(dK.data_ptr()) if dK is not None else None, dev.stream_ptr())
    check(rc, 'gp_matern_dense')
    if verbose:
        print('Generated dense correlation matrix of size: %d.' % n)
    out = DeviceCorrelation(n, K, points=dpts, correlation_scale=correlation_scale, nu=nu)
    return (out, dK) if with_derivative else out
