"""TEST INFRASTRUCTURE. CPU oracle for generate_correlation (dense and kernel-threshold sparse).

Follows gaussian_proc/generate_correlation/generate_correlation.py:32-222 (dispatch, scalar scale -> array :191-196),
_kernels.pyx:17-136, _generate_dense_correlation.pyx:23-162 and _generate_sparse_correlation.pyx:208-594 (with the two
call-site fixes of SURVEY Q7). Closed-form nu run through oracle/cmatern.c (bit-identical to the compiled reference);
general nu uses scipy.special.kv / gamma exactly as _kernels.pyx:87-88 does.
"""

import ctypes
import os
import subprocess

import numpy
import scipy.sparse
import scipy.special

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, '_build')
_SO = os.path.join(_BUILD, 'liboracle.so')
_lib = None


def build_c(force=False):
    """gcc -O2 -ffp-contract=off -fopenmp-free build of cmatern.c into oracle/_build/liboracle.so."""
    src = os.path.join(_HERE, 'cmatern.c')
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        os.makedirs(_BUILD, exist_ok=True)
        subprocess.check_call(['gcc', '-O2', '-ffp-contract=off', '-fPIC', '-shared', '-o', _SO, src, '-lm'])
    return _SO


def _c():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(build_c())
        lib.oracle_matern_kernel.restype = ctypes.c_double
        lib.oracle_matern_kernel.argtypes = [ctypes.c_double, ctypes.c_double]
        lib.oracle_dense.restype = None
        lib.oracle_dense.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_double,
                                     ctypes.c_void_p]
        lib.oracle_sparse_rows.restype = None
        lib.oracle_sparse_rows.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p,
                                           ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_int64,
                                           ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                           ctypes.c_void_p]
        _lib = lib
    return _lib


def _closed_form(nu):
    return nu in (0.5, 1.5, 2.5) or nu >= 100


def matern_kernel(x, nu):
    """Scalar Matern correlation of a scaled distance, _kernels.pyx:17-100."""
    x = float(x)
    if x == 0:
        return 1.0
    if _closed_form(nu):
        return _c().oracle_matern_kernel(x, float(nu))
    return float((2.0 ** (1.0 - nu)) / scipy.special.gamma(nu) * ((numpy.sqrt(2.0 * nu) * x) ** nu)
                 * scipy.special.kv(nu, numpy.sqrt(2.0 * nu) * x))


def scaled_distance_matrix(points, correlation_scale):
    """sqrt(sum_k ((p_ik - p_jk) / scale_k)^2), divide-then-square, k-ordered sum (_kernels.pyx:130-136)."""
    n, d = points.shape
    s = numpy.zeros((n, n))
    for k in range(d):
        s += ((points[:, None, k] - points[None, :, k]) / correlation_scale[k]) ** 2
    return numpy.sqrt(s)


def _scale_array(points, correlation_scale):
    if numpy.isscalar(correlation_scale):
        return numpy.repeat(numpy.array([correlation_scale], dtype=float), points.shape[1])
    return numpy.ascontiguousarray(correlation_scale, dtype=float)


def generate_dense_correlation(points, correlation_scale, nu):
    points = numpy.ascontiguousarray(points, dtype=float)
    scale = _scale_array(points, correlation_scale)
    n, d = points.shape
    if _closed_form(nu):
        K = numpy.zeros((n, n))
        _c().oracle_dense(points.ctypes.data, n, d, scale.ctypes.data, float(nu), K.ctypes.data)
        return K
    x = scaled_distance_matrix(points, scale)
    y = numpy.sqrt(2.0 * nu) * x
    with numpy.errstate(invalid='ignore', over='ignore'):
        K = (2.0 ** (1.0 - nu)) / scipy.special.gamma(nu) * (y ** nu) * scipy.special.kv(nu, y)
    K[x == 0] = 1.0
    return K


def matern_derivative_rho(points, rho, nu):
    """dK/d(rho) for an isotropic scale (extension; SURVEY 8a A9 closed forms, general nu via K_{nu-1})."""
    x = scaled_distance_matrix(points, numpy.repeat(float(rho), points.shape[1]))
    if nu == 0.5:
        return x / rho * numpy.exp(-x)
    if nu == 1.5:
        return 3.0 * x ** 2 / rho * numpy.exp(-numpy.sqrt(3.0) * x)
    if nu == 2.5:
        return 5.0 * x ** 2 / (3.0 * rho) * (1.0 + numpy.sqrt(5.0) * x) * numpy.exp(-numpy.sqrt(5.0) * x)
    if nu >= 100:
        return x ** 2 / rho * numpy.exp(-0.5 * x ** 2)
    y = numpy.sqrt(2.0 * nu) * x
    with numpy.errstate(invalid='ignore', over='ignore'):
        dK = (2.0 ** (1.0 - nu)) / scipy.special.gamma(nu) * (y ** (nu + 1.0)) * scipy.special.kv(nu - 1.0, y) / rho
    dK[x == 0] = 0.0
    return dK


def matern_derivative_scale(points, scale, nu, k):
    """dK/d(scale[k]) for one correlation scale per dimension (extension of the reference kernel, _kernels.pyx:107-136:
    x = sqrt(sum_j ((p_j - q_j) / scale_j)^2)). With g(x) = -x K'(x) (so that the isotropic derivative is g(x) / rho):
    dK/d scale_k = g(x) u_k^2 / (x^2 scale_k), u_k = (p_k - q_k) / scale_k. Pinned by finite differences of the pinned
    generator (tests/test_oracle_golden.py)."""
    scale = numpy.asarray(scale, dtype=float)
    x = scaled_distance_matrix(points, scale)
    u = (points[:, None, k] - points[None, :, k]) / scale[k]
    if nu == 0.5:
        g = x * numpy.exp(-x)
    elif nu == 1.5:
        g = 3.0 * x ** 2 * numpy.exp(-numpy.sqrt(3.0) * x)
    elif nu == 2.5:
        g = 5.0 * x ** 2 / 3.0 * (1.0 + numpy.sqrt(5.0) * x) * numpy.exp(-numpy.sqrt(5.0) * x)
    elif nu >= 100:
        g = x ** 2 * numpy.exp(-0.5 * x ** 2)
    else:
        y = numpy.sqrt(2.0 * nu) * x
        with numpy.errstate(invalid='ignore', over='ignore'):
            g = (2.0 ** (1.0 - nu)) / scipy.special.gamma(nu) * (y ** (nu + 1.0)) * scipy.special.kv(nu - 1.0, y)
    with numpy.errstate(invalid='ignore', divide='ignore'):
        dK = g * u ** 2 / (x ** 2 * scale[k])
    dK[x == 0] = 0.0
    return dK


# ---- sparse ------------------------------------------------------------------------------------------------

def gamma_function(dimension):
    """Gamma(d/2 + 1), _generate_sparse_correlation.pyx:208-233."""
    if dimension % 2 == 0:
        k = 0.5 * dimension
        g = 1.0
        while k > 0.0:
            g *= k
            k -= 1.0
    else:
        k = numpy.ceil(0.5 * dimension)
        g = numpy.sqrt(numpy.pi)
        while k > 0.0:
            g *= k - 0.5
            k -= 1.0
    return g


def ball_radius(volume, dimension):
    """:240-260"""
    return (gamma_function(dimension) * volume) ** (1.0 / dimension) / numpy.sqrt(numpy.pi)


def ball_volume(radius, dimension):
    """:267-287"""
    return (radius * numpy.sqrt(numpy.pi)) ** dimension / gamma_function(dimension)


def estimate_kernel_threshold(matrix_size, dimension, density, correlation_scale, nu):
    """tau = matern(kernel_radius), _generate_sparse_correlation.pyx:294-413 (with _ball_volume given its dimension)."""
    adjacency_volume = density * matrix_size
    if adjacency_volume < 1.0:
        raise ValueError('Adjacency: %0.2f. Correlation matrix will become identity since kernel radius is less '
                         'than grid size.' % adjacency_volume)
    geometric_mean_radius = numpy.prod(correlation_scale) ** (1.0 / dimension)
    adjacency_volume /= ball_volume(geometric_mean_radius, dimension)
    adjacency_radius = ball_radius(adjacency_volume, dimension)
    grid_size = 1.0 / (matrix_size ** (1.0 / dimension) - 1.0)
    return matern_kernel(grid_size * adjacency_radius, nu)


def generate_sparse_correlation(points, correlation_scale, nu, density, kernel_threshold=None):
    """Canonical CSR (int32 indices / indptr, float64 data) of {K_ij : K_ij > tau}, :472-594."""
    points = numpy.ascontiguousarray(points, dtype=float)
    scale = _scale_array(points, correlation_scale)
    n, d = points.shape
    tau = estimate_kernel_threshold(n, d, density, scale, nu) if kernel_threshold is None else kernel_threshold
    if _closed_form(nu):
        lib = _c()
        counts = numpy.zeros(n, dtype=numpy.int64)
        lib.oracle_sparse_rows(points.ctypes.data, n, d, scale.ctypes.data, float(nu), tau, 0, 0, n, counts.ctypes.data,
                               None, None, None)
        indptr = numpy.zeros(n + 1, dtype=numpy.int64)
        numpy.cumsum(counts, out=indptr[1:])
        indices = numpy.zeros(indptr[-1], dtype=numpy.int32)
        data = numpy.zeros(indptr[-1])
        lib.oracle_sparse_rows(points.ctypes.data, n, d, scale.ctypes.data, float(nu), tau, 1, 0, n, None,
                               indptr.ctypes.data, indices.ctypes.data, data.ctypes.data)
        return scipy.sparse.csr_matrix((data, indices, indptr.astype(numpy.int32)), shape=(n, n))
    K = generate_dense_correlation(points, scale, nu)
    return scipy.sparse.csr_matrix(numpy.where(K > tau, K, 0.0))


def generate_correlation(points, correlation_scale=0.1, nu=0.5, grid=True, sparse=False, density=0.001):
    """generate_correlation.py:32-222 minus plotting; `grid` is accepted and unused like in the reference."""
    if sparse:
        return generate_sparse_correlation(points, correlation_scale, nu, density)
    return generate_dense_correlation(points, correlation_scale, nu)
