"""
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): CPU restatement of the stochastic Lanczos quadrature behind the sparse
MixedCorrelation (imate 'slq', reference call sites gaussian_proc/_mixed_correlation/mixed_correlation.py:138-147,
204-213,263-272 -- dead code there; **parity unpinned**: imate is absent, the estimator follows its documented
scheme). It reproduces, on the host and with plain NumPy / SciPy,
  * the Rademacher probes of the device (counter-based hash of (seed, probe id, row), csrc/gp_sparse_la.cu
    rademacher_kernel), so that the GPU samples can be compared PER PROBE, not only in distribution;
  * m Lanczos steps without re-orthogonalisation and the Gauss quadrature of log, 1/x, 1/x^2.
"""

import numpy
import scipy.linalg

__all__ = ['rademacher', 'lanczos', 'slq_samples']

_M64 = (1 << 64) - 1


def rademacher(n, B, seed, probe0):
    """V[i][c] = +-1 from splitmix-style mixing of (seed, probe0 + c, i): bit-for-bit the device probes."""
    V = numpy.empty((n, B))
    for c in range(B):
        pid = probe0 + c
        for i in range(n):
            z = (seed * 0x9E3779B97F4A7C15 + pid * 0xBF58476D1CE4E5B9 + i * 0x94D049BB133111EB + 0x2545F4914F6CDD1D) & _M64
            z ^= z >> 30
            z = (z * 0xBF58476D1CE4E5B9) & _M64
            z ^= z >> 27
            z = (z * 0x94D049BB133111EB) & _M64
            z ^= z >> 31
            V[i, c] = 1.0 if (z & 1) else -1.0
    return V


def lanczos(A, v, m):
    """alpha (m), beta (m) of m Lanczos steps of the symmetric operator A started at v / ||v||; beta[j] = norm of the
    (j+1)-th unnormalised vector."""
    q = v / numpy.linalg.norm(v)
    qprev = numpy.zeros_like(q)
    bprev = 0.0
    alpha, beta = numpy.zeros(m), numpy.zeros(m)
    for j in range(m):
        w = A @ q
        alpha[j] = q @ w
        w = w - alpha[j] * q - bprev * qprev
        beta[j] = numpy.linalg.norm(w)
        if not beta[j] > 1e-300:
            break
        qprev, q, bprev = q, w / beta[j], beta[j]
    return alpha, beta


def slq_samples(K, eta, seed, probe0, B, m):
    """(B x 3) per-probe estimates n * [v^T log(Kn) v, v^T Kn^-1 v, v^T Kn^-2 v] / ||v||^2 with Kn = K + eta I."""
    n = K.shape[0]
    V = rademacher(n, B, seed, probe0)
    out = numpy.empty((B, 3))

    class Op(object):
        def __matmul__(self, x):
            return K @ x + eta * x

    for c in range(B):
        a, b = lanczos(Op(), V[:, c], m)
        theta, Y = scipy.linalg.eigh_tridiagonal(a, b[:m - 1])
        w = Y[0, :] ** 2
        out[c] = n * numpy.array([numpy.sum(w * numpy.log(theta)), numpy.sum(w / theta), numpy.sum(w / theta ** 2)])
    return out
