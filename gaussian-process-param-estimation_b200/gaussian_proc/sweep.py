"""
Hyper-parameter grid sweep (BASELINE.json configs[2]; the workload of the reference's legacy
examples/FindOptimalCovarianceParameters.py:632-702, a multiprocessing.Pool over independent cells): profile
log-likelihood and its derivatives with respect to eta and the correlation scale rho on a (rho x eta) grid.

Cells are grouped by rho; each rank owns a contiguous range of rho values, generates K(rho) once on its GPU and
loops over eta with one Cholesky each. No data-path collective; the per-cell results are all-gathered at the end.
"""

import numpy

from . import _distributed as gpd

__all__ = ['likelihood_grid']


def _gpu_cell_evaluator(points, z, X, nu):
    from .generate_correlation.generate_correlation import generate_dense_correlation
    from ._mixed_correlation import MixedCorrelation
    from ._likelihood import ProfileLikelihood
    points = numpy.ascontiguousarray(points, dtype=float)
    state = {}

    def evaluate(rho, eta):
        if state.get('rho') != rho:
            state.clear()      # release the previous matrix before allocating the next one
            K = generate_dense_correlation(points, numpy.repeat(float(rho), points.shape[1]), float(nu))
            state.update(rho=rho, K_mixed=MixedCorrelation(K))
        return ProfileLikelihood.log_likelihood_and_gradient(z, X, state['K_mixed'], eta)
    return evaluate


def likelihood_grid(points, z, X, nu, rhos, etas, evaluate=None):
    """Returns an array (len(rhos), len(etas), 3) with [l^(sigma_hat, eta), d l^/d eta, d l^/d rho] per cell, identical
    on every rank. `evaluate(rho, eta)` may be injected (tests); by default it is the fused GPU evaluator."""
    rhos = numpy.asarray(rhos, dtype=float)
    etas = numpy.asarray(etas, dtype=float)
    rank, world = gpd.rank_world()
    begin, end = gpd.partition_cells(len(rhos), world, rank)
    if evaluate is None:
        evaluate = _gpu_cell_evaluator(points, z, X, nu)
    local = numpy.empty(((end - begin) * len(etas), 5))
    k = 0
    for i in range(begin, end):
        for j, eta in enumerate(etas):
            local[k, :2] = (i, j)
            local[k, 2:] = evaluate(rhos[i], eta)
            k += 1
    rows = gpd.allgather_rows(local)
    out = numpy.full((len(rhos), len(etas), 3), numpy.nan)
    out[rows[:, 0].astype(int), rows[:, 1].astype(int)] = rows[:, 2:]
    return out
