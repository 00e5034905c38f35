"""CPU tests of the host-side logic of the product: the root finder against the reference-pinned oracle restatement,
and the (m+1)x(m+1) algebra of _fused.FusedQuantities fed with device outputs simulated in NumPy."""

import numpy
import pytest

from oracle import likelihood as L
from oracle import matern
from oracle import data_utilities as du


def test_bracketing_matches_oracle_on_all_branches():
    from gaussian_proc._likelihood._root_finding import find_interval_with_sign_change as prod
    fns = [lambda x: x - 0.3,                       # immediate sign change
           lambda x: (x - 0.6) ** 2 - 0.01,         # midpoint brackets
           lambda x: (x - 2.5) * 0.1,               # extrapolate right
           lambda x: -(x + 1.7) * 0.2,              # extrapolate left
           lambda x: x * x + 1.0,                   # never
           lambda x: (x - 0.45) ** 2 + 0.01]        # shrink, never
    for f in fns:
        calls_p, calls_o = [], []
        rp = prod(lambda x: (calls_p.append(x), f(x))[1], [0.0, 1.0], 3)
        ro = L.find_interval_with_sign_change(lambda x: (calls_o.append(x), f(x))[1], [0.0, 1.0], 3)
        assert rp[0] == ro[0] and rp[1] == ro[1] and rp[2] == ro[2]
        assert calls_p == calls_o                   # same evaluation sequence = same number of factorizations


def test_chandrupatla_matches_oracle():
    from gaussian_proc._likelihood._root_finding import chandrupatla_method as prod
    fns = [(lambda x: numpy.cos(x) - x, [0.0, 1.0]), (lambda x: x ** 3 - 2 * x - 5, [2.0, 3.0]),
           (lambda x: numpy.tanh(5 * (x - 0.123)), [-4.0, 3.0]), (lambda x: numpy.exp(-x) - 1e-3, [-4.0, 9.0])]
    for f, br in fns:
        vals = [f(br[0]), f(br[1])]
        cp, co = [], []
        rp = prod(lambda x: (cp.append(x), f(x))[1], br, vals, eps_m=1e-6, eps_a=1e-6, maxiter=100)
        ro = L.chandrupatla_method(lambda x: (co.append(x), f(x))[1], br, vals, eps_m=1e-6, eps_a=1e-6, maxiter=100)
        assert cp == co and rp['root'] == ro['root'] and rp['iterations'] == ro['iterations']
        assert abs(f(rp['root'])) < 1e-4
    with pytest.raises(AssertionError):
        prod(lambda x: x * x + 1, [0.0, 1.0], None)
    r = prod(lambda x: x - 0.25, [0.0, 1.0], None, eps_m=1e-12, eps_a=1e-12)
    assert abs(r['root'] - 0.25) < 1e-10


def _simulate_device_out(K, dK, X, z, eta):
    """What gp_loglik_dense returns (include/gpgp.h), computed in NumPy."""
    n, m = X.shape
    p = m + 1
    Kn = K + eta * numpy.eye(n)
    Kinv = numpy.linalg.inv(Kn)
    R = numpy.c_[X, z]
    S = Kinv @ R
    out = numpy.zeros(8 + 4 * p * p)
    out[0] = numpy.linalg.slogdet(Kn)[1]
    out[1] = numpy.trace(Kinv)
    out[2] = numpy.sum(Kinv * Kinv)
    out[3] = numpy.sum(Kinv * dK)
    out[8:8 + p * p] = (R.T @ S).ravel()
    out[8 + p * p:8 + 2 * p * p] = (S.T @ S).ravel()
    out[8 + 2 * p * p:8 + 3 * p * p] = (S.T @ dK @ S).ravel()
    out[8 + 3 * p * p:] = (S.T @ Kinv @ S).ravel()
    return out


class _HostOperator(object):
    """Duck-typed K_mixed for host-only tests of the likelihood algebra: the fused evaluator is simulated in NumPy."""
    sparse = False
    interpolate = False

    def __init__(self, K, dK):
        self.K, self.dK = K, dK

    def dot(self, eta, x, exponent=1):
        return self.K @ x + eta * x

    def trace(self, eta, exponent=1):
        A = self.K + eta * numpy.eye(self.K.shape[0])
        return numpy.trace(A) if exponent == 1 else numpy.sum(A * A)


def test_second_derivatives_from_moments(monkeypatch):
    """Hessian (variance space = the reference's numbers, and sigma space with chain_rule=True), the second
    eta-derivative and the sigma -> 0 limit, all assembled from the moments of one fused evaluation."""
    from gaussian_proc._likelihood import _fused, DirectLikelihood, ProfileLikelihood
    numpy.random.seed(5)
    pts = numpy.random.rand(120, 2)
    z = du.generate_data(pts, 0.2)
    X = du.generate_basis_functions(pts, 2)
    n, m = X.shape
    K = matern.generate_dense_correlation(pts, 0.1, 1.5)
    dK = matern.matern_derivative_rho(pts, 0.1, 1.5)
    op = _HostOperator(K, dK)
    Ko = L.MixedCorrelation(K, 'cholesky')

    def fake_evaluate(z_, X_, K_mixed, eta, traceinv=False, inverse=False, drho=False, cubic=False):
        return _fused.FusedQuantities(_simulate_device_out(K_mixed.K, K_mixed.dK, X_, z_, eta), n, m, eta, 15)
    monkeypatch.setattr(_fused, 'evaluate', fake_evaluate)
    for h in ([0.3, 0.2], [0.7, 0.1], [0.2, 0.5]):
        ref = L.DirectLikelihood.log_likelihood_hessian(z, X, Ko, False, h)
        got = DirectLikelihood.log_likelihood_hessian(z, X, op, False, h)
        assert numpy.max(numpy.abs(got - ref)) <= 1e-8 * numpy.max(numpy.abs(ref))
        assert (DirectLikelihood.log_likelihood_hessian(z, X, op, True, h) == -got).all()
        # sigma-space Hessian = finite difference of the sigma-space gradient of the oracle's l
        gs = lambda x: L.DirectLikelihood.log_likelihood_jacobian(z, X, Ko, False, list(x)) * 2.0 * numpy.asarray(x)  # noqa: E731
        e = 1e-5
        fd = numpy.array([(gs([h[0] + e, h[1]]) - gs([h[0] - e, h[1]])) / (2 * e),
                          (gs([h[0], h[1] + e]) - gs([h[0], h[1] - e])) / (2 * e)])
        hs = DirectLikelihood.log_likelihood_hessian(z, X, op, False, h, chain_rule=True)
        assert numpy.max(numpy.abs(hs - fd)) <= 2e-6 * numpy.max(numpy.abs(fd))
    # second eta-derivative: the reference formula (:183) written with the explicit projected precision
    eta = 0.4
    Kinv = numpy.linalg.inv(K + eta * numpy.eye(n))
    Y = Kinv @ X
    M = Kinv - Y @ numpy.linalg.inv(X.T @ Y) @ Y.T
    Mz = M @ z
    zMz, zM3z = z @ Mz, Mz @ (M @ Mz)
    ref = (0.5 / (zMz / (n - m))) * ((numpy.sum(M * M) / (n - m) + (numpy.trace(M) / (n - m)) ** 2) * zMz - 2.0 * zM3z)
    got = ProfileLikelihood.log_likelihood_der2_eta(z, X, op, eta)
    assert abs(got - ref) <= 1e-9 * abs(ref)
    # sigma -> 0: the closed-form limit against finite differences of l(a, b), Sigma = a K + b I, around a = 0
    def ell(a, b):
        Sg = a * K + b * numpy.eye(n)
        Si = numpy.linalg.inv(Sg)
        Bm = X.T @ Si @ X
        Pz = Si @ z - Si @ X @ numpy.linalg.solve(Bm, X.T @ (Si @ z))
        return -0.5 * (n - m) * numpy.log(2 * numpy.pi) - 0.5 * numpy.linalg.slogdet(Sg)[1] \
            - 0.5 * numpy.linalg.slogdet(Bm)[1] - 0.5 * z @ Pz
    b0, e = 0.09, 2e-4
    g0, h0 = DirectLikelihood._degenerate_derivatives(z, X, op, numpy.sqrt(b0))
    fd_g = numpy.array([(ell(e, b0) - ell(-e, b0)) / (2 * e), (ell(0.0, b0 + e) - ell(0.0, b0 - e)) / (2 * e)])
    fd_h = numpy.array([[(ell(e, b0) - 2 * ell(0.0, b0) + ell(-e, b0)) / e ** 2,
                         (ell(e, b0 + e) - ell(e, b0 - e) - ell(-e, b0 + e) + ell(-e, b0 - e)) / (4 * e * e)],
                        [0.0, (ell(0.0, b0 + e) - 2 * ell(0.0, b0) + ell(0.0, b0 - e)) / e ** 2]])
    fd_h[1, 0] = fd_h[0, 1]
    assert numpy.max(numpy.abs(g0 - fd_g)) <= 1e-5 * numpy.max(numpy.abs(g0))
    assert numpy.max(numpy.abs(h0 - fd_h)) <= 1e-3 * numpy.max(numpy.abs(h0))


def test_fused_algebra_reproduces_reference_formulas():
    from gaussian_proc._likelihood._fused import FusedQuantities
    numpy.random.seed(2)
    pts = numpy.random.rand(150, 2)
    z = du.generate_data(pts, 0.2)
    X = du.generate_basis_functions(pts, 2)
    n, m = X.shape
    K = matern.generate_dense_correlation(pts, 0.1, 1.5)
    dK = matern.matern_derivative_rho(pts, 0.1, 1.5)
    Ko = L.MixedCorrelation(K, 'cholesky')
    for sigma, sigma0 in [(0.3, 0.2), (0.7, 0.1)]:
        eta = (sigma0 / sigma) ** 2
        q = FusedQuantities(_simulate_device_out(K, dK, X, z, eta), n, m, eta, 7)
        # direct likelihood and jacobian, assembled exactly as gaussian_proc/_likelihood/_direct_likelihood.py does
        lp = -0.5 * (n - m) * numpy.log(2 * numpy.pi) - 0.5 * (n * numpy.log(sigma ** 2) + q.logdet_Kn) \
            - 0.5 * numpy.log(numpy.linalg.det(q.B / sigma ** 2)) - 0.5 * q.zMz / sigma ** 2
        assert abs(lp - L.DirectLikelihood.log_likelihood(z, X, Ko, False, [sigma, sigma0])) <= 1e-10 * abs(lp)
        trace_M = q.trace_M / sigma ** 2
        jac = [-0.5 * ((n - m) / sigma ** 2 - eta * trace_M) + 0.5 * q.zMKMz / sigma ** 4,
               -0.5 * trace_M + 0.5 * q.zM2z / sigma ** 4]
        ref = L.DirectLikelihood.log_likelihood_jacobian(z, X, Ko, False, [sigma, sigma0])
        assert numpy.max(numpy.abs(jac - ref)) <= 1e-9 * numpy.max(numpy.abs(ref))
        drho = -0.5 * q.trace_MdK + 0.5 * q.zMdKMz / sigma ** 2
        ref = L.DirectLikelihood.log_likelihood_der1_rho(z, X, Ko, dK, [sigma, sigma0])
        assert abs(drho - ref) <= 1e-9 * abs(ref)
        # profile quantities
        sig2 = q.zMz / (n - m)
        d = -0.5 * (q.trace_M - q.zM2z / sig2)
        assert abs(d - L.ProfileLikelihood.log_likelihood_der1_eta(z, X, Ko, numpy.log10(eta))) <= 1e-9 * abs(d)
        ref = L.ProfileLikelihood.log_likelihood_der1_rho(z, X, Ko, dK, eta)
        assert abs(-0.5 * q.trace_MdK + 0.5 * q.zMdKMz / sig2 - ref) <= 1e-9 * abs(ref)


def test_bench_inputs_equal_reference_generators():
    import bench
    pts, z, X = bench.make_inputs(400)
    numpy.random.seed(0)
    p2 = numpy.random.rand(400, 2)
    assert (pts == p2).all() and (z == du.generate_data(p2, 0.2)).all()
    assert numpy.max(numpy.abs(X - du.generate_basis_functions(p2, 2))) == 0.0


def test_trace_interpolation_rational_polynomial():
    """MixedCorrelation(interpolate=True) (reference mixed_correlation.py:52-66,167-170 -> imate.InterpolateTraceInv):
    the rational-polynomial interpolant reproduces tr (K + eta I)^-1 at the interpolant points exactly and to a few
    1e-3 in between (legacy points of examples/CompareVariousNumberOfPoints.py:68); parity unpinned (imate absent)."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location(
        'itp', os.path.join(root, 'gaussian-process-param-estimation_b200', 'gaussian_proc', '_mixed_correlation',
                            '_interpolate_traceinv.py'))
    itp = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(itp)
    from oracle import matern
    numpy.random.seed(0)
    pts = numpy.random.rand(400, 2)
    lam = numpy.linalg.eigvalsh(matern.generate_dense_correlation(pts, numpy.array([0.1, 0.1]), 0.5))
    calls = []

    def exact(eta):
        calls.append(eta)
        return float(numpy.sum(1.0 / (lam + eta)))

    P = [1.0, 10.0, 40.0, 100.0, 1000.0]
    it = itp.InterpolateTraceInv(exact, 400, P)
    assert sorted(calls) == P                                   # one evaluation per interpolant point
    for t in P:
        assert it.interpolate(t) == exact(t)
    etas = numpy.logspace(0, 3, 40)
    err = numpy.abs(it.interpolate(etas) - numpy.array([exact(e) for e in etas])) / numpy.array([exact(e) for e in etas])
    assert err.max() <= 5e-3
    n_calls = len(calls)
    assert abs(it.interpolate(0.5) - exact(0.5)) == 0.0 and len(calls) == n_calls + 2     # below the anchor: direct
    six = itp.InterpolateTraceInv(exact, 400, [1.0, 5.0, 10.0, 40.0, 100.0, 1000.0])      # q = 5: least squares
    assert abs(six.interpolate(20.0) - exact(20.0)) <= 1e-2 * exact(20.0)
    with pytest.raises(ValueError):
        itp.InterpolateTraceInv(exact, 400, [1.0, 2.0])


def test_lanczos_block_quadrature_and_solution_coefficients():
    """Host part of SLQ (gaussian_proc/_sparse.py::lanczos_block_quadrature) on a NumPy Lanczos run: with m = n steps the
    Gauss quadratures are exact, the coefficients reproduce A^-1 v from the UNNORMALISED Lanczos vectors, the residual
    estimate beta_m |y_m| matches the true residual for m < n, and a shift of the diagonal (T(eta) = T + eta I) serves
    another eta from the same run (shift invariance, mixed_correlation.py:44 AffineMatrixFunction)."""
    from gaussian_proc._sparse import lanczos_block_quadrature, lanczos_quadrature
    rng = numpy.random.RandomState(1)
    n, B = 40, 3
    M = rng.randn(n, n)
    A = M @ M.T / n + 0.5 * numpy.eye(n)
    V = rng.choice([-1.0, 1.0], size=(n, B))

    def lanczos(A, V, m):
        al, be = numpy.zeros((m, V.shape[1])), numpy.zeros((m, V.shape[1]))
        U = numpy.zeros((m, n, V.shape[1]))
        for c in range(V.shape[1]):
            u, uprev, s, sprev, bprev = V[:, c].copy(), numpy.zeros(n), 1.0 / numpy.linalg.norm(V[:, c]), 0.0, 0.0
            for j in range(m):
                U[j, :, c] = u
                w = s * (A @ u)
                al[j, c] = s * (u @ w)
                unext = w - al[j, c] * s * u - bprev * sprev * uprev
                # full reorthogonalisation keeps the m = n run exact
                for i in range(j + 1):
                    q = U[i, :, c] / numpy.linalg.norm(U[i, :, c])
                    unext -= (q @ unext) * q
                be[j, c] = numpy.linalg.norm(unext)
                uprev, sprev, bprev = u, s, be[j, c]
                u, s = unext, (1.0 / be[j, c] if be[j, c] > 1e-300 else 0.0)
        return al, be, U

    al, be, U = lanczos(A, V, n)
    quad, tmin, coef, resid = lanczos_block_quadrature(al, be, numpy.sqrt(float(n)))
    lam = numpy.linalg.eigvalsh(A)
    W = numpy.linalg.eigh(A)[1]
    for c in range(B):
        w2 = (W.T @ V[:, c]) ** 2 / n
        assert abs(quad[c, 0] - numpy.sum(w2 * numpy.log(lam))) <= 1e-10
        assert abs(quad[c, 1] - numpy.sum(w2 / lam)) <= 1e-10
        x = numpy.einsum('j,ji->i', coef[:, c], U[:, :, c])
        assert numpy.max(numpy.abs(x - numpy.linalg.solve(A, V[:, c]))) <= 1e-9
        ref, t, k = lanczos_quadrature(al[:, c], be[:, c], [numpy.log], return_size=True)
        assert abs(ref[0] - quad[c, 0]) <= 1e-12
    # truncated run: residual estimate == true relative residual; shifted tridiagonal == run on A + eta I
    m = 12
    al, be, U = lanczos(A, V, m)
    eta = 0.7
    _, _, coef, resid = lanczos_block_quadrature(al + eta, be, numpy.sqrt(float(n)))
    for c in range(B):
        x = numpy.einsum('j,ji->i', coef[:, c], U[:, :, c])
        true = numpy.linalg.norm(V[:, c] - (A + eta * numpy.eye(n)) @ x) / numpy.linalg.norm(V[:, c])
        assert abs(resid[c] - true) <= 1e-8 * max(true, 1e-3)
    al2, be2, _ = lanczos(A + eta * numpy.eye(n), V, m)
    assert numpy.max(numpy.abs(al2 - (al + eta))) <= 1e-10 and numpy.max(numpy.abs(be2 - be)) <= 1e-10


def test_recorded_bench_line_follows_the_contract():
    """The committed bench record (profiles/r02_bench_1gpu_builder.json, written by `python bench.py --steps 20 --warmup 5` on a B200) carries
    every key of the driver's contract: headline metric, roofline, cpu_baseline, e2e, clocks, launch count."""
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    line = json.loads(open(os.path.join(root, 'profiles', 'r02_bench_1gpu_builder.json')).read().strip().splitlines()[-1])
    for k in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
              'vs_baseline', 'dtype', 'data', 'config', 'roofline', 'cpu_baseline', 'e2e', 'gpu_launches', 'clocks'):
        assert k in line, k
    assert line['unit'] == 'evals/s' and line['dtype'] == 'f64' and line['higher_is_better'] is True
    assert line['vs_baseline'] is None and 'workload' in line['config'] and 'l2' in line['config']
    r = line['roofline']
    assert r['bound'] == 'tensor' and r['unit'] == 'TFLOP/s' and abs(r['frac'] - r['achieved'] / r['peak']) < 1e-12
    assert r['traffic'] is None or r['traffic'] > 0
    assert set(('value', 'unit', 'cores', 'kind', 'sample')) <= set(line['cpu_baseline'])
    assert set(('value', 'unit', 'h2d_bytes_per_step', 'd2h_bytes_per_step')) <= set(line['e2e'])
    assert line['e2e']['h2d_bytes_per_step'] > 0 and line['e2e']['d2h_bytes_per_step'] > 0
    assert line['gpu_launches'] > 0 and not line['clocks']['reasons']
    assert abs(line['value'] - line['steps'] * line['n_gpus'] / (line['ms_per_step'] * line['steps'] * 1e-3)) < 1e-9
    sh = line['secondary']['sharded']                      # the workloads that shard, measured at every N
    assert set(('C3_sweep_n8k', 'C4_sparse_n1M_probe_split', 'C4_sparse_sweep_n1M', 'C5_blockcyclic')) <= set(sh)
    assert 'extrapolated' in line['cpu_baseline']
    ref = json.loads(open(os.path.join(root, 'profiles', 'r02_bench_reference_builder.json')).read().strip().splitlines()[-1])
    assert ref['impl'] == 'reference' and ref['config'] == line['config'] and ref['cpu_baseline']['extrapolated'] is False
    assert ref['ms_per_step'] * ref['steps'] * 1e-3 < 300          # the bounded samples, not the n = 20 000 evaluation


def test_sparse_engine_host_helpers():
    """Pure host helpers of the sparse engine: power-of-two probe chunks, the first chunk a rank evaluates (what the
    side-stream prefetch uses), and the content fingerprint of cached host arrays."""
    from gaussian_proc._sparse import SparseEngine, DEFAULTS
    from gaussian_proc import _device as dev
    assert SparseEngine._chunks(0, 16, 16) == [(0, 16)]
    assert SparseEngine._chunks(5, 13, 8) == [(5, 8), (13, 4), (17, 1)]
    assert SparseEngine._chunks(3, 0, 8) == []
    eng = object.__new__(SparseEngine)
    eng.opt = dict(DEFAULTS, max_num_samples=50, batch=16)
    eng.probe_range = None
    assert eng._first_chunk() == (0, 16)
    eng.probe_range = (1, 2)          # second of two ranks: the round of 16 probes is cut into 8 + 8
    assert eng._first_chunk() == (8, 8)
    eng.opt = dict(DEFAULTS, max_num_samples=10, batch=16)
    eng.probe_range = (1, 4)          # 10 probes over 4 ranks: 3 each, rank 1 takes ids 3, 4, 5 -> first chunk (3, 2)
    assert eng._first_chunk() == (3, 2)
    a = numpy.arange(100.0)
    k1 = dev.host_key(a)
    assert dev.host_key(a) == k1
    for pos in (0, 37, 99):            # an in-place edit of ANY entry changes the key (full-content digest)
        b = a.copy()
        b[pos] = -1.0
        assert dev.host_key(b) != k1
    assert dev.host_key(a.copy()) == k1            # same content, another object: same device copy can be reused
    assert dev.host_key(numpy.zeros((0, 3)))[0] == (0, 3)


def test_row_slab_geometry_covers_every_row_once():
    """gaussian_proc/_slab.py: uniform slabs of 16-row blocks; the ranks' row ranges tile [0, n) (last slab may be short)."""
    from gaussian_proc._slab import slab_geometry
    for n in (16, 17, 3000, 131072, 2 ** 20, 2 ** 20 + 5):
        for world in (1, 2, 3, 4, 8):
            geo = [slab_geometry(n, world, r) for r in range(world)]
            slab = geo[0][0]
            assert slab % 16 == 0 and slab < 2 ** 28 and all(g[0] == slab for g in geo)
            assert geo[0][1] == 0 and geo[-1][2] == n
            assert all(geo[r][2] == geo[r + 1][1] for r in range(world - 1))
            assert all(g[2] - g[1] <= slab for g in geo)
            # owner / local row of a global row, as gp_slab_encode_columns computes it
            rows = numpy.array([0, n // 3, n - 1])
            owner = rows // slab
            assert all(geo[o][1] <= r < geo[o][2] for o, r in zip(owner, rows))


def test_row_block_layout_round_trip_on_the_host():
    """The storage of the row-blocked operator as documented (csrc/gp_sparse_la.cu bval_pos): block rb holds
    operator rows 16 rb .. 16 rb + 15, its block-columns are padded to a multiple of 4, and value (slot, row k) sits at
    (slot >> 2) * 64 + (k >> 3) * 32 + (k & 7) * 4 + (slot & 3). Built here by hand from a small symmetric matrix and read back
    by DeviceRowBlocks.to_scipy (host code; CPU tensors stand in for the device arrays)."""
    import torch
    from gaussian_proc._sparse import DeviceRowBlocks
    rng = numpy.random.RandomState(4)
    n, R = 41, 16
    A = rng.rand(n, n)
    A = numpy.where(A + A.T > 1.5, A + A.T, 0.0)
    A[numpy.arange(n), numpy.arange(n)] = 1.0
    order = rng.permutation(n).astype(numpy.int32)              # operator row r is original row order[r]
    inv = numpy.empty(n, dtype=numpy.int32)
    inv[order] = numpy.arange(n, dtype=numpy.int32)
    P = A[numpy.ix_(order, order)]                               # the spatially ordered matrix
    nrb = (n + R - 1) // R
    bptr, bidx, vals = [0], [], []
    for rb in range(nrb):
        rows = numpy.arange(rb * R, min(n, rb * R + R))
        cols = numpy.nonzero(numpy.any(P[rows] != 0.0, axis=0))[0]
        cols = rng.permutation(cols)                             # any order of the block-columns is valid
        padded = (len(cols) + 3) // 4 * 4
        blockvals = numpy.zeros(padded * R)
        for slot, c in enumerate(cols):
            for k, r in enumerate(rows):
                blockvals[(slot >> 2) * 64 + (k >> 3) * 32 + (k & 7) * 4 + (slot & 3)] = P[r, c]
        bidx += list(cols) + [int(cols[0])] * (padded - len(cols))
        vals.append(blockvals)
        bptr.append(bptr[-1] + padded)
    K = DeviceRowBlocks(n, torch.tensor(bptr, dtype=torch.int64), torch.tensor(bidx, dtype=torch.int32),
                        torch.from_numpy(numpy.concatenate(vals)), None, torch.from_numpy(order), torch.from_numpy(inv),
                        int((A != 0).sum()), 0.0, 0, n)
    S = K.to_scipy()
    assert S.nnz == K.nnz and (S.toarray() == A).all()
