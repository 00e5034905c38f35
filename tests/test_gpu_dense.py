"""GPU parity tests (dense path): CUDA product through the C ABI / Python mirror vs the CPU oracle and the committed
golden fixtures. Tolerances: generator <= 4 ulp-level absolute (values in [0,1]); log-likelihood, derivatives,
logdet, traces, solves: 1e-9 relative (BASELINE.json north_star), written next to each assert."""

import numpy
import pytest

pytestmark = pytest.mark.gpu

RTOL = 1e-9


def rel(a, b):
    a, b = numpy.asarray(a, dtype=float), numpy.asarray(b, dtype=float)
    return float(numpy.max(numpy.abs(a - b)) / max(numpy.max(numpy.abs(b)), 1e-300))


@pytest.fixture(scope='module')
def gp():
    import gaussian_proc
    return gaussian_proc


@pytest.fixture(scope='module')
def problem():
    from oracle import data_utilities as du
    numpy.random.seed(0)
    pts = numpy.random.rand(300, 2)
    return pts, du.generate_data(pts, 0.2), du.generate_basis_functions(pts, 2)


# ---------------------------------------------------------------------------------------------- generator
@pytest.mark.parametrize('nu', [0.5, 1.5, 2.5, 200.0, 3.3, 0.8])
def test_dense_generator_matches_reference_golden(gp, golden_generate, nu):
    pts = golden_generate['points2d']
    K = gp.generate_correlation(pts, numpy.array([0.1, 0.17]), nu)
    Kref = golden_generate['dense_nu%g' % nu]
    assert K.shape == Kref.shape and K.dtype == numpy.float64 and K.flags['C_CONTIGUOUS']
    tol = 4e-16 if nu in (0.5, 1.5, 2.5, 200.0) else 2e-13   # closed forms: ~ulp; Bessel branch: device K_nu
    assert numpy.max(numpy.abs(K - Kref)) <= tol
    assert (K == K.T).all()                      # exact symmetry (reference mirrors one evaluation, Q13)
    assert (numpy.diag(K) == 1.0).all()          # x == 0 -> exactly 1 (_kernels.pyx:73-74)


def test_dense_generator_3d_and_scalar_scale(gp, golden_generate):
    K = gp.generate_correlation(golden_generate['points3d'], numpy.array([0.2, 0.3, 0.25]), 1.5)
    assert numpy.max(numpy.abs(K - golden_generate['dense3d_nu1.5'])) <= 4e-16
    from oracle import matern
    pts = golden_generate['points2d']
    K1 = gp.generate_correlation(pts, 0.1, 2.5)        # scalar -> repeated (generate_correlation.py:191-196)
    assert numpy.max(numpy.abs(K1 - matern.generate_dense_correlation(pts, 0.1, 2.5))) <= 4e-16


def test_dense_generator_edge_sizes(gp):
    from oracle import matern
    for n in (1, 2, 127, 128, 129, 257):
        numpy.random.seed(n)
        pts = numpy.random.rand(n, 2)
        K = gp.generate_correlation(pts, 0.2, 1.5)
        assert K.shape == (n, n)
        assert numpy.max(numpy.abs(K - matern.generate_dense_correlation(pts, 0.2, 1.5))) <= 4e-16
    # duplicate points: distance exactly zero off the diagonal -> exactly 1
    pts = numpy.array([[0.1, 0.2], [0.1, 0.2], [0.5, 0.5]])
    K = gp.generate_correlation(pts, 0.2, 2.5)
    assert K[0, 1] == 1.0 and K[1, 0] == 1.0


def test_generator_dK_matches_oracle(gp, problem):
    from gaussian_proc.generate_correlation.generate_correlation import generate_dense_correlation
    from oracle import matern
    pts = problem[0]
    for nu in (0.5, 1.5, 2.5, 200.0, 3.3):
        Kd, dK = generate_dense_correlation(pts, numpy.array([0.1, 0.1]), nu, with_derivative=True)
        n = pts.shape[0]
        ref = matern.matern_derivative_rho(pts, 0.1, nu)
        assert rel(dK[:n, :n].cpu().numpy(), ref) <= 1e-11, nu


# ------------------------------------------------------------------------------- mixed correlation operator
@pytest.mark.parametrize('case', [0, 1, 2, 3, 4])
def test_mixed_correlation_against_reference_vectors(gp, golden_likelihood, case):
    from gaussian_proc._mixed_correlation import MixedCorrelation
    g = golden_likelihood
    nu, rho = g['cases'][case]
    pts, z, X = g['points'], g['z'], g['X']
    K = gp.generate_correlation(pts, rho, nu, device=True)
    Km = MixedCorrelation(K, imate_method='cholesky')
    assert Km.get_matrix_size() == pts.shape[0]
    for method in ('eigenvalue', 'cholesky'):          # the reference's two deterministic methods agree
        tag = 'c%d_%s_' % (case, method)
        for i, t in enumerate(g['log_etas']):
            eta = 10.0 ** t
            if case == 4 and t < 0:
                continue                                # Gaussian kernel, tiny eta: kappa ~ 1e12, outside the 1e-9 regime
            assert abs(Km.logdet(eta) - g[tag + 'logdet'][i]) <= RTOL * abs(g[tag + 'logdet'][i]) + 1e-9
            assert rel(Km.traceinv(eta), g[tag + 'traceinv'][i]) <= (1e-8 if t < 0 else RTOL)
            assert rel(Km.traceinv(eta, exponent=2), g[tag + 'traceinv2'][i]) <= (1e-7 if t < 0 else RTOL)
    if case != 4:
        sol = Km.solve(0.1, numpy.c_[X, z])
        assert rel(sol, g['c%d_solve_eta0.1' % case]) <= RTOL
        assert rel(Km.solve(0.1, z), g['c%d_solve_eta0.1' % case][:, -1]) <= RTOL
    # dot / trace against the oracle
    from oracle import matern
    Kh = matern.generate_dense_correlation(pts, rho, nu)
    assert rel(Km.dot(0, z), Kh @ z) <= 1e-12
    assert rel(Km.dot(0.3, X), Kh @ X + 0.3 * X) <= 1e-12
    assert rel(Km.trace(0.5), numpy.trace(Kh) + 0.5 * Kh.shape[0]) <= 1e-13
    assert rel(Km.trace(0.5, exponent=2), numpy.sum((Kh + 0.5 * numpy.eye(Kh.shape[0])) ** 2)) <= 1e-12


def test_not_positive_definite_raises(gp):
    from gaussian_proc._mixed_correlation import MixedCorrelation
    A = numpy.eye(200)
    A[5, 5] = -1.0
    Km = MixedCorrelation(A)
    with pytest.raises(numpy.linalg.LinAlgError):
        Km.logdet(0.0)


# ---------------------------------------------------------------------------------------------- likelihood
@pytest.mark.parametrize('case', [0, 1, 2, 3])
def test_likelihood_against_reference_vectors(gp, golden_likelihood, case):
    from gaussian_proc._mixed_correlation import MixedCorrelation
    from gaussian_proc._likelihood import DirectLikelihood, ProfileLikelihood
    g = golden_likelihood
    nu, rho = g['cases'][case]
    pts, z, X = g['points'], g['z'], g['X']
    Km = MixedCorrelation(gp.generate_correlation(pts, rho, nu, device=True))
    tag = 'c%d_cholesky_' % case
    for i, h in enumerate(g['hyper_direct']):
        eta = (h[1] / h[0]) ** 2
        tol = RTOL if eta >= 1e-2 else 1e-7       # agreement between two FP64 algorithms ~ kappa*eps
        ll = DirectLikelihood.log_likelihood(z, X, Km, False, list(h))
        assert abs(ll - g[tag + 'direct_ll'][i]) <= tol * abs(g[tag + 'direct_ll'][i])
        assert DirectLikelihood.log_likelihood(z, X, Km, True, list(h)) == -ll
        jac = DirectLikelihood.log_likelihood_jacobian(z, X, Km, False, list(h))
        assert rel(jac, g[tag + 'direct_jac'][i]) <= tol
        hess = DirectLikelihood.log_likelihood_hessian(z, X, Km, False, list(h))
        assert rel(hess, g[tag + 'direct_hess'][i]) <= max(tol, 1e-8)
    for i, h in enumerate(g['hyper_profile']):
        tol = RTOL if h[1] >= 1e-2 else 1e-7
        ll = ProfileLikelihood.log_likelihood(z, X, Km, False, list(h))
        assert abs(ll - g[tag + 'profile_ll'][i]) <= tol * abs(g[tag + 'profile_ll'][i])
    for i, t in enumerate(g['log_etas']):
        d = ProfileLikelihood.log_likelihood_der1_eta(z, X, Km, t)
        ref = g[tag + 'profile_der1'][i]
        assert abs(d - ref) <= (RTOL if t >= -1 else 1e-7) * max(abs(ref), 1.0)


def test_profile_root_against_reference(gp, golden_likelihood):
    from gaussian_proc._mixed_correlation import MixedCorrelation
    from gaussian_proc._likelihood import ProfileLikelihood
    g = golden_likelihood
    pts, z, X = g['points'], g['z'], g['X']
    for case in (0, 1, 2):
        ref = g['c%d_eigenvalue_root' % case]
        if numpy.isnan(ref).any():
            continue
        nu, rho = g['cases'][case]
        Km = MixedCorrelation(gp.generate_correlation(pts, rho, nu, device=True))
        res = ProfileLikelihood.find_log_likelihood_der1_zeros(z, X, Km, [1e-4, 1e3])
        got = numpy.array([res['sigma'], res['sigma0'], res['eta']])
        assert rel(got, ref) <= 1e-7       # root tolerance is 1e-6 in log10(eta)


def test_golden_pickle_cells_on_gpu(gp, golden_pickles):
    """Shipped OptimalCovariance_WithoutPrior.pickle cells (general-nu Bessel branch, n = 900 grid): full GPU path
    (device K_nu generator -> Cholesky -> root find -> profile l)."""
    from oracle import data_utilities as du
    from gaussian_proc._mixed_correlation import MixedCorrelation
    from gaussian_proc._likelihood import ProfileLikelihood
    pts = du.generate_points(30, 2, grid=True)
    z = du.generate_data(pts, 0.2)
    X = du.generate_basis_functions(pts, 2)
    for (i, j) in [(0, 0), (10, 7), (33, 20), (60, 59), (45, 3)]:
        rho, nu = golden_pickles['rho'][i], golden_pickles['nu'][j]
        Km = MixedCorrelation(gp.generate_correlation(pts, rho, nu, device=True))
        res = ProfileLikelihood.find_log_likelihood_der1_zeros(z, X, Km, [1e-3, 1e3])
        lp = ProfileLikelihood.log_likelihood(z, X, Km, False, [res['sigma'], res['eta']])
        assert abs(lp - golden_pickles['Lp_noprior'][i, j]) <= 1e-9 * abs(lp)


# ------------------------------------------------------------------------------- d/d rho (extension) + fused
@pytest.mark.parametrize('nu', [0.5, 1.5, 2.5])
def test_gradient_rho_matches_oracle(gp, problem, nu):
    from oracle import matern, likelihood as L
    from gaussian_proc._mixed_correlation import MixedCorrelation
    from gaussian_proc._likelihood import DirectLikelihood, ProfileLikelihood
    pts, z, X = problem
    rho = 0.1
    Kh = matern.generate_dense_correlation(pts, rho, nu)
    dKh = matern.matern_derivative_rho(pts, rho, nu)
    Ko = L.MixedCorrelation(Kh, 'cholesky')
    Km = MixedCorrelation(gp.generate_correlation(pts, rho, nu, device=True))
    for h in [(0.3, 0.2), (0.5, 0.5)]:
        lp, jac, drho = DirectLikelihood.log_likelihood_and_gradient(z, X, Km, list(h))
        assert abs(lp - L.DirectLikelihood.log_likelihood(z, X, Ko, False, list(h))) <= RTOL * abs(lp)
        assert rel(jac, L.DirectLikelihood.log_likelihood_jacobian(z, X, Ko, False, list(h))) <= RTOL
        assert rel(drho, L.DirectLikelihood.log_likelihood_der1_rho(z, X, Ko, dKh, list(h))) <= RTOL
    for eta in (0.05, 1.0, 20.0):
        lp, deta, drho = ProfileLikelihood.log_likelihood_and_gradient(z, X, Km, eta)
        sig = L.ProfileLikelihood.find_optimal_sigma(z, X, Ko, eta)
        assert abs(lp - L.ProfileLikelihood.log_likelihood(z, X, Ko, False, [sig, eta])) <= RTOL * abs(lp)
        assert rel(deta, L.ProfileLikelihood.log_likelihood_der1_eta(z, X, Ko, numpy.log10(eta))) <= RTOL
        assert rel(drho, L.ProfileLikelihood.log_likelihood_der1_rho(z, X, Ko, dKh, eta)) <= RTOL


def test_gradient_rho_set_kernel_and_errors(gp, problem):
    from gaussian_proc._mixed_correlation import MixedCorrelation
    from gaussian_proc._likelihood import ProfileLikelihood
    pts, z, X = problem
    Kh = gp.generate_correlation(pts, 0.1, 1.5)              # plain ndarray: generator parameters unknown
    Km = MixedCorrelation(Kh)
    with pytest.raises(ValueError):
        ProfileLikelihood.log_likelihood_der1_rho(z, X, Km, 0.5)
    Km.set_kernel(pts, 0.1, 1.5)
    a = ProfileLikelihood.log_likelihood_der1_rho(z, X, Km, 0.5)
    Kd = MixedCorrelation(gp.generate_correlation(pts, 0.1, 1.5, device=True))
    assert rel(a, ProfileLikelihood.log_likelihood_der1_rho(z, X, Kd, 0.5)) <= 1e-12


def test_gaussian_process_train_profiled(gp, golden_likelihood):
    g = golden_likelihood
    pts, z, X = g['points'], g['z'], g['X']
    nu, rho = g['cases'][1]
    K = gp.generate_correlation(pts, rho, nu)
    res = gp.GaussianProcess(X, K, likelihood_method='profiled').train(z)
    ref = g['c1_eigenvalue_root']
    assert rel([res['sigma'], res['sigma0'], res['eta']], ref) <= 1e-7


# ------------------------------------------------------------------ full size (BASELINE config 2): properties
def test_full_size_n20k_properties(gp):
    """n = 20 000, nu = 2.5, rho = 0.1 (BASELINE.json configs[1]): size-independent identities instead of the oracle.
    (a) solve residual ||Kn x - b|| / ||b||, (b) tr(Kn^-1 Kn) = n via tr Kn^-1 K = n - eta tr Kn^-1 with the SpMV-free
    identity z^T M K M z + eta z^T M^2 z = z^T M z, (c) d l/d eta against a central difference of l."""
    from oracle import data_utilities as du
    from gaussian_proc._mixed_correlation import MixedCorrelation
    from gaussian_proc._likelihood import ProfileLikelihood, _fused
    n = 20000
    numpy.random.seed(0)
    pts = numpy.random.rand(n, 2)
    z = du.generate_data(pts, 0.2)
    X = du.generate_basis_functions(pts, 2)
    Km = MixedCorrelation(gp.generate_correlation(pts, 0.1, 2.5, device=True))
    eta = 0.1
    x = Km.solve(eta, z)
    r = Km.dot(eta, x) - z
    assert numpy.linalg.norm(r) / numpy.linalg.norm(z) <= 1e-10
    q = _fused.evaluate(z, X, Km, eta, traceinv=True, drho=True)
    assert abs(q.zMKMz + eta * q.zM2z - q.zMz) <= 1e-9 * abs(q.zMz)        # M Kn M = M
    assert q.trace_Kninv > 0 and q.trace_Kninv2 > 0
    assert abs(q.trace_Kninv - Km.traceinv(eta)) <= 1e-10 * q.trace_Kninv     # ||inv L||_F^2 == tr of explicit inverse
    h = 1e-4
    lp1 = ProfileLikelihood.log_likelihood_and_gradient(z, X, Km, eta * (1 + h), with_rho=False)[0]
    lp0 = ProfileLikelihood.log_likelihood_and_gradient(z, X, Km, eta * (1 - h), with_rho=False)[0]
    _, deta, drho = ProfileLikelihood.log_likelihood_and_gradient(z, X, Km, eta)
    assert abs((lp1 - lp0) / (2 * h * eta) - deta) <= 1e-5 * abs(deta)
    # (d) d l^/d rho against a Richardson-extrapolated central difference of l^ over REGENERATED K(rho +- h):
    # truncation O(h^4), round-off ~ eps |l^| / h; tolerance 1e-6 relative
    del Km
    rho, hr = 0.1, 1e-3

    def lp_at(r):
        Kr = MixedCorrelation(gp.generate_correlation(pts, r, 2.5, device=True))
        return ProfileLikelihood.log_likelihood_and_gradient(z, X, Kr, eta, with_rho=False)[0]
    d1 = (lp_at(rho + hr) - lp_at(rho - hr)) / (2 * hr)
    d2 = (lp_at(rho + hr / 2) - lp_at(rho - hr / 2)) / hr
    fd = (4 * d2 - d1) / 3
    assert abs(fd - drho) <= 1e-6 * abs(drho), (fd, drho)


def _oracle_cell(pts, z, X, rho, nu, eta):
    """[l^, d l^/d eta, d l^/d rho] of one cell from the CPU oracle (reference formulas, one dposv per solve)."""
    from oracle import likelihood as L, matern
    Ko = L.MixedCorrelation(matern.generate_dense_correlation(pts, rho, nu), 'cholesky')
    dK = matern.matern_derivative_rho(pts, rho, nu)
    sig = L.ProfileLikelihood.find_optimal_sigma(z, X, Ko, eta)
    return [L.ProfileLikelihood.log_likelihood(z, X, Ko, False, [sig, eta]),
            L.ProfileLikelihood.log_likelihood_der1_eta(z, X, Ko, numpy.log10(eta)),
            L.ProfileLikelihood.log_likelihood_der1_rho(z, X, Ko, dK, eta)]


@pytest.mark.parametrize('rho,eta', [(0.1, 0.1), (0.1, 1.0), (0.2, 1e-2)])
def test_config2_n8k_cells_match_oracle(gp, rho, eta):
    """BASELINE configs[2] cell size (n = 8000, nu = 2.5): l^, d l^/d eta, d l^/d rho of three (rho, eta) cells against
    the CPU oracle, 1e-9 relative (eta >= 1e-2)."""
    from oracle import data_utilities as du
    from gaussian_proc._mixed_correlation import MixedCorrelation
    from gaussian_proc._likelihood import ProfileLikelihood
    n = 8000
    numpy.random.seed(0)
    pts = numpy.random.rand(n, 2)
    z, X = du.generate_data(pts, 0.2), du.generate_basis_functions(pts, 2)
    Km = MixedCorrelation(gp.generate_correlation(pts, rho, 2.5, device=True))
    got = ProfileLikelihood.log_likelihood_and_gradient(z, X, Km, eta)
    ref = _oracle_cell(pts, z, X, rho, 2.5, eta)
    for g_, r_ in zip(got, ref):
        assert abs(g_ - r_) <= RTOL * max(abs(r_), 1.0), (got, ref)


def test_n12k_cell_matches_oracle(gp):
    """One oracle comparison above n = 12 000 (same generator of synthetic inputs as configs[1]): 1e-9 relative."""
    from oracle import data_utilities as du
    from gaussian_proc._mixed_correlation import MixedCorrelation
    from gaussian_proc._likelihood import ProfileLikelihood
    n = 12500                                  # not a multiple of 128: the identity padding is exercised too
    numpy.random.seed(0)
    pts = numpy.random.rand(n, 2)
    z, X = du.generate_data(pts, 0.2), du.generate_basis_functions(pts, 2)
    Km = MixedCorrelation(gp.generate_correlation(pts, 0.1, 2.5, device=True))
    got = ProfileLikelihood.log_likelihood_and_gradient(z, X, Km, 0.1)
    ref = _oracle_cell(pts, z, X, 0.1, 2.5, 0.1)
    for g_, r_ in zip(got, ref):
        assert abs(g_ - r_) <= RTOL * max(abs(r_), 1.0), (got, ref)


@pytest.mark.parametrize('rows', [(0, 16), (16, 31), (31, 46), (46, 61)])
def test_golden_pickle_whole_grid_on_gpu(gp, golden_pickles, rows):
    """ALL 61 x 60 cells of the reference's shipped answer sheet data/OptimalCovariance_WithoutPrior.pickle
    (examples/FindOptimalCovarianceParameters.py:632-702: n = 900 grid points, general-nu Matern, profile likelihood at
    its eta root) through the GPU path, 1e-9 relative."""
    from oracle import data_utilities as du
    from gaussian_proc._mixed_correlation import MixedCorrelation
    from gaussian_proc._likelihood import ProfileLikelihood
    import contextlib
    import io
    pts = du.generate_points(30, 2, grid=True)
    z = du.generate_data(pts, 0.2)
    X = du.generate_basis_functions(pts, 2)
    worst = 0.0
    with contextlib.redirect_stdout(io.StringIO()):
        for i in range(rows[0], rows[1]):
            for j in range(60):
                rho, nu = golden_pickles['rho'][i], golden_pickles['nu'][j]
                Km = MixedCorrelation(gp.generate_correlation(pts, rho, nu, device=True))
                res = ProfileLikelihood.find_log_likelihood_der1_zeros(z, X, Km, [1e-3, 1e3])
                lp = ProfileLikelihood.log_likelihood(z, X, Km, False, [res['sigma'], res['eta']])
                ref = golden_pickles['Lp_noprior'][i, j]
                worst = max(worst, abs(lp - ref) / abs(ref))
                assert abs(lp - ref) <= 1e-9 * abs(ref), (i, j, lp, ref)
    assert worst <= 1e-9


# ------------------------------------------------------------ configs[0]: ~1k points, nu = 1.5, full MLE drivers
def test_config0_direct_and_profiled_training_match_oracle_drivers(gp):
    """BASELINE configs[0] (the reference's own CPU-runnable case): n = 1000 random 2-D points, nu = 1.5, rho = 0.1.
    The same optimiser / root-finder code drives the GPU evaluations and the oracle's; results must agree."""
    import scipy.optimize
    from functools import partial
    from oracle import data_utilities as du, likelihood as L, matern
    numpy.random.seed(0)
    pts = numpy.random.rand(1000, 2)
    z = du.generate_data(pts, 0.2)
    X = du.generate_basis_functions(pts, 2)
    K = gp.generate_correlation(pts, 0.1, 1.5)
    Ko = L.MixedCorrelation(matern.generate_dense_correlation(pts, 0.1, 1.5), 'cholesky')
    # profiled: root of d l/d eta on [1e-4, 1e3] (likelihood.py:86-94)
    res = gp.GaussianProcess(X, K, likelihood_method='profiled').train(z)
    ref = L.ProfileLikelihood.find_log_likelihood_der1_zeros(z, X, Ko, [1e-4, 1e3])
    assert rel([res['sigma'], res['sigma0'], res['eta']], [ref['sigma'], ref['sigma0'], ref['eta']]) <= 1e-7
    # direct: trust-exact from (0.2, 0.2) with the (variance-space) jacobian and hessian, tol 1e-3 (:378-384)
    res = gp.GaussianProcess(X, K, likelihood_method='direct').train(z)
    o = scipy.optimize.minimize(partial(L.DirectLikelihood.log_likelihood, z, X, Ko, True), [0.2, 0.2], method='trust-exact',
                                tol=1e-3, jac=partial(L.DirectLikelihood.log_likelihood_jacobian, z, X, Ko, True),
                                hess=partial(L.DirectLikelihood.log_likelihood_hessian, z, X, Ko, True))
    assert rel([res['sigma'], res['sigma0']], o.x) <= 1e-6 and abs(res['max_lp'] + o.fun) <= 1e-7 * abs(o.fun)


def test_grid_sweep_matches_oracle(gp):
    """configs[2] in miniature: (rho x eta) grid of profile likelihoods and both derivatives."""
    from gaussian_proc.sweep import likelihood_grid
    from oracle import data_utilities as du, likelihood as L, matern
    numpy.random.seed(4)
    pts = numpy.random.rand(400, 2)
    z = du.generate_data(pts, 0.2)
    X = du.generate_basis_functions(pts, 2)
    rhos, etas = numpy.linspace(0.05, 0.3, 3), numpy.logspace(-1, 1, 3)
    G = likelihood_grid(pts, z, X, 2.5, rhos, etas)
    assert G.shape == (3, 3, 3) and numpy.isfinite(G).all()
    for i, rho in enumerate(rhos):
        Ko = L.MixedCorrelation(matern.generate_dense_correlation(pts, rho, 2.5), 'cholesky')
        dK = matern.matern_derivative_rho(pts, rho, 2.5)
        for j, eta in enumerate(etas):
            sig = L.ProfileLikelihood.find_optimal_sigma(z, X, Ko, eta)
            ref = [L.ProfileLikelihood.log_likelihood(z, X, Ko, False, [sig, eta]),
                   L.ProfileLikelihood.log_likelihood_der1_eta(z, X, Ko, numpy.log10(eta)),
                   L.ProfileLikelihood.log_likelihood_der1_rho(z, X, Ko, dK, eta)]
            assert rel(G[i, j], ref) <= 1e-8, (rho, eta)


def test_block_cyclic_single_rank_matches_dense_engine(gp):
    """The distributed Cholesky + gradient on ONE rank (the multi-rank algorithm itself is covered on CPU/gloo in
    tests/test_blockcyclic_cpu.py and on 2-8 GPUs by bench.py --gpus N / tools/gpu_check_blockcyclic.py): GPU ops through
    the C ABI - look-ahead streams, replicated panels, rows of inv(L) with staircase k ranges, regenerated dK panels."""
    from oracle import data_utilities as du, likelihood as L, matern
    from gaussian_proc._blockcyclic import BlockCyclicCholesky
    numpy.random.seed(0)
    pts = numpy.random.rand(1000, 2)
    z = du.generate_data(pts, 0.2)
    X = du.generate_basis_functions(pts, 2)
    Kh = matern.generate_dense_correlation(pts, 0.1, 2.5)
    Ko = L.MixedCorrelation(Kh, 'cholesky')
    dK = matern.matern_derivative_rho(pts, 0.1, 2.5)
    s0 = L.ProfileLikelihood.find_optimal_sigma(z, X, Ko, 0.3)
    ref = [L.ProfileLikelihood.log_likelihood(z, X, Ko, False, [s0, 0.3]),
           L.ProfileLikelihood.log_likelihood_der1_eta(z, X, Ko, numpy.log10(0.3)),
           L.ProfileLikelihood.log_likelihood_der1_rho(z, X, Ko, dK, 0.3)]
    Kinv = numpy.linalg.inv(Kh + 0.3 * numpy.eye(1000))
    for nb in (128, 256):
        bc = BlockCyclicCholesky(pts, 0.1, 2.5, nb=nb)
        lp, sig = bc.profile_log_likelihood(z, X, 0.3)
        assert abs(lp - ref[0]) <= RTOL * abs(lp) and abs(sig - s0) <= RTOL * s0
        assert abs(bc.logdet() - Ko.logdet(0.3)) <= RTOL * abs(Ko.logdet(0.3))
        assert rel(bc.solve(numpy.c_[X, z]), Ko.solve(0.3, numpy.c_[X, z])) <= RTOL
        t1, t2 = bc.inverse_traces()
        assert abs(t1 - numpy.trace(Kinv)) <= RTOL * numpy.trace(Kinv)
        assert abs(t2 - numpy.sum(Kinv * dK)) <= RTOL * abs(numpy.sum(Kinv * dK))
        got = bc.profile_log_likelihood_and_gradient(z, X, 0.3)
        assert rel(got, ref) <= RTOL
    with pytest.raises(numpy.linalg.LinAlgError):
        bc.factor(-2.0)


def test_eigenvalue_method_matches_reference_vectors(gp, golden_likelihood):
    """imate_method='eigenvalue' (what the reference's Likelihood hard-codes, likelihood.py:41)."""
    from gaussian_proc._mixed_correlation import MixedCorrelation
    from gaussian_proc._likelihood import DirectLikelihood, Likelihood
    g = golden_likelihood
    nu, rho = g['cases'][1]
    pts, z, X = g['points'], g['z'], g['X']
    K = gp.generate_correlation(pts, rho, nu, device=True)
    Km = MixedCorrelation(K, imate_method='eigenvalue')
    for i, t in enumerate(g['log_etas']):
        eta = 10.0 ** t
        assert rel(Km.logdet(eta), g['c1_eigenvalue_logdet'][i]) <= RTOL
        assert rel(Km.traceinv(eta), g['c1_eigenvalue_traceinv'][i]) <= (1e-8 if t < 0 else RTOL)
        assert rel(Km.traceinv(eta, exponent=2), g['c1_eigenvalue_traceinv2'][i]) <= (1e-7 if t < 0 else RTOL)
    lam_ref = numpy.linalg.eigvalsh(K.to_numpy())
    assert rel(Km.trace(0.5, exponent=3), numpy.sum((lam_ref + 0.5) ** 3)) <= 1e-10
    # the spectrum itself (own tridiagonalisation + bisection) against LAPACK: absolute accuracy ~ eps ||K||
    lam = Km.K_eigenvalues.cpu().numpy()
    assert numpy.max(numpy.abs(numpy.sort(lam) - lam_ref)) <= 1e-12 * lam_ref[-1]
    h = list(g['hyper_direct'][0])
    lk = Likelihood(X, K, likelihood_method='direct', imate_method='eigenvalue')
    assert rel(lk.likelihood(z, h), g['c1_eigenvalue_direct_ll'][0]) <= RTOL
    assert rel(DirectLikelihood.log_likelihood_hessian(z, X, Km, False, h), g['c1_eigenvalue_direct_hess'][0]) <= 1e-8


def test_trace_interpolation_through_mixed_correlation(gp, problem):
    """MixedCorrelation(interpolate=True, interpolant_points=...) (mixed_correlation.py:52-66,167-170): traceinv is
    interpolated from Cholesky evaluations at the interpolant points; d l^/d eta with the interpolated trace stays within
    the interpolation error of the exact one, and its root is found by the unchanged driver."""
    from gaussian_proc._mixed_correlation import MixedCorrelation
    from gaussian_proc._likelihood import ProfileLikelihood
    pts, z, X = problem[:3]
    K = gp.generate_correlation(pts, 0.1, 0.5, device=True)
    P = [1e-2, 1e-1, 1.0, 10.0, 100.0]
    Ki = MixedCorrelation(K, interpolate=True, interpolant_points=P)
    Ke = MixedCorrelation(K)
    for eta in (0.03, 0.5, 4.0, 60.0):
        a, b = Ki.traceinv(eta), Ke.traceinv(eta)
        assert abs(a - b) <= 2e-2 * abs(b)
    for eta in P:
        assert abs(Ki.traceinv(eta) - Ke.traceinv(eta)) <= 1e-9 * abs(Ke.traceinv(eta))
    with pytest.raises(TypeError):
        MixedCorrelation(K, interpolate=True)
    n, m = X.shape
    d_i = ProfileLikelihood.log_likelihood_der1_eta(z, X, Ki, numpy.log10(0.5))
    d_e = ProfileLikelihood.log_likelihood_der1_eta(z, X, Ke, numpy.log10(0.5))
    assert abs(d_i - d_e) <= 0.5 * 2e-2 * abs(Ke.traceinv(0.5)) + 1e-9


def test_eigen_engine_matches_cholesky_engine(gp):
    """imate_method='eigenvalue': the fused evaluation on ONE eigendecomposition (every eta O(n^2 p)) returns the numbers
    of the Cholesky path: l^, d l^/d eta, d l^/d rho to 1e-9 (eta >= 1e-2), also through likelihood_grid."""
    from gaussian_proc._mixed_correlation import MixedCorrelation
    from gaussian_proc._likelihood import ProfileLikelihood, DirectLikelihood
    from gaussian_proc.sweep import likelihood_grid
    from oracle import data_utilities as du
    numpy.random.seed(3)
    pts = numpy.random.rand(1100, 2)
    z, X = du.generate_data(pts, 0.2), du.generate_basis_functions(pts, 2)
    K = gp.generate_correlation(pts, 0.12, 2.5, device=True)
    Ke, Kc = MixedCorrelation(K, imate_method='eigenvalue'), MixedCorrelation(K)
    for eta in (1e-2, 0.1, 1.0, 10.0, 300.0):
        a = ProfileLikelihood.log_likelihood_and_gradient(z, X, Ke, eta)
        b = ProfileLikelihood.log_likelihood_and_gradient(z, X, Kc, eta)
        tol = 1e-9 if eta >= 0.1 else 1e-8
        for x, y in zip(a, b):
            assert abs(x - y) <= tol * max(abs(y), 1.0), (eta, a, b)
    h = [0.3, 0.2]
    ja = DirectLikelihood.log_likelihood_jacobian(z, X, Ke, False, h)
    jb = DirectLikelihood.log_likelihood_jacobian(z, X, Kc, False, h)
    assert rel(ja, jb) <= 1e-9
    with pytest.raises(numpy.linalg.LinAlgError):
        ProfileLikelihood.log_likelihood_and_gradient(z, X, Ke, -2.0)
    rhos, etas = [0.1, 0.15], numpy.logspace(-1, 1, 5)
    Ge = likelihood_grid(pts, z, X, 2.5, rhos, etas, method='eigenvalue')
    Gc = likelihood_grid(pts, z, X, 2.5, rhos, etas)
    assert rel(Ge, Gc) <= 1e-9
    Gn = likelihood_grid(pts, z, X, 2.5, rhos, etas, method='eigenvalue', with_rho=False)      # spectrum only: no d/d rho
    assert rel(Gn[:, :, :2], Gc[:, :, :2]) <= 1e-9 and numpy.isnan(Gn[:, :, 2]).all()
    # odd size (leading dimension padding of the work copy) and tiny sizes
    for nn in (3, 4, 131):
        numpy.random.seed(nn)
        A = numpy.random.rand(nn, nn)
        A = A @ A.T + numpy.eye(nn)
        Kn = MixedCorrelation(A, imate_method='eigenvalue')
        ref = numpy.linalg.eigvalsh(A)
        assert numpy.max(numpy.abs(numpy.sort(Kn.K_eigenvalues.cpu().numpy()) - ref)) <= 1e-12 * ref[-1]
        assert rel(Kn.logdet(0.3), numpy.sum(numpy.log(ref + 0.3))) <= 1e-11


def test_in_place_edit_of_z_is_noticed(gp, problem):
    """The device copy of [X z] is cached by CONTENT (full digest): an in-place edit of a single entry of z between two
    evaluations must give the numbers of a fresh evaluation (the reference re-reads z on every call)."""
    from gaussian_proc._mixed_correlation import MixedCorrelation
    from gaussian_proc._likelihood import ProfileLikelihood
    pts, z, X = problem
    z = z.copy()
    K = gp.generate_correlation(pts, 0.1, 1.5, device=True)
    Km = MixedCorrelation(K)
    a = ProfileLikelihood.log_likelihood_and_gradient(z, X, Km, 0.5)
    z[137] += 0.5                                   # not one of any "sampled" positions
    b = ProfileLikelihood.log_likelihood_and_gradient(z, X, Km, 0.5)
    fresh = ProfileLikelihood.log_likelihood_and_gradient(z.copy(), X, MixedCorrelation(K), 0.5)
    assert b != a
    assert rel(b, fresh) <= 1e-13


def test_hessian_one_factorisation_and_sigma_space_optimiser(gp):
    """f-3: l + jacobian + hessian cost one fused evaluation each (no generic solves), der2_eta matches the explicit
    formula, and with chain_rule=True (derivatives in the optimiser's own variables) trust-exact ends with success: True
    at the oracle's maximum."""
    from oracle import data_utilities as du, likelihood as L, matern
    from gaussian_proc._mixed_correlation import MixedCorrelation
    from gaussian_proc._likelihood import DirectLikelihood, ProfileLikelihood
    numpy.random.seed(0)
    pts = numpy.random.rand(600, 2)
    z, X = du.generate_data(pts, 0.2), du.generate_basis_functions(pts, 2)
    n, m = X.shape
    Kd = gp.generate_correlation(pts, 0.1, 1.5, device=True)
    Km = MixedCorrelation(Kd)
    Kh = matern.generate_dense_correlation(pts, 0.1, 1.5)
    Ko = L.MixedCorrelation(Kh, 'cholesky')
    h = [0.25, 0.15]
    calls = []
    orig = Km.solve
    Km.solve = lambda *a, **k: (calls.append(1), orig(*a, **k))[1]
    H = DirectLikelihood.log_likelihood_hessian(z, X, Km, False, h)
    assert not calls                                          # nothing goes through the generic per-call solves
    assert rel(H, L.DirectLikelihood.log_likelihood_hessian(z, X, Ko, False, h)) <= 1e-8
    eta = 0.3
    Kinv = numpy.linalg.inv(Kh + eta * numpy.eye(n))
    Y = Kinv @ X
    M = Kinv - Y @ numpy.linalg.inv(X.T @ Y) @ Y.T
    Mz = M @ z
    ref = 0.5 * (n - m) / (z @ Mz) * ((numpy.sum(M * M) / (n - m) + (numpy.trace(M) / (n - m)) ** 2) * (z @ Mz) - 2 * Mz @ M @ Mz)
    assert rel(ProfileLikelihood.log_likelihood_der2_eta(z, X, Km, eta), ref) <= 1e-8
    res = DirectLikelihood.maximize_log_likelihood(z, X, Km, chain_rule=True)
    assert res['success']
    root = L.ProfileLikelihood.find_log_likelihood_der1_zeros(z, X, Ko, [1e-4, 1e3])
    # sigma enters only as sigma^2: the optimiser may land on either sign
    assert abs(res['eta'] - root['eta']) <= 2e-2 * root['eta'] and abs(abs(res['sigma']) - root['sigma']) <= 1e-2 * root['sigma']


@pytest.mark.parametrize('nu', [0.5, 2.5, 3.3])
def test_anisotropic_gradient_matches_oracle(gp, problem, nu):
    """One correlation scale per dimension (the reference kernel, _kernels.pyx:107-136): d l^/d scale[k] for every k from
    ONE factorisation against the oracle (M built as the reference's M_dot builds it, dK_k analytic), 1e-9; an isotropic
    scale still returns the scalar d/d rho, equal to the sum of the per-dimension derivatives."""
    from oracle import matern, likelihood as L
    from gaussian_proc._mixed_correlation import MixedCorrelation
    from gaussian_proc._likelihood import ProfileLikelihood, DirectLikelihood
    pts, z, X = problem
    scale = numpy.array([0.08, 0.13])
    Ko = L.MixedCorrelation(matern.generate_dense_correlation(pts, scale, nu), 'cholesky')
    Km = MixedCorrelation(gp.generate_correlation(pts, scale, nu, device=True))
    eta = 0.4
    lp, deta, drho = ProfileLikelihood.log_likelihood_and_gradient(z, X, Km, eta)
    assert numpy.shape(drho) == (2,)
    for k in range(2):
        ref = L.ProfileLikelihood.log_likelihood_der1_rho(z, X, Ko, matern.matern_derivative_scale(pts, scale, nu, k), eta)
        assert abs(drho[k] - ref) <= RTOL * abs(ref), (k, drho, ref)
    assert rel(deta, L.ProfileLikelihood.log_likelihood_der1_eta(z, X, Ko, numpy.log10(eta))) <= RTOL
    h = [0.3, 0.2]
    _, _, dr = DirectLikelihood.log_likelihood_and_gradient(z, X, Km, h)
    for k in range(2):
        ref = L.DirectLikelihood.log_likelihood_der1_rho(z, X, Ko, matern.matern_derivative_scale(pts, scale, nu, k), h)
        assert abs(dr[k] - ref) <= RTOL * abs(ref)
    Ki = MixedCorrelation(gp.generate_correlation(pts, 0.1, nu, device=True))
    Ka = MixedCorrelation(gp.generate_correlation(pts, numpy.array([0.1, 0.1 * (1 + 1e-15)]), nu, device=True))
    iso = ProfileLikelihood.log_likelihood_and_gradient(z, X, Ki, eta)[2]
    assert numpy.isscalar(iso) or numpy.ndim(iso) == 0


def test_legacy_rho_nu_surface_and_restartable_sweeps(gp, golden_pickles, tmp_path):
    """The reference's legacy workload as one call: profile_likelihood_surface over a corner of the (rho, nu) grid of
    data/OptimalCovariance_WithoutPrior.pickle (examples/FindOptimalCovarianceParameters.py:632-702) equals the shipped
    answer sheet (1e-9); a sweep restarted from its checkpoint directory recomputes only the missing rows; nu as a third
    axis of likelihood_grid equals the per-nu grids."""
    from oracle import data_utilities as du
    from gaussian_proc.sweep import likelihood_grid, profile_likelihood_surface
    pts = du.generate_points(30, 2, grid=True)
    z = du.generate_data(pts, 0.2)
    X = du.generate_basis_functions(pts, 2)
    rhos, nus = golden_pickles['rho'][[0, 30, 60]], golden_pickles['nu'][[0, 20, 59]]
    ck = str(tmp_path / 'surface')
    Lp, eta_hat = profile_likelihood_surface(pts, z, X, rhos, nus, checkpoint=ck)
    ref = golden_pickles['Lp_noprior'][numpy.ix_([0, 30, 60], [0, 20, 59])]
    assert rel(Lp, ref) <= RTOL and (eta_hat > 1e-3).all() and (eta_hat < 1e3).all()
    import os
    os.remove(os.path.join(ck, 'row_1.npy'))
    Lp2, _ = profile_likelihood_surface(pts, z, X, rhos, nus, checkpoint=ck)          # rows 0 and 2 come from the files
    assert numpy.array_equal(Lp2, Lp)
    etas = numpy.logspace(-1, 1, 3)
    G3 = likelihood_grid(pts, z, X, [1.5, 2.5], [0.1, 0.2], etas, checkpoint=str(tmp_path / 'grid'))
    assert G3.shape == (2, 2, 3, 3)
    assert rel(G3[1], likelihood_grid(pts, z, X, 2.5, [0.1, 0.2], etas)) <= 1e-12
    assert rel(G3[0], likelihood_grid(pts, z, X, 1.5, [0.1, 0.2], etas)) <= 1e-12
    assert numpy.array_equal(likelihood_grid(pts, z, X, [1.5, 2.5], [0.1, 0.2], etas, checkpoint=str(tmp_path / 'grid')), G3)
