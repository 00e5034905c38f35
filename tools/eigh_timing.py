"""Developer timing: the library's own symmetric eigenvalue path (gp_sytrd_f64 + gp_stebz_f64, csrc/gp_eig.cu) against the
cuSOLVER routines behind torch.linalg.eigvalsh / eigh (comparison only - the package does not call them) and against one
Cholesky-based cell."""
import json, os, sys, time, numpy
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'gaussian-process-param-estimation_b200'))
import torch
import bench, gaussian_proc
from gaussian_proc._dense import EigenEngine
from gaussian_proc._mixed_correlation import MixedCorrelation
from gaussian_proc._likelihood import ProfileLikelihood
for n in [int(a) for a in sys.argv[1:]] or [4000, 8000]:
    pts, z, X = bench.make_inputs(n)
    K = gaussian_proc.generate_correlation(pts, 0.1, 2.5, device=True)
    A = K.data[:n, :n].contiguous()
    out = {'n': n}
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        eng = EigenEngine(K)
        torch.cuda.synchronize(); out['own_sytrd_stebz_s'] = time.perf_counter() - t0
    torch.cuda.synchronize(); t0 = time.perf_counter()
    lam2 = torch.linalg.eigvalsh(A)
    torch.cuda.synchronize(); out['cusolver_eigvalsh_s'] = time.perf_counter() - t0
    torch.cuda.synchronize(); t0 = time.perf_counter()
    lam, V = torch.linalg.eigh(A)
    torch.cuda.synchronize(); out['cusolver_eigh_s'] = time.perf_counter() - t0
    out['max_abs_diff_vs_cusolver'] = float((eng.lam.sort()[0] - lam2).abs().max())
    out['bytes_algorithmic'] = 16.0 * n ** 3 / 3.0
    out['sytrd_GBs'] = out['bytes_algorithmic'] / out['own_sytrd_stebz_s'] * 1e-9
    Km = MixedCorrelation(K, imate_method='eigenvalue')
    ProfileLikelihood.log_likelihood_and_gradient(z, X, Km, 0.1, with_rho=False)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for eta in numpy.logspace(-2, 2, 32):
        ProfileLikelihood.log_likelihood_and_gradient(z, X, Km, eta, with_rho=False)
    torch.cuda.synchronize(); out['per_eta_ms_after_reduction'] = (time.perf_counter() - t0) / 32 * 1e3
    Kc = MixedCorrelation(K)
    ProfileLikelihood.log_likelihood_and_gradient(z, X, Kc, 0.1, with_rho=False)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ProfileLikelihood.log_likelihood_and_gradient(z, X, Kc, 0.2, with_rho=False)
    torch.cuda.synchronize(); out['cholesky_cell_ms'] = (time.perf_counter() - t0) * 1e3
    print(json.dumps(out), flush=True)
