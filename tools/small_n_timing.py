"""Developer timing: loglik+grad evaluation latency at small n (configs[0] scale) on the GPU vs the CPU oracle."""
import os, sys, time
import numpy
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'gaussian-process-param-estimation_b200'))
import torch
from bench import make_inputs
import gaussian_proc
from gaussian_proc._mixed_correlation import MixedCorrelation
from gaussian_proc._likelihood import ProfileLikelihood
from oracle import likelihood as L, matern
for n in (500, 1000, 2000, 4000):
    pts, z, X = make_inputs(n)
    K = gaussian_proc.generate_correlation(pts, 0.1, 1.5, device=True)
    Km = MixedCorrelation(K)
    ProfileLikelihood.log_likelihood_and_gradient(z, X, Km, 0.1)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(10):
        ProfileLikelihood.log_likelihood_and_gradient(z, X, Km, 0.1 + 0.01 * i)
    torch.cuda.synchronize(); tg = (time.perf_counter() - t0) / 10
    t0 = time.perf_counter()
    for i in range(5):
        ProfileLikelihood.log_likelihood_der1_eta(z, X, Km, numpy.log10(0.1 + 0.01 * i))
    torch.cuda.synchronize(); td = (time.perf_counter() - t0) / 5
    Kh = matern.generate_dense_correlation(pts, numpy.array([0.1, 0.1]), 1.5)
    Ko = L.MixedCorrelation(Kh, 'cholesky')
    t0 = time.perf_counter()
    L.ProfileLikelihood.log_likelihood_der1_eta(z, X, Ko, numpy.log10(0.1))
    tc = time.perf_counter() - t0
    print('n=%d  GPU loglik+grad(eta,rho) %.2f ms   GPU dl/deta %.2f ms   CPU oracle dl/deta %.2f ms' % (n, tg * 1e3, td * 1e3, tc * 1e3), flush=True)
