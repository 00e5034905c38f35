"""Small end-to-end case for compute-sanitizer (racecheck / memcheck): the dense fused evaluation (look-ahead potrf, recursive
trtri on side streams, TMA GEMM), the eigenvalue path, the distributed path on one rank and the sparse evaluation (Krylov
side-stream overlap). python tools/sanitize_case.py [dense|eigen|dist|sparse ...]"""
import os, sys
import numpy
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'gaussian-process-param-estimation_b200'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch
import bench, gaussian_proc
from gaussian_proc._mixed_correlation import MixedCorrelation
from gaussian_proc._likelihood import ProfileLikelihood, DirectLikelihood
what = sys.argv[1:] or ['dense', 'eigen', 'dist', 'sparse']
if 'dense' in what:
    pts, z, X = bench.make_inputs(1100)
    Km = MixedCorrelation(gaussian_proc.generate_correlation(pts, 0.1, 2.5, device=True))
    print('dense', ProfileLikelihood.log_likelihood_and_gradient(z, X, Km, 0.1))
    print('hess', DirectLikelihood.log_likelihood_hessian(z, X, Km, False, [0.3, 0.2]).ravel())
if 'eigen' in what:
    pts, z, X = bench.make_inputs(500)
    Km = MixedCorrelation(gaussian_proc.generate_correlation(pts, 0.1, 2.5, device=True), imate_method='eigenvalue')
    print('eigen', ProfileLikelihood.log_likelihood_and_gradient(z, X, Km, 0.1, with_rho=False))
if 'dist' in what:
    from gaussian_proc._blockcyclic import BlockCyclicCholesky
    pts, z, X = bench.make_inputs(700)
    print('dist', BlockCyclicCholesky(pts, 0.1, 2.5, nb=128).profile_log_likelihood_and_gradient(z, X, 0.3))
if 'sparse' in what:
    from gaussian_proc._sparse import generate_sparse_correlation
    pts, z, X = bench.make_inputs(3000)
    K = generate_sparse_correlation(pts, numpy.array([0.03, 0.03]), 0.5, 0.01, device=True, with_derivative=True)
    Km = MixedCorrelation(K, imate_method='slq', imate_options={'seed': 0})
    print('sparse', ProfileLikelihood.log_likelihood_and_gradient(z, X, Km, 2.0))
torch.cuda.synchronize()
print('done')
