"""Developer profile target: ONE profile log-likelihood + gradient evaluation at a new rho through the public API on the
sparse n = 2^20 workload (use under ncu --metrics gpu__time_duration.sum for the launch list)."""
import os, sys, time
import numpy
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'gaussian-process-param-estimation_b200'))
import torch
from bench import make_inputs
from gaussian_proc._sparse import generate_sparse_correlation
from gaussian_proc._mixed_correlation import MixedCorrelation
from gaussian_proc._likelihood import ProfileLikelihood

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2 ** 20
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
pts, z, X = make_inputs(n)
scale = numpy.array([0.005, 0.005])
for r in range(reps):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    K = generate_sparse_correlation(pts, scale, 0.5, 1e-3, device=True, with_derivative=True)
    Km = MixedCorrelation(K, imate_method='slq', imate_options={'seed': 0, 'lanczos_degree': 30})
    out = ProfileLikelihood.log_likelihood_and_gradient(z, X, Km, 10.0)
    torch.cuda.synchronize()
    print('eval %d: %.1f ms' % (r, (time.perf_counter() - t0) * 1e3), out)
    del K, Km
