"""The reference's headline driver on the sparse n = 2^20 workload: maximum profile likelihood by the root of d l^/d eta
(ProfileLikelihood.find_log_likelihood_der1_zeros, reference _profile_likelihood.py:244-415; the 'presented method' of
examples/CompareVariousNumberOfPoints.py whose legacy CPU times are in BASELINE.md: 7 940 s pre-computation + 2 093 s
root finding at n = 1 048 576)."""
import json, os, sys, time
import numpy
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'gaussian-process-param-estimation_b200'))
import torch
from bench import make_inputs
from gaussian_proc import _device as dev
from gaussian_proc._sparse import generate_sparse_correlation
from gaussian_proc._mixed_correlation import MixedCorrelation
from gaussian_proc._likelihood import ProfileLikelihood

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2 ** 20
pts, z, X = make_inputs(n)
scale = numpy.array([0.005, 0.005]) * numpy.sqrt(2 ** 20 / float(n))
dens = 1e-3 * 2 ** 20 / float(n)
out = {'n': n}
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    K = generate_sparse_correlation(pts, scale, 0.5, dens, device=True)
    Km = MixedCorrelation(K, imate_method='slq', imate_options={'seed': 0, 'lanczos_degree': 30})
    torch.cuda.synchronize(); t1 = time.perf_counter()
    l0 = dev.lib.gp_launch_count()
    res = ProfileLikelihood.find_log_likelihood_der1_zeros(z, X, Km, [10.0, 1e3])
    torch.cuda.synchronize(); t2 = time.perf_counter()
    out['run%d' % rep] = {'generate_and_build_s': t1 - t0, 'root_find_s': t2 - t1, 'launches': int(dev.lib.gp_launch_count() - l0),
                          'result': {k: (float(v) if not isinstance(v, bool) else v) for k, v in res.items()}, 'nnz': K.nnz}
    del K, Km
print(json.dumps(out))
