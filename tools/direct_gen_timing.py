"""Developer timing at n = 2^20: CSR generation + row-blocked build against the direct row-block generation."""
import json, os, sys, time
import numpy
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'gaussian-process-param-estimation_b200'))
import torch
from gaussian_proc._sparse import generate_sparse_correlation, generate_sparse_operator, SparseEngine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2 ** 20
numpy.random.seed(0)
pts = numpy.random.rand(n, 2)
pts.setflags(write=False)
rho = 0.005 * numpy.sqrt(2 ** 20 / n)
dens = min(0.05, 1e-3 * 2 ** 20 / n)
sc = numpy.array([rho, rho])
def t(fn, reps=4):
    best = 1e30
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0); del r
    return best * 1e3
out = {'n': n}
for dk in (False, True):
    tag = '_dk' if dk else ''
    out['csr_generate_ms' + tag] = t(lambda: generate_sparse_correlation(pts, sc, 0.5, dens, device=True, with_derivative=dk))
    out['csr_generate_plus_build_ms' + tag] = t(lambda: SparseEngine(generate_sparse_correlation(pts, sc, 0.5, dens, device=True, with_derivative=dk), 'slq', {}))
    out['direct_ms' + tag] = t(lambda: SparseEngine(generate_sparse_operator(pts, sc, 0.5, dens, with_derivative=dk), 'slq', {}))
K = generate_sparse_operator(pts, sc, 0.5, dens, with_derivative=True)
out['nnz'] = K.nnz; out['block_columns'] = K.bidx.numel(); out['fill_ratio'] = K.bidx.numel() * 16 / K.nnz
e1 = SparseEngine(generate_sparse_correlation(pts, sc, 0.5, dens, device=True, with_derivative=True), 'slq', {})
e2 = SparseEngine(K, 'slq', {})
V = e1.probes(0, 8)
out['spmm_max_abs_diff'] = float((e1.spmm(1.0, V) - e2.spmm(1.0, V)).abs().max().item())
out['dspmm_max_abs_diff'] = float((e1.spmm(0.0, V, derivative=True) - e2.spmm(0.0, V, derivative=True)).abs().max().item())
print(json.dumps(out))
