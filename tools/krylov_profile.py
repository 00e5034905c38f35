"""Developer profile target: one short Lanczos run and one CG solve on the row-blocked operator at n = 2^20, B = 16
(use under ncu --metrics gpu__time_duration.sum for the per-kernel launch list)."""
import ctypes, os, sys
import numpy
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'gaussian-process-param-estimation_b200'))
import torch
from gaussian_proc import _device as dev
from gaussian_proc._sparse import generate_sparse_correlation, SparseEngine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2 ** 20
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
numpy.random.seed(0)
pts = numpy.random.rand(n, 2)
K = generate_sparse_correlation(pts, numpy.array([0.005, 0.005]), 0.5, 1e-3, device=True, with_derivative=True)
e = SparseEngine(K, 'slq', {'lanczos_degree': 6, 'cg_maxiter': 8})
print('samples', e._slq_samples(10.0, 0, B)[0])
try:
    e.solve_dev(10.0, e.probes(0, B))
except Exception as ex:  # noqa: BLE001  (maxiter is reached on purpose)
    print('cg:', str(ex)[:60])
torch.cuda.synchronize()
