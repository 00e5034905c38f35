"""ncu target: the direct row-block generator (count + fill) and one peer-gather SpMM (row-slab engine with a group of one)
at n = 2^20. profiles/r02_sparse_frontend_ncu_summary.md was produced from
  ncu --metrics <list> -k regex:"sparse_blocks|bcsr8_spmm" python tools/sparse_frontend_profile.py"""
import os, sys
import numpy
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'gaussian-process-param-estimation_b200'))
import torch
from gaussian_proc._sparse import generate_sparse_operator, SparseEngine
from gaussian_proc._slab import SlabSparseEngine
n = 2 ** 20
numpy.random.seed(0)
pts = numpy.random.rand(n, 2)
pts.setflags(write=False)
sc = numpy.array([0.005, 0.005])
K = generate_sparse_operator(pts, sc, 0.5, 1e-3, with_derivative=True)
eng = SparseEngine(K, 'slq', {})
V = eng.probes(0, 16)
eng.spmm(10.0, V)
Ks = generate_sparse_operator(pts, sc, 0.5, 1e-3, with_derivative=True, row_slab=(0, 1))
slab = SlabSparseEngine(Ks, 'slq', {}, rank=0, world=1)
slab.spmm(10.0, V)
torch.cuda.synchronize()
print('done')
