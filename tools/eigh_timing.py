"""Developer timing: library symmetric eigensolver (cuSOLVER behind torch.linalg.eigh) against one Cholesky-based cell."""
import os, sys, time, numpy
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'gaussian-process-param-estimation_b200'))
import torch
import bench, gaussian_proc
for n in (4000, 8000):
    pts, z, X = bench.make_inputs(n)
    K = gaussian_proc.generate_correlation(pts, 0.1, 2.5, device=True)
    A = K.data[:n, :n].contiguous()
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        lam, V = torch.linalg.eigh(A)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        lam2 = torch.linalg.eigvalsh(A)
        torch.cuda.synchronize(); t2 = time.perf_counter()
    print('n=%d eigh %.3f s  eigvalsh %.3f s' % (n, t1 - t0, t2 - t1), flush=True)
