"""Developer profile target: a few grid-sweep cells at n = 8000 (use under ncu --metrics gpu__time_duration.sum)."""
import os, sys, numpy
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'gaussian-process-param-estimation_b200'))
import torch
import bench
from gaussian_proc.sweep import likelihood_grid
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8000
pts, z, X = bench.make_inputs(n)
G = likelihood_grid(pts, z, X, 2.5, [0.1], numpy.logspace(-2, 2, 4), concurrency=1)
torch.cuda.synchronize()
print(G[0, :, 0])
