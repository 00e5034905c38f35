import sys, time, numpy, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/gaussian-process-param-estimation_b200')
import bench
from gaussian_proc.sweep import likelihood_grid
for n in (8000, 20000):
    pts, z, X = bench.make_inputs(n)
    etas = numpy.logspace(-2, 2, 8)
    ref = None
    for conc in (1, 2, 3):
        likelihood_grid(pts, z, X, 2.5, [0.1], etas[:2], concurrency=conc)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        G = likelihood_grid(pts, z, X, 2.5, [0.1, 0.2], etas, concurrency=conc)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        if ref is None: ref = G
        print(n, 'concurrency', conc, 'cells/s %.2f' % (16 / dt), 'TF %.2f' % (16 * n ** 3 / dt * 1e-12), 'maxdiff', float(numpy.max(numpy.abs(G - ref) / numpy.abs(ref))))
