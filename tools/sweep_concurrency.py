"""Developer timing: cells/s of the grid sweep against the number of cells in flight per GPU."""
import sys, time, numpy, torch, os
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'gaussian-process-param-estimation_b200'))
import bench
from gaussian_proc.sweep import likelihood_grid
for n, nrho, neta in ((8000, 3, 32), (20000, 2, 12)):
    pts, z, X = bench.make_inputs(n)
    etas = numpy.logspace(-2, 2, neta)
    rhos = numpy.linspace(0.1, 0.2, nrho)
    ref = None
    for conc in (1, 2, 3, 4, 6):
        likelihood_grid(pts, z, X, 2.5, rhos[:1], etas[:conc + 1], concurrency=conc)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        G = likelihood_grid(pts, z, X, 2.5, rhos, etas, concurrency=conc)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        if ref is None: ref = G
        cells = nrho * neta
        print(n, 'concurrency', conc, 'cells/s %.2f' % (cells / dt), 'TF %.2f' % (cells * n ** 3 / dt * 1e-12), 'maxdiff', float(numpy.max(numpy.abs(G - ref) / numpy.abs(ref))), flush=True)
