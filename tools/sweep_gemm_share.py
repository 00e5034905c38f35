"""Developer timing: during the n = 8000 grid sweep, what share of the wall time has a DMMA GEMM in flight, and at what
executed-tile rate (gp_gemm_profile_*)."""
import ctypes, json, os, sys, time
import numpy
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'gaussian-process-param-estimation_b200'))
import torch
import bench
from gaussian_proc import _device as dev
from gaussian_proc.sweep import likelihood_grid
lib = dev.lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8000
pts, z, X = bench.make_inputs(n)
etas = numpy.logspace(-2, 2, 16)
rhos = numpy.linspace(0.1, 0.2, 2)
for conc in (1, 4):
    likelihood_grid(pts, z, X, 2.5, rhos[:1], etas[:conc + 1], concurrency=conc)
    torch.cuda.synchronize()
    lib.gp_gemm_profile_enable(1)
    t0 = time.perf_counter()
    likelihood_grid(pts, z, X, 2.5, rhos, etas, concurrency=conc)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    gm, gu, gf, gl = ctypes.c_double(), ctypes.c_double(), ctypes.c_double(), ctypes.c_longlong()
    lib.gp_gemm_profile_read(ctypes.byref(gm), ctypes.byref(gu), ctypes.byref(gf), ctypes.byref(gl))
    lib.gp_gemm_profile_enable(0)
    cells = 32
    print(json.dumps({'n': n, 'concurrency': conc, 'wall_ms_per_cell': dt * 1e3 / cells, 'gemm_union_ms_per_cell': gu.value / cells,
                      'gemm_sum_ms_per_cell': gm.value / cells, 'gemm_share_of_wall': gu.value / (dt * 1e3),
                      'executed_tile_TF_over_union': gf.value / (gu.value * 1e-3) * 1e-12,
                      'executed_over_algorithmic': gf.value / (cells * float(n) ** 3), 'launches_per_cell': gl.value / cells,
                      'cell_TF': cells * float(n) ** 3 / dt * 1e-12}))
