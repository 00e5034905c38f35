"""BASELINE configs[2] at scale (torchrun, one rank per GPU): the (rho x eta) grid of profile log-likelihood + gradient
cells at n = 8000, contiguous rho groups per rank, no data-path collective, results all-gathered.
  torchrun --nproc-per-node N tools/gpu_sweep_scale.py [n_rho] [n_eta] [n]"""
import json, os, sys, time
import numpy
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'gaussian-process-param-estimation_b200'))
import torch
import torch.distributed as dist
world = int(os.environ.get('WORLD_SIZE', '1')); rank = int(os.environ.get('RANK', '0')); local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
from bench import make_inputs
from gaussian_proc.sweep import likelihood_grid

n_rho = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n_eta = int(sys.argv[2]) if len(sys.argv) > 2 else 64
n = int(sys.argv[3]) if len(sys.argv) > 3 else 8000
sparse = len(sys.argv) > 4 and sys.argv[4] == 'sparse'
eigen = len(sys.argv) > 4 and sys.argv[4] == 'eigen'
pts, z, X = make_inputs(n)
kw = {}
if sparse:      # configs[3] family: nu = 0.5, rho around 0.005, density 1e-3, eta >= 10 (hard-thresholded K is indefinite)
    rhos = numpy.linspace(0.004, 0.006, n_rho) * numpy.sqrt(2 ** 20 / float(n))
    etas = numpy.logspace(1, 3, n_eta)
    nu = 0.5
    kw = dict(sparse=True, density=1e-3 * 2 ** 20 / float(n), imate_options={'seed': 0, 'lanczos_degree': 30})
else:
    rhos = numpy.linspace(0.05, 0.3, n_rho)
    etas = numpy.logspace(-2, 2, n_eta)
    nu = 2.5
    if eigen:
        kw = dict(method='eigenvalue')
likelihood_grid(pts, z, X, nu, rhos[:world], etas[:min(4, n_eta)], **kw)            # warm-up: one small row per rank
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
G = likelihood_grid(pts, z, X, nu, rhos, etas, **kw)
torch.cuda.synchronize()
dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device='cuda')
if world > 1:
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
dt = float(dt.item())
if rank == 0:
    i, j = numpy.unravel_index(numpy.nanargmax(G[:, :, 0]), G[:, :, 0].shape)
    print(json.dumps({'world': world, 'n': n, 'sparse': sparse, 'method': 'eigenvalue' if eigen else ('slq' if sparse else 'cholesky'), 'cells': n_rho * n_eta, 'seconds': dt, 'cells_per_s': n_rho * n_eta / dt,
                      'tflops_total': None if sparse else n_rho * n_eta * float(n) ** 3 / dt * 1e-12, 'finite': bool(numpy.isfinite(G).all()),
                      'argmax': {'rho': float(rhos[i]), 'eta': float(etas[j]), 'lp': float(G[i, j, 0]),
                                 'dlp_deta': float(G[i, j, 1]), 'dlp_drho': float(G[i, j, 2])}}))
if world > 1:
    dist.destroy_process_group()
