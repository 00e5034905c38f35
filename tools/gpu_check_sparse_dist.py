"""Developer check (multi-GPU, torchrun): the stochastic estimators with probes split across ranks must reproduce the
single-rank estimate (same probe ids -> same samples; only the summation order differs)."""
import json, os, sys, time
import numpy
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'gaussian-process-param-estimation_b200'))
import torch
import torch.distributed as dist
world = int(os.environ.get('WORLD_SIZE', '1')); rank = int(os.environ.get('RANK', '0')); local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
from gaussian_proc._sparse import generate_sparse_correlation, SparseEngine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2 ** 18
numpy.random.seed(0)
pts = numpy.random.rand(n, 2)
rho = 0.005 * numpy.sqrt(2 ** 20 / n)
K = generate_sparse_correlation(pts, numpy.array([rho, rho]), 0.5, 1e-3 * 2 ** 20 / n, device=True, with_derivative=True)
opts = {'seed': 0, 'lanczos_degree': 30, 'min_num_samples': 32, 'max_num_samples': 32, 'batch': 8}
out = {'world': world, 'n': n, 'nnz': K.nnz}
for name, pr in (('split', (rank, world)), ('single', None)):
    eng = SparseEngine(K, 'slq', opts, probe_range=pr)
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    t0 = time.perf_counter()
    ld = eng.logdet(10.0)
    tr = eng.traceinv_dK(10.0)
    torch.cuda.synchronize()
    out[name] = {'logdet': ld, 'tr_dK': tr, 't_s': time.perf_counter() - t0, 'samples': eng.last_info['num_samples']}
out['rel_diff'] = [abs(out['split'][k] - out['single'][k]) / abs(out['single'][k]) for k in ('logdet', 'tr_dK')]
out['speedup'] = out['single']['t_s'] / out['split']['t_s']
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
