"""Developer check (multi-GPU, run under torchrun): distributed Cholesky + gradient vs the single-GPU dense engine, then a
timed large-n evaluation.  torchrun --nproc-per-node N tools/gpu_check_blockcyclic.py [n_big] [nb] [n_check ...]"""
import json, os, sys, time
import numpy
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'gaussian-process-param-estimation_b200'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch
import torch.distributed as dist

world = int(os.environ.get('WORLD_SIZE', '1'))
rank = int(os.environ.get('RANK', '0'))
local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
import bench
from gaussian_proc._blockcyclic import BlockCyclicCholesky
from gaussian_proc import generate_correlation
from gaussian_proc._mixed_correlation import MixedCorrelation
from gaussian_proc._likelihood import ProfileLikelihood

nbig = int(sys.argv[1]) if len(sys.argv) > 1 else 40960
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 512
checks = [int(a) for a in sys.argv[3:]] or [6000]


def say(obj):
    if rank == 0:
        print(json.dumps(obj), flush=True)


# ---- correctness against the single-GPU path ----------------------------------------------------------------------
for n in checks:
    pts, z, X = bench.make_inputs(n)
    out = {'world': world, 'check_n': n, 'nb': nb}
    try:
        bc = BlockCyclicCholesky(pts, 0.1, 2.5, nb=nb)
        grad = bc.profile_log_likelihood_and_gradient(z, X, 0.1)
        ld = bc.logdet()
        sol = bc.solve(z)
        Km = MixedCorrelation(generate_correlation(pts, 0.1, 2.5, device=True))
        ref_grad = ProfileLikelihood.log_likelihood_and_gradient(z, X, Km, 0.1)
        ref_ld = Km.logdet(0.1)
        ref_sol = Km.solve(0.1, z)
        out.update({'grad_rel': [abs(a - b) / abs(b) for a, b in zip(grad, ref_grad)], 'logdet_rel': abs(ld - ref_ld) / abs(ref_ld),
                    'solve_rel': float(numpy.max(numpy.abs(sol - ref_sol)) / numpy.max(numpy.abs(ref_sol))), 'stats': dict(bc.stats)})
        del Km, bc
    except Exception as exc:  # noqa: BLE001
        out['error'] = repr(exc)[:300]
    torch.cuda.empty_cache()
    say(out)
# ---- timing ------------------------------------------------------------------------------------------------------
if nbig > 0:
    pts, z, X = bench.make_inputs(nbig)
    bc = BlockCyclicCholesky(pts, 0.1, 2.5, nb=nb)
    for rep in range(2):
        out = {'world': world, 'n': nbig, 'nb': nb, 'rep': rep}
        try:
            torch.cuda.synchronize()
            if world > 1: dist.barrier()
            t0 = time.perf_counter()
            g = bc.profile_log_likelihood_and_gradient(z, X, 0.1)
            torch.cuda.synchronize()
            if world > 1: dist.barrier()
            t1 = time.perf_counter()
            st = dict(bc.stats)
            out.update({'t_total_s': t1 - t0, 'loglik_grad': [float(v) for v in g], 'stats': st,
                        'potrf_tflops_total': (nbig ** 3 / 3.0) / st['factor_s'] * 1e-12,
                        'total_tflops_per_gpu': float(nbig) ** 3 / (t1 - t0) / world * 1e-12})
        except Exception as exc:  # noqa: BLE001
            out['error'] = repr(exc)[:300]
        say(out)
if world > 1:
    dist.destroy_process_group()
