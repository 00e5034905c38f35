"""CPU tests of the multi-GPU plumbing with the gloo backend, world_size = 2: cell partition + final gather of the
grid sweep, the (count, sum, sum-of-squares) all-reduce of the stochastic estimators, and probe sharding."""

import os
import socket

import numpy
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _cell(rho, eta):
    return (numpy.sin(3 * rho) + numpy.log(eta), rho * eta, rho - eta)


def _fake_samples(first, width):
    ids = numpy.arange(first, first + width, dtype=float)
    return numpy.stack([numpy.cos(ids) + 5.0, 0.1 * numpy.sin(ids) + 2.0], axis=1)


def _worker(rank, world, port, out):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, 'gaussian-process-param-estimation_b200'))
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from gaussian_proc import _distributed as gpd
    from gaussian_proc.sweep import likelihood_grid
    from gaussian_proc._sparse import SparseEngine, DEFAULTS
    res = {}
    rhos, etas = numpy.linspace(0.05, 0.3, 5), numpy.logspace(-2, 2, 4)
    res['grid'] = likelihood_grid(None, None, None, 2.5, rhos, etas, evaluate=_cell)
    res['sum'] = gpd.allreduce_sum(numpy.array([rank + 1.0, 10.0]))
    res['rows'] = gpd.allgather_rows(numpy.full((rank + 1, 2), float(rank)))
    eng = object.__new__(SparseEngine)
    eng.opt = dict(DEFAULTS, min_num_samples=10, max_num_samples=24, batch=4, error_rtol=1e-9)
    eng.probe_range = (rank, world)
    res['est'] = eng._run_estimator(_fake_samples, 2)[:3]

    # a breakdown on ONE rank (indefinite K + eta I) must raise on EVERY rank after the collective, not deadlock
    def breaking(first, width):
        if rank == 1:
            raise numpy.linalg.LinAlgError('non-positive Ritz value')
        return _fake_samples(first, width)
    try:
        eng._run_estimator(breaking, 2)
        res['breakdown'] = 'no error'
    except numpy.linalg.LinAlgError as exc:
        res['breakdown'] = 'raised: ' + str(exc)[:40]
    dist.barrier()
    dist.destroy_process_group()
    out[rank] = res


def test_two_rank_gloo_matches_single_process():
    world = 2
    from conftest import RankResults
    out = RankResults()          # no multiprocessing.Manager: it would fork() this process
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, 'gaussian-process-param-estimation_b200'))
    from gaussian_proc.sweep import likelihood_grid
    from gaussian_proc._sparse import SparseEngine, DEFAULTS
    rhos, etas = numpy.linspace(0.05, 0.3, 5), numpy.logspace(-2, 2, 4)
    single = likelihood_grid(None, None, None, 2.5, rhos, etas, evaluate=_cell)
    for r in range(world):
        assert numpy.array_equal(out[r]['grid'], single)              # identical on every rank, no NaN holes
        assert numpy.array_equal(out[r]['sum'], numpy.array([3.0, 20.0]))
        assert numpy.array_equal(out[r]['rows'], numpy.array([[0.0, 0.0], [1.0, 1.0], [1.0, 1.0]]))
    eng = object.__new__(SparseEngine)
    eng.opt = dict(DEFAULTS, min_num_samples=10, max_num_samples=24, batch=4, error_rtol=1e-9)
    eng.probe_range = None
    mean1, half1, n1, state = eng._run_estimator(_fake_samples, 2)
    # continuing a finished run changes nothing (max_num_samples reached)
    mean2, half2, n2, _ = eng._run_estimator(_fake_samples, 2, state=state)
    assert n2 == n1 and numpy.array_equal(mean2, mean1)
    for r in range(world):
        assert out[r]['breakdown'].startswith('raised'), out[r]['breakdown']
        mean, half, n = out[r]['est']
        assert n == n1 == 24
        assert numpy.allclose(mean, mean1, rtol=0, atol=1e-12) and numpy.allclose(half, half1, rtol=0, atol=1e-12)


def test_partition_covers_all_cells():
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, 'gaussian-process-param-estimation_b200'))
    from gaussian_proc._distributed import partition_cells
    for n in (1, 5, 64, 61):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                b, e = partition_cells(n, world, r)
                seen += list(range(b, e))
            assert seen == list(range(n))
