"""CPU (gloo, world_size 2, 3 and 4) tests of the block-cyclic distributed Cholesky + gradient: the distributed algorithm
(snake ownership, look-ahead order, replicated panels, rows of inv(L), staircase k ranges, trace reductions) is run with
the NumPy `ops` of tests/numpy_ops.py and checked against the oracle's dense values."""

import os
import socket
import sys

import numpy
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _problem(n):
    from oracle import data_utilities as du
    numpy.random.seed(0)
    pts = numpy.random.rand(n, 2)
    return pts, du.generate_data(pts, 0.2), du.generate_basis_functions(pts, 2)


def _worker(rank, world, port, n, nb, out):
    for p in (ROOT, os.path.join(ROOT, 'gaussian-process-param-estimation_b200'), os.path.join(ROOT, 'tests')):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from gaussian_proc._blockcyclic import BlockCyclicCholesky
    from numpy_ops import NumpyOps
    pts, z, X = _problem(n)
    bc = BlockCyclicCholesky(pts, 0.1, 2.5, nb=nb, ops=NumpyOps())
    bc.factor(0.3)
    res = {'grid': (bc.P_r, bc.P_c), 'logdet': bc.logdet(), 'solve': bc.solve(numpy.c_[X, z]), 'vec': bc.solve(z),
           'traces': bc.inverse_traces(), 'recv': bc.bytes_received, 'rows': list(bc.J_loc)}
    res['lp'] = bc.profile_log_likelihood(z, X, 0.3)
    res['grad'] = bc.profile_log_likelihood_and_gradient(z, X, 0.3)
    try:
        BlockCyclicCholesky(pts, 0.1, 2.5, nb=nb, ops=NumpyOps()).factor(-2.0)   # not positive definite
        res['raised'] = False
    except numpy.linalg.LinAlgError:
        res['raised'] = True
    dist.barrier()
    dist.destroy_process_group()
    out[rank] = res


@pytest.mark.parametrize('world,n,nb', [(2, 700, 128), (3, 1000, 128), (4, 1100, 256)])
def test_block_cyclic_matches_dense_oracle(world, n, nb):
    from oracle import likelihood as L, matern
    from conftest import RankResults
    out = RankResults()          # no multiprocessing.Manager: it would fork() this process
    mp.spawn(_worker, args=(world, _free_port(), n, nb, out), nprocs=world, join=True)
    pts, z, X = _problem(n)
    Ko = L.MixedCorrelation(matern.generate_dense_correlation(pts, 0.1, 2.5), 'cholesky')
    ld = Ko.logdet(0.3)
    sol = Ko.solve(0.3, numpy.c_[X, z])
    sig = L.ProfileLikelihood.find_optimal_sigma(z, X, Ko, 0.3)
    lp = L.ProfileLikelihood.log_likelihood(z, X, Ko, False, [sig, 0.3])
    dK = matern.matern_derivative_rho(pts, 0.1, 2.5)
    Kinv = numpy.linalg.inv(Ko.K + 0.3 * numpy.eye(n))
    grad = [lp, L.ProfileLikelihood.log_likelihood_der1_eta(z, X, Ko, numpy.log10(0.3)),
            L.ProfileLikelihood.log_likelihood_der1_rho(z, X, Ko, dK, 0.3)]
    seen = []
    for r in range(world):
        res = out[r]
        seen += res['rows']
        assert res['grid'] == (1, world)
        assert abs(res['traces'][0] - numpy.trace(Kinv)) <= 1e-9 * numpy.trace(Kinv)
        assert abs(res['traces'][1] - numpy.sum(Kinv * dK)) <= 1e-9 * abs(numpy.sum(Kinv * dK))
        assert numpy.max(numpy.abs(numpy.array(res['grad']) - grad) / numpy.abs(grad)) <= 1e-9
        assert abs(res['logdet'] - ld) <= 1e-10 * abs(ld)
        assert numpy.max(numpy.abs(res['solve'] - sol)) <= 1e-9 * numpy.max(numpy.abs(sol))
        assert numpy.max(numpy.abs(res['vec'] - sol[:, -1])) <= 1e-9 * numpy.max(numpy.abs(sol))
        assert abs(res['lp'][0] - lp) <= 1e-10 * abs(lp) and abs(res['lp'][1] - sig) <= 1e-10
        assert res['raised'] and res['recv'] > 0
    assert sorted(seen) == list(range((n + nb - 1) // nb))           # every block column / row has exactly one owner


def test_snake_ownership_balances_the_trailing_work():
    sys.path.insert(0, os.path.join(ROOT, 'gaussian-process-param-estimation_b200'))
    from gaussian_proc._blockcyclic import process_grid, snake_owner
    assert [process_grid(w) for w in (1, 2, 4, 8)] == [(1, 1), (1, 2), (1, 4), (1, 8)]
    assert [snake_owner(j, 4) for j in range(10)] == [0, 1, 2, 3, 3, 2, 1, 0, 0, 1]
    NB, P = 196, 8                                                  # n = 100 000, nb = 512 on 8 GPUs
    work = numpy.zeros(P)
    for j in range(NB):
        work[snake_owner(j, P)] += (NB - j) ** 2                   # trailing-update work ~ sum over steps of the column height
    assert work.max() / work.mean() <= 1.01                         # plain cyclic order: 1.06
