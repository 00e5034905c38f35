"""
Multi-GPU plumbing (one process per GPU, torch.distributed): the hot path shards only where it is naturally
parallel -- independent (eta, rho) cells of a hyper-parameter grid, and independent probe vectors of the stochastic
estimators -- so the only collectives are a final gather of per-cell results and an all-reduce of
(count, sum, sum of squares). NCCL on GPUs; the same code runs on gloo for the CPU tests.
"""

import numpy

__all__ = ['is_distributed', 'rank_world', 'allreduce_sum', 'allgather_rows', 'partition_cells']


def _dist():
    import torch.distributed as dist
    return dist


def is_distributed():
    dist = _dist()
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def rank_world():
    dist = _dist()
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def _device_for_backend():
    import torch
    dist = _dist()
    return torch.device('cuda', torch.cuda.current_device()) if dist.get_backend() == 'nccl' else torch.device('cpu')


def allreduce_sum(values):
    """Element-wise sum of a small float64 host vector over all ranks (identity when not distributed)."""
    values = numpy.asarray(values, dtype=numpy.float64)
    if not is_distributed():
        return values
    import torch
    t = torch.from_numpy(values.copy()).to(_device_for_backend())
    _dist().all_reduce(t)
    return t.cpu().numpy()


def allgather_rows(rows):
    """Concatenates per-rank (k_r x w) float64 arrays in rank order; k_r may differ between ranks."""
    rows = numpy.ascontiguousarray(rows, dtype=numpy.float64).reshape(-1, rows.shape[-1] if rows.ndim > 1 else 1)
    if not is_distributed():
        return rows
    import torch
    dist = _dist()
    world = dist.get_world_size()
    d = _device_for_backend()
    cnt = torch.tensor([rows.shape[0]], dtype=torch.int64, device=d)
    counts = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(counts, cnt)
    kmax = int(max(int(c.item()) for c in counts))
    pad = torch.zeros((kmax, rows.shape[1]), dtype=torch.float64, device=d)
    pad[:rows.shape[0]] = torch.from_numpy(rows).to(d)
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    return numpy.concatenate([b[:int(c.item())].cpu().numpy() for b, c in zip(bufs, counts)], axis=0)


def partition_cells(num_rho, world, rank):
    """Contiguous rho-groups per rank: the correlation matrix is generated once per rho and reused for
    every eta of that group. Returns the [begin, end) range of rho indices owned by `rank`."""
    base, rem = divmod(num_rho, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)
