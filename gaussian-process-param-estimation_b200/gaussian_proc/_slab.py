"""
Row-slab sparse engine: ONE sparse likelihood evaluation (n = 2^20 in BASELINE configs[3]) on several GPUs.

The reference evaluates it on one host through imate SLQ and scipy CG over the whole matrix
(/root/reference/gaussian_proc/_mixed_correlation/mixed_correlation.py:193-209, 263-299). Here the rows of the spatially
ordered operator (Hilbert / Z-order, the order SparseEngine uses anyway) are cut into `world` contiguous slabs of 16-row
blocks, one per GPU (one process per GPU). Every n-vector of the Krylov iterations exists only as slabs; what crosses
GPUs is

* the HALO rows of the SpMM input: a block-column index carries (owner rank, row in the owner's slab) and the kernel
  loads the row from the owner's memory over NVLink (the peers' arenas are mapped through CUDA IPC, csrc/gp_peer.cu) -
  the spatial order keeps most columns of a slab inside the slab;
* the column reductions (Lanczos alpha / beta, CG p^T A p / r^T r, dots, Gram blocks): pushed into every rank's mailbox
  and added in rank order inside the reduction kernel itself (csrc/gp_peer.cuh) - bit-identical on every rank, so the
  ranks run the same recurrences, take the same stopping decisions and return the same numbers.

torch.distributed is used for the set-up (exchange of 64-byte IPC handles, one barrier) and for the host API's
all-gather of a solution; nothing on the evaluation path calls a library collective.
"""

import ctypes

import numpy

from . import _device as dev
from ._sparse import SparseEngine, DeviceCSR, DeviceRowBlocks, check, _p

lib = dev.lib

__all__ = ['SlabSparseEngine', 'slab_geometry', 'PeerArena']


def slab_geometry(n, world, rank, block_rows=16):
    """(slab, first, last): uniform slabs of `slab` rows (a multiple of the row-block height), this rank's rows are
    [first, last) of the ordered operator. The last slab may be short."""
    per = (n + world - 1) // world
    slab = (per + block_rows - 1) // block_rows * block_rows
    first = min(n, rank * slab)
    return slab, first, min(n, first + slab)


class PeerArena(object):
    """This rank's arena + the mapped arenas of the peers (see csrc/gp_peer.cu). One per process and slab size, reused by
    every engine (an optimiser builds a new operator per rho)."""

    _cache = {}

    def __init__(self, rank, world, nloc_max):
        torch = dev.require_cuda()
        self.rank, self.world, self.nloc_max = int(rank), int(world), int(nloc_max)
        self.ctx = lib.gp_peer_create(self.rank, self.world, self.nloc_max)
        if not self.ctx:
            raise MemoryError('gp_peer_create failed (arena for slabs of %d rows)' % nloc_max)
        hb = int(lib.gp_peer_handle_bytes())
        mine = numpy.zeros(hb, dtype=numpy.uint8)
        check(lib.gp_peer_handle(self.ctx, dev.host_ptr(mine)), 'gp_peer_handle')
        if self.world > 1:
            import torch.distributed as dist
            d = torch.device('cuda', torch.cuda.current_device()) if dist.get_backend() == 'nccl' else torch.device('cpu')
            t = torch.from_numpy(mine).to(d)
            bufs = [torch.empty_like(t) for _ in range(self.world)]
            dist.all_gather(bufs, t)
            handles = numpy.ascontiguousarray(numpy.stack([b.cpu().numpy() for b in bufs]))
            check(lib.gp_peer_connect(self.ctx, dev.host_ptr(handles)), 'gp_peer_connect (CUDA IPC)')
            dist.barrier()        # every arena is zeroed and mapped everywhere before the first kernel touches one
        else:
            check(lib.gp_peer_connect(self.ctx, None), 'gp_peer_connect')

    @classmethod
    def get(cls, rank, world, nloc_max, tag='main'):
        key = (int(rank), int(world), tag)
        cur = cls._cache.get(key)
        if cur is None or cur.nloc_max < nloc_max:
            # (a larger arena replaces the old one on every rank at the same call: nloc_max depends on n and world only)
            if cur is not None:
                cur.close()
            cur = cls(rank, world, nloc_max)
            cls._cache[key] = cur
        return cur

    def vec(self, k, rows, B):
        """exchange vector k of this rank as an address (rows x B doubles live there)"""
        if rows > self.nloc_max or B > 32:
            raise ValueError('exchange vector too small')
        return ctypes.c_void_p(lib.gp_peer_vec(self.ctx, k))

    def barrier(self):
        check(lib.gp_peer_barrier(self.ctx, dev.stream_ptr()), 'gp_peer_barrier')

    def check_error(self):
        if lib.gp_peer_error(self.ctx, dev.stream_ptr()) != 0:
            raise RuntimeError('row-slab engine: a peer exchange timed out (the ranks issued different kernel sequences)')

    def close(self):
        """collective: every rank closes its arena at the same point (nobody may still gather from it)"""
        if self.ctx:
            if self.world > 1:
                import torch.distributed as dist
                dev.torch.cuda.synchronize()
                dist.barrier()
            lib.gp_peer_destroy(self.ctx)
            self.ctx = None
            if self.world > 1:
                dist.barrier()


class SlabSparseEngine(SparseEngine):
    """SparseEngine on this rank's slab of rows. Same interface and - to rounding of the reordered sums - the same numbers
    as SparseEngine on one GPU; every rank of the group must make the same calls in the same order."""

    def __init__(self, K, imate_method='slq', imate_options=None, rank=None, world=None):
        from ._distributed import rank_world
        r, w = rank_world()
        self.rank = int(r if rank is None else rank)
        self.world = int(w if world is None else world)
        if not isinstance(K, (DeviceCSR, DeviceRowBlocks)) or K.order is None:
            raise ValueError('the row-slab engine needs a device handle from generate_sparse_operator / '
                             'generate_sparse_correlation(..., device=True).')
        opts = dict(imate_options or {})
        if int(opts.get('block_rows', 16)) != 16:
            raise ValueError('the row-slab engine works on 16-row blocks.')
        opts['block_rows'] = 16
        self.slab, self.first_row, self.last_row = slab_geometry(K.n, self.world, self.rank)
        if self.last_row <= self.first_row:
            raise ValueError('n = %d is too small for %d slabs of 16-row blocks.' % (K.n, self.world))
        part = getattr(K, 'row_slab', None)         # a handle generated with row_slab=(rank, world) holds only this slab
        if part is not None and part != (self.rank, self.world, self.first_row, self.last_row):
            raise ValueError('this handle holds the rows of slab %r, not those of rank %d of %d.'
                             % (part, self.rank, self.world))
        if isinstance(K, DeviceRowBlocks) and (K.first_row, K.last_row) != (self.first_row, self.last_row):
            raise ValueError('a directly generated operator must be generated for this slab (row_slab=(rank, world)).')
        self.peer = PeerArena.get(self.rank, self.world, self.slab)
        # the first SLQ block runs on a side stream next to the CG for [X z] (SparseEngine.prefetch_slq): its exchanges go
        # through a second arena - mailboxes and exchange vectors of two concurrent Krylov runs must not mix
        self.peer_side = (PeerArena.get(self.rank, self.world, self.slab, tag='side')
                          if bool(opts.get('overlap', True)) else None)
        SparseEngine.__init__(self, K, imate_method, opts, probe_range=None)

    # ---- operator -------------------------------------------------------------------------------------------------
    def _halo_ws(self):
        """scratch of the halo statistics: two counters + one bit per row of the operator"""
        return dev.torch.empty(2 + (self.n + 63) // 64 + 1, dtype=dev.torch.int64, device='cuda')

    def _build_blocked(self, K, R):
        torch = dev.torch
        n, r0, r1 = self.n, self.first_row, self.last_row
        nloc = r1 - r0
        s = dev.stream_ptr()
        if isinstance(K, DeviceRowBlocks):         # generated as the row blocks of this slab: only the column encoding is left
            total = K.bidx.numel()
            if K.encoded is None:
                halo = (ctypes.c_int64 * 2)()
                check(lib.gp_slab_encode_columns(_p(K.bidx), total, self.slab, self.rank, n, halo, _p(self._halo_ws()), s),
                      'gp_slab_encode_columns')
                K.encoded = (self.slab, self.rank, int(halo[0]), int(halo[1]))
            elif K.encoded[:2] != (self.slab, self.rank):
                raise ValueError('this operator was encoded for another slab geometry')
            self.halo_blocks, self.halo_rows, self.total_blocks = K.encoded[2], K.encoded[3], total
            self.halo_fraction = self.halo_blocks / float(max(total, 1))
            self.R, self.rows = R, nloc
            self.blocked = (K.bptr, K.bidx, K.bvals, K.bdvals)
            self.fill_ratio = total * R / float(max(K.nnz, 1))
            self.order, self.inv_order = K.order, K.inv_order
            self._my_rows = K.order[r0:r1]
            return
        inv = getattr(K, 'inv_order', None)
        if inv is None:
            inv = torch.empty(n, dtype=torch.int32, device='cuda')
            check(lib.gp_inverse_permutation(_p(K.order), n, _p(inv), s), 'gp_inverse_permutation')
        my_rows = K.order[r0:r1]                              # original ids of this rank's rows (contiguous int32 view)
        nrb = (nloc + R - 1) // R
        nblk = torch.empty(nrb, dtype=torch.int32, device='cuda')
        flag = torch.zeros(1, dtype=torch.int32, device='cuda')
        check(lib.gp_bcsr_count(R, nloc, _p(my_rows), _p(inv), _p(K.indptr), _p(K.indices), _p(nblk), _p(flag), s),
              'gp_bcsr_count')
        if not K.sorted_rows and int(flag.item()) != 0:
            K.canonicalize()
            check(lib.gp_bcsr_count(R, nloc, _p(my_rows), _p(inv), _p(K.indptr), _p(K.indices), _p(nblk), _p(flag), s),
                  'gp_bcsr_count')
        bptr = torch.empty(nrb + 1, dtype=torch.int64, device='cuda')
        check(lib.gp_scan_counts(_p(nblk), nrb, _p(bptr), s), 'gp_scan_counts')
        total = int(bptr[-1].item())
        bidx = torch.empty(total, dtype=torch.int32, device='cuda')
        bvals = torch.empty(total * R, dtype=torch.float64, device='cuda')
        bdvals = torch.empty(total * R, dtype=torch.float64, device='cuda') if K.ddata is not None else None
        check(lib.gp_bcsr_fill(R, nloc, _p(my_rows), _p(inv), _p(K.indptr), _p(K.indices), _p(K.data),
                               _p(K.ddata) if K.ddata is not None else None, _p(bptr), total, _p(bidx), _p(bvals),
                               _p(bdvals) if bdvals is not None else None, s), 'gp_bcsr_fill')
        halo = (ctypes.c_int64 * 2)()   # block-columns whose row of X lives on another rank, and the distinct rows among them
        check(lib.gp_slab_encode_columns(_p(bidx), total, self.slab, self.rank, n, halo, _p(self._halo_ws()), s),
              'gp_slab_encode_columns')
        self.halo_blocks, self.halo_rows, self.total_blocks = int(halo[0]), int(halo[1]), total
        self.halo_fraction = halo[0] / float(max(total, 1))
        self.R = R
        self.rows = nloc
        self.blocked = (bptr, bidx, bvals, bdvals)
        self.fill_ratio = total * R / float(max(K.nnz, 1)) * (1 if getattr(K, 'row_slab', None) else self.world)
        self.order, self.inv_order = K.order, inv
        self._my_rows = my_rows

    def spmm(self, eta, X_dev, derivative=False):
        torch = dev.torch
        B = X_dev.shape[1]
        X_dev = X_dev.contiguous()
        Y = torch.empty_like(X_dev)
        bptr, bidx, bvals, bdvals = self.blocked
        check(lib.gp_slab_spmm(self.peer.ctx, _p(bptr), _p(bidx), _p(bdvals if derivative else bvals), self.rows, float(eta),
                               _p(X_dev), B, _p(Y), dev.stream_ptr()), 'gp_slab_spmm')
        return Y

    def to_op(self, X_dev):
        """this rank's rows (operator order) of a FULL (n x B) block given in the caller's row order"""
        X_dev = X_dev.contiguous()
        B = X_dev.numel() // self.n
        Y = dev.torch.empty((self.rows, B), dtype=dev.torch.float64, device='cuda')
        check(lib.gp_gather_rows(_p(X_dev), _p(self._my_rows), self.rows, B, _p(Y), dev.stream_ptr()), 'gp_gather_rows')
        return Y

    def from_op(self, X_dev):
        """the FULL (n x B) block in the caller's row order from the ranks' slabs (host API only: all-gather)"""
        torch = dev.torch
        B = X_dev.shape[1]
        pad = torch.zeros((self.slab, B), dtype=torch.float64, device='cuda')
        pad[:self.rows].copy_(X_dev)
        if self.world > 1:
            import torch.distributed as dist
            full = torch.empty((self.world * self.slab, B), dtype=torch.float64, device='cuda')
            dist.all_gather_into_tensor(full, pad)
        else:
            full = pad
        Y = torch.empty((self.n, B), dtype=torch.float64, device='cuda')
        check(lib.gp_gather_rows(_p(full), _p(self.inv_order), self.n, B, _p(Y), dev.stream_ptr()), 'gp_gather_rows')
        return Y

    def probes(self, first, B, out=None):
        torch = dev.torch
        V = torch.empty((self.rows, B), dtype=torch.float64, device='cuda') if out is None else out
        check(lib.gp_rademacher(_p(V), self.rows, B, int(self.opt['seed']), int(first), _p(self._my_rows),
                                dev.stream_ptr()), 'gp_rademacher')
        return V

    # ---- Krylov drivers ---------------------------------------------------------------------------------------------
    def _lanczos_launch(self, eta, V, m, basis=None, alpha=None, beta=None, side=False):
        torch = dev.torch
        B = V.shape[1]
        alpha = torch.empty((m, B), dtype=torch.float64, device='cuda') if alpha is None else alpha
        beta = torch.empty((m, B), dtype=torch.float64, device='cuda') if beta is None else beta
        bptr, bidx, bvals, _ = self.blocked
        peer = self.peer_side if side else self.peer
        check(lib.gp_slab_lanczos(peer.ctx, _p(bptr), _p(bidx), _p(bvals), self.rows, float(eta), _p(V), B, m, _p(alpha),
                                  _p(beta), _p(basis) if basis is not None else None, _p(self._workspace(B, side)),
                                  dev.stream_ptr()), 'gp_slab_lanczos')
        return alpha, beta

    def col_dot(self, X, Y):
        torch = dev.torch
        B = X.shape[1]
        out = torch.empty(32, dtype=torch.float64, device='cuda')
        check(lib.gp_slab_col_dot(self.peer.ctx, _p(X), _p(Y), self.rows, B, _p(out), _p(self._workspace(B)),
                                  dev.stream_ptr()), 'gp_slab_col_dot')
        return out[:B].cpu().numpy()

    def solve_dev(self, eta, R_dev):
        torch = dev.torch
        B = R_dev.shape[1]
        X = torch.empty_like(R_dev)
        it = ctypes.c_int64()
        bptr, bidx, bvals, _ = self.blocked
        rc = lib.gp_slab_cg_solve(self.peer.ctx, _p(bptr), _p(bidx), _p(bvals), self.rows, float(eta), _p(R_dev), _p(X), B,
                                  float(self.opt['cg_tol']), int(self.opt['cg_maxiter']), ctypes.byref(it),
                                  _p(self._workspace(B)), dev.stream_ptr())
        check(rc, 'gp_slab_cg_solve')
        self.last_cg_iterations = it.value
        if rc == 2:
            raise numpy.linalg.LinAlgError(
                'K + eta*I (eta=%g) is not positive definite (CG met p^T A p <= 0). The thresholded Matern matrix is '
                'indefinite; use a larger eta (reference: _generate_sparse_correlation.pyx:516-523).' % eta)
        if rc == 1:
            raise numpy.linalg.LinAlgError('CG did not converge in %d iterations (eta=%g)' % (it.value, eta))
        return X

    def gram(self, X_dev, Y_dev):
        torch = dev.torch
        B = X_dev.shape[1]
        if '_gram' not in self._ws:
            self._ws['_gram'] = torch.empty(lib.gp_gram_workspace_bytes(16) // 8, dtype=torch.float64, device='cuda')
        out = torch.empty(B * B, dtype=torch.float64, device='cuda')
        s = dev.stream_ptr()
        check(lib.gp_gram_skinny(_p(X_dev), _p(Y_dev), self.rows, B, _p(out), _p(self._ws['_gram']), s), 'gp_gram_skinny')
        check(lib.gp_peer_allreduce(self.peer.ctx, _p(out), B * B, s), 'gp_peer_allreduce')
        return out.cpu().numpy().reshape(B, B)

    def fused(self, eta, X, z, traceinv=True, drho=True, cubic=False):
        out = SparseEngine.fused(self, eta, X, z, traceinv=traceinv, drho=drho, cubic=cubic)
        self.peer.check_error()
        if self.peer_side is not None:
            self.peer_side.check_error()
        return out

    def trace_K(self):
        raise NotImplementedError('trace_K is a host reduction of the single-GPU engine.')
