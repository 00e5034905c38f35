"""
Sparse (kernel-threshold) path: device CSR generation and the stochastic estimators behind MixedCorrelation for a
sparse K (SpMM + batched Lanczos / CG in csrc/gp_sparse.cu, csrc/gp_sparse_la.cu).

Reference behaviour restated here:
  generate_sparse_correlation  gaussian_proc/generate_correlation/_generate_sparse_correlation.pyx:472-594
  logdet / traceinv (slq, hutchinson), solve (CG, tol 1e-6)
                               gaussian_proc/_mixed_correlation/mixed_correlation.py:193-209,263-268;
                               _linear_solver.py:49-68
imate (absent, unpinned dependency) is followed through its documented estimator: Rademacher probes, Lanczos
quadrature, min/max number of samples, relative error tolerance at a confidence level. parity unpinned:
the reference's own 'slq' branches are dead code (mixed_correlation.py:141,207,266); the confidence band is checked
against exact values from the oracle in tests/test_gpu_sparse.py.
"""

import ctypes

import numpy
import scipy.linalg
import scipy.sparse
import scipy.special

from . import _device as dev
from ._device import lib, check

__all__ = ['DeviceCSR', 'DeviceRowBlocks', 'SparseEngine', 'generate_sparse_correlation', 'generate_sparse_operator',
           'estimate_kernel_threshold']

# imate's documented defaults for the stochastic estimators
DEFAULTS = dict(min_num_samples=10, max_num_samples=50, error_rtol=1e-2, error_atol=None, confidence_level=0.95,
                lanczos_degree=20, seed=0, batch=16, cg_tol=1e-6, cg_maxiter=2000, block_rows=16, reuse_lanczos=True, overlap=True,
                shift_reuse=True, solve_degree=None, eager_rhs_basis=False)


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


_RHS_DEVICE_CACHE = []      # [(key, device block, X, z)]: see SparseEngine._rhs_block


def estimate_kernel_threshold(matrix_size, dimension, density, correlation_scale, nu):
    """tau for the requested density (_generate_sparse_correlation.pyx:294-413); ValueError when density*n < 1."""
    tau = ctypes.c_double()
    scale = dev.host_f64(correlation_scale)
    rc = lib.gp_kernel_threshold(int(matrix_size), int(dimension), float(density), dev.host_ptr(scale), float(nu),
                                 ctypes.byref(tau))
    if rc == -10:
        raise ValueError(
            'Adjacency: %0.2f. Correlation matrix will become identity since kernel radius is less than grid size. '
            'To increase adjacency, consider increasing density or correlation_scale.' % (density * matrix_size))
    check(rc, 'gp_kernel_threshold')
    return tau.value


class DeviceCSR(object):
    """CSR (int32 indptr / int32 indices, float64 data) resident on the GPU; `ddata` optionally holds d/d(rho) of every
    stored entry on the same pattern. The device generator leaves the rows in generation order (``sorted_rows`` False):
    the device operators do not need more; ``canonicalize()`` sorts every row by column - the reference's canonical CSR,
    bit-exact pattern - and is called by ``to_scipy()`` / whenever a consumer needs sorted rows."""

    def __init__(self, n, indptr, indices, data, ddata=None, kernel_threshold=None, order=None, sorted_rows=True):
        self.n = int(n)
        self.indptr, self.indices, self.data, self.ddata = indptr, indices, data, ddata
        self.kernel_threshold = kernel_threshold
        self.order = order      # optional spatially local ordering of the rows (device int32), see SparseEngine
        self.sorted_rows = bool(sorted_rows)

    def canonicalize(self):
        """sort every row by column index, in place (no-op when already sorted)"""
        if not self.sorted_rows:
            torch = dev.torch
            flags = torch.zeros(2, dtype=torch.int32, device='cuda')
            rc = lib.gp_csr_sort_rows(self.n, _p(self.indptr), _p(self.indices), _p(self.data),
                                      _p(self.ddata) if self.ddata is not None else None, _p(flags), dev.stream_ptr())
            check(rc, 'gp_csr_sort_rows')
            self.sorted_rows = True
        return self

    @property
    def shape(self):
        return (self.n, self.n)

    @property
    def nnz(self):
        return int(self.data.shape[0])

    @classmethod
    def from_scipy(cls, K):
        torch = dev.require_cuda()
        K = scipy.sparse.csr_matrix(K)
        K.sort_indices()
        if K.shape[0] != K.shape[1]:
            raise ValueError('K should be a square matrix.')
        return cls(K.shape[0], torch.from_numpy(K.indptr.astype(numpy.int32)).cuda(),
                   torch.from_numpy(K.indices.astype(numpy.int32)).cuda(),
                   torch.from_numpy(K.data.astype(numpy.float64)).cuda())

    def to_scipy(self):
        self.canonicalize()
        return scipy.sparse.csr_matrix((self.data.cpu().numpy(), self.indices.cpu().numpy(), self.indptr.cpu().numpy()),
                                       shape=(self.n, self.n))


class DeviceRowBlocks(object):
    """A sparse correlation matrix generated DIRECTLY as the operator the likelihood works on: 16-row blocks of the
    spatially ordered matrix (bptr, bidx, bvals[, bdvals] in the SpMM's fragment order, csrc/gp_sparse_la.cu), without a CSR.
    Rows [first_row, last_row) of the ordered operator are present (everything, or the slab of one GPU)."""

    R = 16

    def __init__(self, n, bptr, bidx, bvals, bdvals, order, inv_order, nnz, kernel_threshold, first_row, last_row,
                 row_slab=None):
        self.n = int(n)
        self.bptr, self.bidx, self.bvals, self.bdvals = bptr, bidx, bvals, bdvals
        self.ddata = bdvals                    # (engines ask `K.ddata is not None` for "carries dK/drho")
        self.order, self.inv_order = order, inv_order
        self.nnz = int(nnz)                    # entries in the pattern of the rows present
        self.kernel_threshold = kernel_threshold
        self.first_row, self.last_row = int(first_row), int(last_row)
        self.row_slab = row_slab
        self.sorted_rows = True
        self.encoded = None                    # set by the row-slab engine once the columns carry (owner, local row)

    @property
    def shape(self):
        return (self.n, self.n)

    def to_scipy(self):
        """scipy.sparse.csr_matrix of the rows present, original row / column order (host reconstruction; rare path)"""
        if self.encoded is not None:
            raise ValueError('the columns of this operator were re-encoded for the row-slab engine')
        R = self.R
        bptr = self.bptr.cpu().numpy()
        bidx = self.bidx.cpu().numpy().astype(numpy.int64)
        vals = self.bvals.cpu().numpy().reshape(-1, 2, 8, 4)       # [group of 4 slots][fragment][row in fragment][slot in group]
        order = self.order.cpu().numpy().astype(numpy.int64)
        nrb = bptr.size - 1
        slot_block = numpy.repeat(numpy.arange(nrb), numpy.diff(bptr))
        v = vals.transpose(0, 3, 1, 2).reshape(-1, R)               # [slot][row of the block]
        rows_op = self.first_row + slot_block[:, None] * R + numpy.arange(R)[None, :]
        keep = (v != 0.0) & (rows_op < self.last_row)
        rows = order[numpy.minimum(rows_op, self.n - 1)][keep]
        cols = order[numpy.broadcast_to(bidx[:, None], v.shape)[keep]]
        M = scipy.sparse.csr_matrix((v[keep], (rows, cols)), shape=(self.n, self.n))
        M.sort_indices()
        return M


_POINTS_DEVICE_CACHE = []      # [(key, points array, device points, order, position)]


def _device_points(points):
    """(device copy, spatial order, position in the order) of a host point set. None of the three depends on the
    hyper-parameters, and an optimiser or a sweep generates K for the same points at every step: the last two point sets
    are kept (keyed by content, _device.host_key; arrays that cannot be written to are recognised without a digest).
    order = stable sort of the Hilbert / Z-order keys over the bounding box (own kernels, csrc/gp_index.cu): the
    deterministic, spatially local row order of the row-blocked operator; position = its inverse permutation."""
    torch = dev.torch
    key = dev.host_key(points)
    for ent in _POINTS_DEVICE_CACHE:
        if ent[0] == key:
            return ent[2], ent[3], ent[4]
    n, d = points.shape
    s = dev.stream_ptr()
    dpts = torch.from_numpy(numpy.array(points)).cuda()
    box = numpy.empty(2 * d)
    box_ws = torch.empty(148 * 2 * d + 2 * d, dtype=torch.float64, device='cuda')
    check(lib.gp_points_bbox(_p(dpts), n, d, dev.host_ptr(box), _p(box_ws), s), 'gp_points_bbox')
    lo, hi = dev.host_f64(box[:d]), dev.host_f64(box[d:])
    keys = torch.empty(n, dtype=torch.int64, device='cuda')
    check(lib.gp_spatial_keys(_p(dpts), n, d, dev.host_ptr(lo), dev.host_ptr(hi), _p(keys), s), 'gp_spatial_keys')
    order = torch.empty(n, dtype=torch.int32, device='cuda')
    sort_ws = torch.empty(lib.gp_sort_workspace_bytes(n) // 8 + 8, dtype=torch.float64, device='cuda')
    check(lib.gp_sort_keys_u64(_p(keys), n, 64, _p(order), _p(sort_ws), s), 'gp_sort_keys_u64')
    pos = torch.empty(n, dtype=torch.int32, device='cuda')
    check(lib.gp_inverse_permutation(_p(order), n, _p(pos), s), 'gp_inverse_permutation')
    _POINTS_DEVICE_CACHE.append((key, points, dpts, order, pos))
    del _POINTS_DEVICE_CACHE[:-2]
    return dpts, order, pos


def generate_sparse_correlation(points, correlation_scale, nu, density, verbose=False, device=False,
                                with_derivative=False, kernel_threshold=None, sort_rows=None, row_slab=None):
    """Device generator behind the reference's generate_sparse_correlation. Returns scipy.sparse.csr_matrix (or a
    DeviceCSR when ``device``). ``sort_rows``: canonical (column-sorted) rows right away; default: yes for the SciPy
    result, deferred (DeviceCSR.canonicalize) for a device handle.
    ``row_slab`` = (rank, world) (device handle only): generate only this rank's slab of rows of the spatially ordered
    operator (gaussian_proc/_slab.py); the other rows of the handle stay empty."""
    sort_rows = (not device) if sort_rows is None else bool(sort_rows)
    torch = dev.require_cuda()
    points = numpy.ascontiguousarray(points, dtype=numpy.float64)
    scale = dev.host_f64(correlation_scale)
    n, d = points.shape
    tau = estimate_kernel_threshold(n, d, density, scale, nu) if kernel_threshold is None else float(kernel_threshold)
    if with_derivative and not numpy.all(scale == scale[0]):
        raise ValueError('d/d(correlation_scale) is defined for an isotropic correlation_scale.')
    if row_slab is not None and not device:
        raise ValueError('row_slab needs device=True (a slab is not a complete matrix).')
    s = dev.stream_ptr()
    dpts, order, pos = _device_points(points)
    rows = None
    if row_slab is not None:
        from ._slab import slab_geometry
        rank, world = int(row_slab[0]), int(row_slab[1])
        _, first, last = slab_geometry(n, world, rank)
        rows = (pos, first, last)
    ws = torch.empty(lib.gp_sparse_workspace_bytes(n, d) // 8 + 8, dtype=torch.float64, device='cuda')
    indptr = torch.empty(n + 1, dtype=torch.int32, device='cuda')
    nnz = ctypes.c_int64()
    if rows is None:
        rc = lib.gp_matern_sparse_count(_p(dpts), dev.host_ptr(points), n, d, dev.host_ptr(scale), float(nu), tau, _p(ws),
                                        _p(indptr), ctypes.byref(nnz), s)
    else:
        rc = lib.gp_matern_sparse_count_rows(_p(dpts), dev.host_ptr(points), n, d, dev.host_ptr(scale), float(nu), tau, _p(ws),
                                             _p(indptr), ctypes.byref(nnz), _p(rows[0]), rows[1], rows[2], s)
    check(rc, 'gp_matern_sparse_count')
    indices = torch.empty(nnz.value, dtype=torch.int32, device='cuda')
    data = torch.empty(nnz.value, dtype=torch.float64, device='cuda')
    ddata = torch.empty(nnz.value, dtype=torch.float64, device='cuda') if with_derivative else None
    if rows is None:
        rc = lib.gp_matern_sparse_fill(_p(dpts), dev.host_ptr(points), n, d, dev.host_ptr(scale), float(nu), tau, _p(ws),
                                       _p(indptr), _p(indices), _p(data), _p(ddata) if ddata is not None else None,
                                       1 if sort_rows else 0, s)
    else:
        rc = lib.gp_matern_sparse_fill_rows(_p(dpts), dev.host_ptr(points), n, d, dev.host_ptr(scale), float(nu), tau, _p(ws),
                                            _p(indptr), _p(indices), _p(data), _p(ddata) if ddata is not None else None,
                                            1 if sort_rows else 0, _p(rows[0]), rows[1], rows[2], s)
    check(rc, 'gp_matern_sparse_fill')
    K = DeviceCSR(n, indptr, indices, data, ddata, kernel_threshold=tau, order=order, sorted_rows=sort_rows)
    K.inv_order = pos
    if rows is not None:
        K.row_slab = (rank, world, rows[1], rows[2])       # only these rows of the ordered operator are present
    if verbose:
        print('Generated sparse correlation matrix using kernel threshold: %0.4f and sparse density: %0.2e.'
              % (tau, K.nnz / float(n) ** 2))
    return K if device else K.to_scipy()


def generate_sparse_operator(points, correlation_scale, nu, density, with_derivative=False, kernel_threshold=None,
                             row_slab=None, verbose=False):
    """The sparse correlation matrix of generate_sparse_correlation as a device OPERATOR for MixedCorrelation / SparseEngine:
    the row blocks are generated directly from the cell lists (csrc/gp_sparse.cu, sparse_blocks_kernel; same kernel
    arithmetic and pattern rule, no CSR in between), about half the time of CSR generation + block build. Falls back to that
    path (returns a DeviceCSR) for a matrix with entries inside the 8-ulp borderline band of the threshold, which the
    reference's rule decides in host arithmetic. ``row_slab`` = (rank, world): only this rank's slab (gaussian_proc/_slab.py)."""
    torch = dev.require_cuda()
    points = numpy.ascontiguousarray(points, dtype=numpy.float64)
    scale = dev.host_f64(correlation_scale)
    n, d = points.shape
    tau = estimate_kernel_threshold(n, d, density, scale, nu) if kernel_threshold is None else float(kernel_threshold)
    if with_derivative and not numpy.all(scale == scale[0]):
        raise ValueError('d/d(correlation_scale) is defined for an isotropic correlation_scale.')
    s = dev.stream_ptr()
    dpts, order, pos = _device_points(points)
    first, last, slab = 0, n, None
    if row_slab is not None:
        from ._slab import slab_geometry
        rank, world = int(row_slab[0]), int(row_slab[1])
        _, first, last = slab_geometry(n, world, rank)
        slab = (rank, world, first, last)
    ws = torch.empty(lib.gp_sparse_workspace_bytes(n, d) // 8 + 8, dtype=torch.float64, device='cuda')
    nrb = (last - first + 15) // 16
    nblk = torch.empty(nrb, dtype=torch.int32, device='cuda')
    stats = (ctypes.c_int64 * 2)()
    rc = lib.gp_matern_blocks_count(_p(dpts), n, d, dev.host_ptr(scale), float(nu), tau, _p(ws), _p(order), _p(pos), first, last,
                                    _p(nblk), stats, s)
    if rc == 3:
        return generate_sparse_correlation(points, correlation_scale, nu, density, verbose=verbose, device=True,
                                           with_derivative=with_derivative, kernel_threshold=tau, row_slab=row_slab)
    check(rc, 'gp_matern_blocks_count')
    bptr = torch.empty(nrb + 1, dtype=torch.int64, device='cuda')
    check(lib.gp_scan_counts(_p(nblk), nrb, _p(bptr), s), 'gp_scan_counts')
    total = int(bptr[-1].item())
    bidx = torch.empty(total, dtype=torch.int32, device='cuda')
    bvals = torch.empty(total * 16, dtype=torch.float64, device='cuda')
    bdvals = torch.empty(total * 16, dtype=torch.float64, device='cuda') if with_derivative else None
    check(lib.gp_matern_blocks_fill(_p(dpts), n, d, dev.host_ptr(scale), float(nu), tau, _p(ws), _p(order), _p(pos), first, last,
                                    _p(bptr), _p(bidx), _p(bvals), _p(bdvals) if bdvals is not None else None, s),
          'gp_matern_blocks_fill')
    K = DeviceRowBlocks(n, bptr, bidx, bvals, bdvals, order, pos, stats[0], tau, first, last, row_slab=slab)
    if verbose:
        print('Generated sparse correlation operator using kernel threshold: %0.4f and sparse density: %0.2e.'
              % (tau, K.nnz / float(n) / float(last - first)))
    return K


def lanczos_quadrature(alpha, beta, funcs, return_size=False):
    """Gauss quadrature of v^T f(A) v / ||v||^2 from the Lanczos tridiagonal: sum_k tau_k^2 f(theta_k) with theta the
    Ritz values and tau the first components of the Ritz vectors. A beta ~ 0 truncates the recurrence (invariant
    subspace reached)."""
    m = len(alpha)
    for j in range(m - 1):
        if not (beta[j] > 1e-12 * max(abs(alpha[0]), 1.0)):
            m = j + 1
            break
    theta, Y = scipy.linalg.eigh_tridiagonal(alpha[:m], beta[:m - 1]) if m > 1 else (alpha[:1], numpy.ones((1, 1)))
    w = Y[0, :] ** 2
    res = [float(numpy.sum(w * f(theta))) for f in funcs], float(theta.min())
    return res + (m,) if return_size else res


def lanczos_block_quadrature(alpha, beta, vnorm=None):
    """Batched host part of SLQ for a block of B Lanczos runs (alpha, beta: m x B): one stacked symmetric eigensolve of
    the B tridiagonals gives, per column, the Gauss quadratures of [log, 1/x, 1/x^2], the smallest Ritz value and - with
    ``vnorm`` (||v|| per column, or a scalar) - the coefficients of x = A^-1 v in the unnormalised Lanczos vectors
    (u_0 = v, u_j = beta_{j-1} q_j) together with the relative residual beta_m |y_m| of that Lanczos (= CG) solution.
    A column whose recurrence broke down (beta ~ 0: invariant subspace) is truncated there.
    Returns quad (B x 3), tmin (B), coef (m x B) or None, resid (B) or None."""
    m, B = alpha.shape
    ks = numpy.full(B, m, dtype=int)
    for c in range(B):
        small = numpy.nonzero(~(beta[:m - 1, c] > 1e-12 * max(abs(alpha[0, c]), 1.0)))[0]
        if small.size:
            ks[c] = small[0] + 1
    quad, tmin = numpy.empty((B, 3)), numpy.empty(B)
    want = vnorm is not None
    coef = numpy.zeros((m, B)) if want else None
    resid = numpy.zeros(B) if want else None
    vn = numpy.broadcast_to(numpy.asarray(vnorm, dtype=float), (B,)) if want else None
    for k in numpy.unique(ks):
        cols = numpy.nonzero(ks == k)[0]
        T = numpy.zeros((cols.size, k, k))
        idx = numpy.arange(k)
        T[:, idx, idx] = alpha[:k, cols].T
        if k > 1:
            T[:, idx[:-1], idx[1:]] = beta[:k - 1, cols].T
            T[:, idx[1:], idx[:-1]] = beta[:k - 1, cols].T
        theta, Y = numpy.linalg.eigh(T)
        w = Y[:, 0, :] ** 2
        with numpy.errstate(invalid='ignore', divide='ignore'):
            quad[cols, 0] = numpy.sum(w * numpy.log(theta), axis=1)
            quad[cols, 1] = numpy.sum(w / theta, axis=1)
            quad[cols, 2] = numpy.sum(w / theta ** 2, axis=1)
            tmin[cols] = theta.min(axis=1)
            if want:
                y = numpy.einsum('bij,bj->bi', Y, Y[:, 0, :] / theta)          # T^-1 e_1
                scale = numpy.empty((cols.size, k))
                scale[:, 0] = 1.0 / vn[cols]
                if k > 1:
                    scale[:, 1:] = 1.0 / beta[:k - 1, cols].T
                coef[:k, cols] = (vn[cols][:, None] * y * scale).T
                resid[cols] = numpy.abs(beta[k - 1, cols] * y[:, k - 1]) if k == m else 0.0
    return quad, tmin, coef, resid


class SparseEngine(object):
    """Stochastic logdet / traceinv and CG solves for K + eta I with K in device CSR."""

    def __init__(self, K, imate_method='slq', imate_options=None, probe_range=None):
        dev.require_cuda()
        if not isinstance(K, (DeviceCSR, DeviceRowBlocks)):
            K = DeviceCSR.from_scipy(K)
        self.K = K
        self.n = K.n
        self.rows = K.n            # rows of the vectors this engine works on (a slab of n on the row-slab engine, _slab.py)
        self.method = imate_method
        self.opt = dict(DEFAULTS)
        self.opt.update(imate_options or {})
        self._ws = {}
        self._krylov = {}          # shift-invariant Lanczos runs, see _probe_krylov / solve_rhs_block
        self._rhs_etas = {}
        self._slq_cache = {}
        self.last_info = {}
        # multi-GPU: (rank, world) -> this engine evaluates its slice of every round of probes; see _run_estimator
        self.probe_range = probe_range
        # Internally the operator works on the ROW-BLOCKED form of the symmetrically permuted matrix P K P^T (rows in
        # the generator's Z-order): 8 consecutive rows share one column list, so one gathered row of the probe block
        # serves 8 rows of K and four such block-columns are one DMMA.8x8x4 (csrc/gp_sparse_la.cu). Results do not depend
        # on the permutation: probes are hashed with ORIGINAL row ids. block_rows = 1 (or a K without an order, e.g. from
        # SciPy) keeps plain CSR in the original order.
        self.order = self.inv_order = None
        self.R = 1
        self.blocked = None
        R = int(self.opt.get('block_rows', 16))
        if R not in (1, 8, 16):
            raise ValueError('block_rows should be 8 or 16 (row-blocked, FP64 tensor-core SpMM) or 1 (plain CSR).')
        if isinstance(K, DeviceRowBlocks) and R != 16:
            raise ValueError('a directly generated operator (DeviceRowBlocks) has 16-row blocks.')
        if K.order is not None and R > 1:
            self._build_blocked(K, R)

    def _build_blocked(self, K, R):
        torch = dev.torch
        n = self.n
        if isinstance(K, DeviceRowBlocks):         # generated as row blocks: nothing to build
            if (K.first_row, K.last_row) != (0, n) or K.encoded is not None:
                raise ValueError('this operator holds one slab of rows (or its columns were encoded for the row-slab engine): '
                                 'use the row-slab engine (row_slabs=True).')
            self.R = R
            self.blocked = (K.bptr, K.bidx, K.bvals, K.bdvals)
            self.fill_ratio = K.bidx.numel() * R / float(max(K.nnz, 1))
            self.order, self.inv_order = K.order, K.inv_order
            return
        s = dev.stream_ptr()
        inv = getattr(K, 'inv_order', None)
        if inv is None:
            inv = torch.empty(n, dtype=torch.int32, device='cuda')
            check(lib.gp_inverse_permutation(_p(K.order), n, _p(inv), s), 'gp_inverse_permutation')
        nrb = (n + R - 1) // R
        nblk = torch.empty(nrb, dtype=torch.int32, device='cuda')
        flag = torch.zeros(1, dtype=torch.int32, device='cuda')
        check(lib.gp_bcsr_count(R, n, _p(K.order), _p(inv), _p(K.indptr), _p(K.indices), _p(nblk), _p(flag), s),
              'gp_bcsr_count')
        if not K.sorted_rows and int(flag.item()) != 0:
            # some row block is too wide for the hash path: its binary-search fallback needs canonical rows
            K.canonicalize()
            check(lib.gp_bcsr_count(R, n, _p(K.order), _p(inv), _p(K.indptr), _p(K.indices), _p(nblk), _p(flag), s),
                  'gp_bcsr_count')
        bptr = torch.empty(nrb + 1, dtype=torch.int64, device='cuda')
        check(lib.gp_scan_counts(_p(nblk), nrb, _p(bptr), s), 'gp_scan_counts')
        total = int(bptr[-1].item())
        bidx = torch.empty(total, dtype=torch.int32, device='cuda')
        bvals = torch.empty(total * R, dtype=torch.float64, device='cuda')
        bdvals = torch.empty(total * R, dtype=torch.float64, device='cuda') if K.ddata is not None else None
        check(lib.gp_bcsr_fill(R, n, _p(K.order), _p(inv), _p(K.indptr), _p(K.indices), _p(K.data),
                               _p(K.ddata) if K.ddata is not None else None, _p(bptr), total, _p(bidx), _p(bvals),
                               _p(bdvals) if bdvals is not None else None, s), 'gp_bcsr_fill')
        self.R = R
        self.blocked = (bptr, bidx, bvals, bdvals)
        self.fill_ratio = total * R / float(max(K.nnz, 1))     # stored values per original nonzero (>= 1)
        self.order, self.inv_order = K.order, inv            # device int32 row maps of to_op / from_op

    # ---- plumbing ----------------------------------------------------------------------------------------------
    def _workspace(self, B, side=False):
        """Krylov workspace for column blocks of width B; ``side``: the separate one of the prefetch stream (a Lanczos
        run there may overlap a CG / Lanczos run of the same width on the main stream)."""
        torch = dev.torch
        key = ('side', B) if side else B
        if key not in self._ws:
            self._ws[key] = torch.empty(lib.gp_krylov_workspace_bytes(self.rows, B) // 8 + 8, dtype=torch.float64, device='cuda')
        return self._ws[key]

    def spmm(self, eta, X_dev, derivative=False):
        """(K + eta I) X, or (dK/drho + eta I) X with ``derivative``, in OPERATOR space (rows in self.order when the
        operator is permuted)"""
        torch = dev.torch
        B = X_dev.shape[1]
        Y = torch.empty_like(X_dev)
        if self.blocked is not None:
            bptr, bidx, bvals, bdvals = self.blocked
            check(lib.gp_bcsr_spmm(self.R, _p(bptr), _p(bidx), _p(bdvals if derivative else bvals), self.n, float(eta),
                                   _p(X_dev), B, _p(Y), dev.stream_ptr()), 'gp_bcsr_spmm')
        else:
            K = self.K
            check(lib.gp_csr_spmm(_p(K.indptr), _p(K.indices), _p(K.ddata if derivative else K.data), self.n, float(eta),
                                  _p(X_dev), B, _p(Y), dev.stream_ptr()), 'gp_csr_spmm')
        return Y

    def _gather(self, X_dev, rmap):
        X_dev = X_dev.contiguous()
        Y = dev.torch.empty_like(X_dev)
        B = X_dev.numel() // self.n
        check(lib.gp_gather_rows(_p(X_dev), _p(rmap), self.n, B, _p(Y), dev.stream_ptr()), 'gp_gather_rows')
        return Y

    def to_op(self, X_dev):
        """rows of an (n x B) block from the caller's order into operator order"""
        return X_dev if self.order is None else self._gather(X_dev, self.order)

    def from_op(self, X_dev):
        return X_dev if self.order is None else self._gather(X_dev, self.inv_order)

    def probes(self, first, B, out=None):
        torch = dev.torch
        V = torch.empty((self.rows, B), dtype=torch.float64, device='cuda') if out is None else out
        rmap = self.K.order if self.order is not None else None
        check(lib.gp_rademacher(_p(V), self.rows, B, int(self.opt['seed']), int(first), _p(rmap) if rmap is not None else None,
                                dev.stream_ptr()), 'gp_rademacher')
        return V

    # ---- SLQ -----------------------------------------------------------------------------------------------------
    def _lanczos(self, eta, V, m, basis=None):
        alpha, beta = self._lanczos_launch(eta, V, m, basis)
        return alpha.cpu().numpy(), beta.cpu().numpy()

    def _lanczos_launch(self, eta, V, m, basis=None, alpha=None, beta=None, side=False):
        """enqueues the batched Lanczos run on torch's current stream; returns the device (alpha, beta)"""
        torch = dev.torch
        B = V.shape[1]
        alpha = torch.empty((m, B), dtype=torch.float64, device='cuda') if alpha is None else alpha
        beta = torch.empty((m, B), dtype=torch.float64, device='cuda') if beta is None else beta
        bp = _p(basis) if basis is not None else None
        if self.blocked is not None:
            bptr, bidx, bvals, _ = self.blocked
            check(lib.gp_bcsr_lanczos(self.R, _p(bptr), _p(bidx), _p(bvals), self.n, float(eta), _p(V), B, m, _p(alpha),
                                      _p(beta), bp, _p(self._workspace(B, side)), dev.stream_ptr()), 'gp_bcsr_lanczos')
        else:
            K = self.K
            check(lib.gp_lanczos(_p(K.indptr), _p(K.indices), _p(K.data), self.n, float(eta), _p(V), B, m, _p(alpha),
                                 _p(beta), bp, _p(self._workspace(B, side)), dev.stream_ptr()), 'gp_lanczos')
        return alpha, beta

    def _first_chunk(self):
        """(first probe id, width) of the first block of probes this rank evaluates in an estimator run"""
        o = self.opt
        B, hi = int(o['batch']), int(o['max_num_samples'])
        rank, world = self.probe_range if self.probe_range is not None else (0, 1)
        nb = min(B, hi)
        per = (nb + world - 1) // world
        my0 = rank * per
        chunks = self._chunks(my0, max(0, min(nb, my0 + per) - my0), B)
        return chunks[0] if chunks else None

    def prefetch_slq(self, eta):
        """Enqueues the first SLQ block of probes at `eta` on a side stream and returns at once, so that the batched CG
        for [X z] that follows on the main stream (HBM-bound, B = 8) overlaps the Lanczos run (L1-bound, B = 16).
        _slq_samples picks the result up."""
        torch = dev.torch
        if float(eta) in self._slq_cache or getattr(self, '_prefetched', None) is not None:
            return
        chunk = self._first_chunk()
        if chunk is None:
            return
        first, B = chunk
        if ('probes', first, B) in self._krylov and bool(self.opt.get('shift_reuse', True)):
            return                                 # already served by a kept Lanczos run
        m = int(self.opt['lanczos_degree'])
        with_dk = self.K.ddata is not None and bool(self.opt.get('reuse_lanczos', True))
        if not hasattr(self, '_side_stream'):
            self._side_stream = torch.cuda.Stream()
        # every buffer is allocated on the main stream; the side stream only runs kernels on them
        V = torch.empty((self.rows, B), dtype=torch.float64, device='cuda')
        alpha = torch.empty((m, B), dtype=torch.float64, device='cuda')
        beta = torch.empty((m, B), dtype=torch.float64, device='cuda')
        basis = self._new_basis(m, B) if with_dk else None
        self._workspace(B, side=True)
        cur = torch.cuda.current_stream()
        self._side_stream.wait_stream(cur)
        with torch.cuda.stream(self._side_stream):
            self.probes(first, B, out=V)
            self._lanczos_launch(eta, V, m, basis, alpha, beta, side=True)
        self._prefetched = (float(eta), first, B, with_dk, V, alpha, beta, basis)

    # The Krylov space of K + eta I does not depend on eta and the Lanczos tridiagonal only shifts: T(eta) = T(eta_ref) +
    # (eta - eta_ref) I (the reference builds imate.AffineMatrixFunction(K) for this, mixed_correlation.py:44,141,207,266).
    # One batched Lanczos run per block of probes (and one per right-hand-side block) therefore serves EVERY eta asked of
    # this operator: new quadratures on the host, new solutions x(eta) = ||v|| Q (T + d I)^-1 e_1 by one pass over the kept
    # vectors. A root find or a sweep over eta at fixed rho pays the SpMMs once.
    KRYLOV_CACHE_BYTES = 32 << 30      # kept Lanczos vectors per operator (beyond it blocks are recomputed per eta)

    def _krylov_bytes(self):
        return sum(e['basis'].numel() * 8 for e in self._krylov.values() if e.get('basis') is not None)

    def _probe_krylov(self, eta, first, B, keep_basis):
        """(entry, shift) for the block of probes first .. first+B-1: the cached Lanczos run if there is one (shift =
        eta - eta_ref), else a new run at eta (picked up from the side-stream prefetch when it matches)."""
        torch = dev.torch
        m = int(self.opt['lanczos_degree'])
        key = ('probes', first, B)
        ent = self._krylov.get(key)
        if ent is not None and (ent['basis'] is not None or not keep_basis) and bool(self.opt.get('shift_reuse', True)):
            return ent, float(eta) - ent['eta_ref']
        pre = getattr(self, '_prefetched', None)
        self._prefetched = None
        if pre is not None and pre[:4] == (float(eta), first, B, keep_basis):
            V, alpha, beta, basis = pre[4:]
            torch.cuda.current_stream().wait_stream(self._side_stream)
            a, b = alpha.cpu().numpy(), beta.cpu().numpy()
        else:
            if pre is not None:
                self._side_stream.synchronize()        # an unused prefetch still owns the workspace
            V = self.probes(first, B)
            basis = self._new_basis(m, B) if keep_basis else None
            a, b = self._lanczos(eta, V, m, basis)
        ent = {'eta_ref': float(eta), 'a': a, 'b': b, 'basis': basis, 'V': V if keep_basis else None, 'Wd': None}
        if self._krylov_bytes() + (basis.numel() * 8 if basis is not None else 0) <= self.KRYLOV_CACHE_BYTES:
            self._krylov[key] = ent
        return ent, 0.0

    def _new_basis(self, m, B):
        return dev.torch.empty((m, self.rows, B), dtype=dev.torch.float64, device='cuda')

    def _slq_samples(self, eta, first, B, with_dk=False):
        """Per-probe quadratures [log, 1/x, 1/x^2] * n for probes first .. first+B-1. With ``with_dk`` a fourth column:
        the Hutchinson sample v^T Kn^-1 dK v with Kn^-1 v taken from the SAME Lanczos run (x = ||v|| Q T^-1 e_1, the
        Lanczos form of CG) when its residual beta_k |y_k| meets the CG tolerance, else from a batched CG solve."""
        torch = dev.torch
        m = int(self.opt['lanczos_degree'])
        ent, shift = self._probe_krylov(eta, first, B, with_dk)
        a, b = ent['a'] + shift, ent['b']
        out = numpy.empty((B, 4 if with_dk else 3))
        vnorm = numpy.sqrt(float(self.n))          # Rademacher probes
        quad, tmin, coef, res = lanczos_block_quadrature(a, b, vnorm if with_dk else None)
        if not (tmin.min() > 0):
            raise numpy.linalg.LinAlgError(
                'K + eta*I (eta=%g) is not positive definite: Lanczos found a Ritz value %.3e. The thresholded '
                'Matern matrix is indefinite; use a larger eta (reference: _generate_sparse_correlation.pyx:516-523).'
                % (eta, tmin.min()))
        out[:, :3] = quad * self.n
        resid = float(res.max()) if with_dk else 0.0
        if with_dk:
            V = ent['V'] if ent['V'] is not None else self.probes(first, B)
            if ent['Wd'] is None:
                ent['Wd'] = self.spmm(0.0, V, derivative=True)      # dK v does not depend on eta either
            if resid <= float(self.opt['cg_tol']):
                U = torch.empty_like(V)
                cd = torch.from_numpy(coef).cuda()
                check(lib.gp_block_combine(_p(ent['basis']), self.rows, B, m, _p(cd), _p(U), dev.stream_ptr()),
                      'gp_block_combine')
                self.last_dk_solver = 'lanczos'
            else:
                U = self.solve_dev(eta, V.clone())
                self.last_dk_solver = 'cg'
            out[:, 3] = self.col_dot(U, ent['Wd'])
        return out

    def solve_rhs_block(self, eta, Rop, key, refs=None):
        """S = (K + eta I)^-1 R for the (cached) right-hand-side block of the likelihood, operator space. The first eta
        asked of this operator is solved by batched CG; from the second DISTINCT eta on (from the first with the option
        ``eager_rhs_basis``, which sweeps and the root finder set: they know more eta will follow), one batched Lanczos
        run on R (degree ~ twice the CG iteration count, 48 when unknown) is kept and every eta is served from it (shift
        invariance), falling back to CG for an eta whose Lanczos residual misses the CG tolerance."""
        torch = dev.torch
        eta = float(eta)
        if not bool(self.opt.get('shift_reuse', True)):
            return self.solve_dev(eta, Rop.clone())
        B = Rop.shape[1]
        ent = self._krylov.get(('rhs', key))
        if ent is None:
            seen = self._rhs_etas.setdefault(key, (set(), refs))[0]      # refs keep X, z alive: the ids stay unique
            seen.add(eta)
            if len(seen) < 2 and not bool(self.opt.get('eager_rhs_basis', False)):
                return self.solve_dev(eta, Rop.clone())
            m = int(self.opt.get('solve_degree') or min(128, max(int(self.opt['lanczos_degree']),
                                                                   2 * int(getattr(self, 'last_cg_iterations', 24)))))
            basis = self._new_basis(m, B)
            a, b = self._lanczos(eta, Rop, m, basis)
            norms = numpy.sqrt(numpy.maximum(self.col_dot(Rop, Rop), 0.0))
            ent = {'eta_ref': eta, 'a': a, 'b': b, 'basis': basis, 'norms': norms, 'm': m, 'refs': refs}
            if self._krylov_bytes() + basis.numel() * 8 <= self.KRYLOV_CACHE_BYTES:
                self._krylov[('rhs', key)] = ent
        a, b, m = ent['a'] + (eta - ent['eta_ref']), ent['b'], ent['m']
        live = numpy.nonzero(ent['norms'] > 0.0)[0]                 # zero (padding) columns stay zero
        coef = numpy.zeros((m, B))
        _, tmin, cf, res = lanczos_block_quadrature(a[:, live], b[:, live], ent['norms'][live])
        coef[:, live] = cf
        resid = float(res.max()) if (live.size and tmin.min() > 0) else numpy.inf
        if not (resid <= float(self.opt['cg_tol'])):
            self.last_rhs_solver = 'cg'
            return self.solve_dev(eta, Rop.clone())
        S = torch.empty_like(Rop)
        cd = torch.from_numpy(coef).cuda()
        check(lib.gp_block_combine(_p(ent['basis']), self.rows, B, m, _p(cd), _p(S), dev.stream_ptr()), 'gp_block_combine')
        self.last_rhs_solver = 'lanczos'
        return S

    @staticmethod
    def _chunks(first, count, batch):
        """split [first, first + count) into contiguous power-of-two blocks of width <= batch"""
        out = []
        while count > 0:
            w = 1
            while w * 2 <= min(count, batch):
                w *= 2
            out.append((first, w))
            first += w
            count -= w
        return out

    def _run_estimator(self, sample_fn, ncols, check_cols=None, state=None):
        """imate-style sampling loop: rounds of probes until every estimated quantity (the first ``check_cols`` columns)
        satisfies z * s / sqrt(N) <= max(atol, rtol |mean|) (after min_num_samples) or max_num_samples is reached.
        ``state`` = (samples, first) continues an earlier run (its convergence is tested before sampling more).
        Multi-GPU: each round's probe ids are cut into `world` contiguous slices, one per rank; the running
        (count, sum, sum of squares) are all-reduced, so every rank takes the same stopping decision. Probe ids, not
        ranks, seed the random signs: the union of the samples is the same set for any number of GPUs.
        Returns mean, half_width, N, (samples, first)."""
        o = self.opt
        B = int(o['batch'])
        zc = float(numpy.sqrt(2.0) * scipy.special.erfinv(float(o['confidence_level'])))
        lo, hi = int(o['min_num_samples']), int(o['max_num_samples'])
        rank, world = self.probe_range if self.probe_range is not None else (0, 1)
        cc = ncols if check_cols is None else check_cols
        samples, first = (numpy.empty((0, ncols)), 0) if state is None else state
        atol = o['error_atol'] if o['error_atol'] is not None else 0.0

        failure = []      # a breakdown on THIS rank (indefinite K + eta I: non-positive Ritz value / CG breakdown)

        def reduce_or_raise():
            """all-reduce of the running sums together with a 'some rank failed' flag: a local breakdown must not leave
            the other ranks blocked in the collective - every rank raises after it."""
            N, mean, sd, failed = self._reduce(samples, bool(failure))
            if failed:
                raise failure[0] if failure else numpy.linalg.LinAlgError(
                    'K + eta*I is not positive definite (breakdown reported by another rank of the probe split).')
            return N, mean, sd

        def converged():
            N, mean, sd = reduce_or_raise()
            if N < lo:
                return False
            half = zc * sd / numpy.sqrt(N)
            return bool(numpy.all(half[:cc] <= numpy.maximum(atol, o['error_rtol'] * numpy.abs(mean[:cc]))))

        done = (first > 0) and converged()
        while first < hi and not done:
            nb = min(B, hi - first)          # the round size does not depend on the rank count: same samples, same stopping decision
            per = (nb + world - 1) // world
            my0 = first + rank * per
            my1 = min(first + nb, my0 + per)
            for (f, w) in self._chunks(my0, max(0, my1 - my0), B):
                if failure:
                    break
                try:
                    samples = numpy.vstack([samples, sample_fn(f, w)])
                except numpy.linalg.LinAlgError as exc:
                    if world == 1:
                        raise
                    failure.append(exc)
            first += nb
            done = converged()
        N, mean, sd = reduce_or_raise()
        half = zc * sd / numpy.sqrt(max(N, 1))
        return mean, half, int(N), (samples, first)

    def _reduce(self, samples, failed=False):
        """(N, mean, unbiased std, any rank failed) over all ranks: all-reduce of (count, failure flag, sum, sum of
        squares) per quantity."""
        cnt = numpy.array([samples.shape[0], 1.0 if failed else 0.0], dtype=float)
        s1 = samples.sum(axis=0)
        s2 = (samples ** 2).sum(axis=0)
        if self.probe_range is not None and self.probe_range[1] > 1:
            from ._distributed import allreduce_sum
            packed = allreduce_sum(numpy.concatenate([cnt, s1, s2]))
            k = s1.size
            cnt, s1, s2 = packed[:2], packed[2:2 + k], packed[2 + k:]
        N = cnt[0]
        mean = s1 / max(N, 1)
        var = numpy.maximum(s2 - N * mean ** 2, 0.0) / max(N - 1, 1)
        return N, mean, numpy.sqrt(var), bool(cnt[1] > 0)

    def _slq(self, eta):
        """SLQ estimates [logdet, tr Kn^-1, tr Kn^-2] at eta (cached). When the matrix carries dK/drho the same Lanczos
        runs also yield the Hutchinson samples of tr(Kn^-1 dK) (kept for traceinv_dK; they do not influence the
        stopping decision here)."""
        key = float(eta)
        if key not in self._slq_cache:
            with_dk = self.K.ddata is not None and bool(self.opt.get('reuse_lanczos', True))
            mean, half, N, state = self._run_estimator(
                lambda first, width: self._slq_samples(eta, first, width, with_dk), 4 if with_dk else 3, check_cols=3)
            self._slq_cache = {key: (mean[:3], half[:3], N)}
            self._dk_state = {key: (state[0][:, 3:4].copy(), state[1])} if with_dk else {}
        mean, half, N = self._slq_cache[key]
        self.last_info = {'num_samples': N, 'half_width': half, 'confidence_level': self.opt['confidence_level'],
                          'lanczos_degree': self.opt['lanczos_degree']}
        return mean

    def logdet(self, eta):
        """SLQ estimate of logdet(K + eta I) (also for method 'hutchinson')."""
        return float(self._slq(eta)[0])

    def traceinv(self, eta, exponent=1):
        if exponent not in (1, 2):
            raise ValueError('traceinv on the sparse engine supports exponent 1 and 2.')
        if self.method == 'slq' or exponent == 2:
            return float(self._slq(eta)[exponent])

        def fn(first, width):   # hutchinson: v^T Kn^-1 v with CG solves
            V = self.probes(first, width)
            U = self.solve_dev(eta, V.clone())
            return self.col_dot(U, V).reshape(-1, 1)
        mean, half, N, _ = self._run_estimator(fn, 1)
        self.last_info = {'num_samples': N, 'half_width': half}
        return float(mean[0])

    def traceinv_dK(self, eta):
        """Hutchinson estimate of tr((K + eta I)^-1 dK/d rho): mean over probes of (Kn^-1 v)^T (dK v). Samples that the
        SLQ run at this eta already produced (same probe ids, Kn^-1 v from its Lanczos basis) are used first; further
        rounds, if its own stopping rule asks for them, go through the same kept-Lanczos path (CG when the Lanczos
        residual misses the tolerance or with reuse_lanczos=False)."""
        if self.K.ddata is None:
            raise ValueError('needs a DeviceCSR generated with with_derivative=True')

        def fn(first, width):
            if bool(self.opt.get('reuse_lanczos', True)):
                # through the (kept, shift-invariant) Lanczos run of this probe block: reused for every later eta
                return self._slq_samples(eta, first, width, True)[:, 3:4]
            V = self.probes(first, width)
            U = self.solve_dev(eta, V.clone())
            Wd = self.spmm(0.0, V, derivative=True)
            return self.col_dot(U, Wd).reshape(-1, 1)
        state = getattr(self, '_dk_state', {}).get(float(eta))
        mean, half, N, _ = self._run_estimator(fn, 1, state=state)
        self.last_info = {'num_samples': N, 'half_width': half}
        return float(mean[0])

    # ---- CG ------------------------------------------------------------------------------------------------------
    def col_dot(self, X, Y):
        torch = dev.torch
        B = X.shape[1]
        out = torch.empty(32, dtype=torch.float64, device='cuda')
        check(lib.gp_col_dot(_p(X), _p(Y), self.n, B, _p(out), _p(self._workspace(B)), dev.stream_ptr()), 'gp_col_dot')
        return out[:B].cpu().numpy()

    def solve_dev(self, eta, R_dev):
        """(K + eta I)^-1 R for a device block (n x B, B a power of two <= 32); R_dev is overwritten."""
        torch = dev.torch
        B = R_dev.shape[1]
        X = torch.empty_like(R_dev)
        it = ctypes.c_int64()
        if self.blocked is not None:
            bptr, bidx, bvals, _ = self.blocked
            rc = lib.gp_bcsr_cg_solve(self.R, _p(bptr), _p(bidx), _p(bvals), self.n, float(eta), _p(R_dev), _p(X), B,
                                      float(self.opt['cg_tol']), int(self.opt['cg_maxiter']), ctypes.byref(it),
                                      _p(self._workspace(B)), dev.stream_ptr())
        else:
            K = self.K
            rc = lib.gp_cg_solve(_p(K.indptr), _p(K.indices), _p(K.data), self.n, float(eta), _p(R_dev), _p(X), B,
                                 float(self.opt['cg_tol']), int(self.opt['cg_maxiter']), ctypes.byref(it),
                                 _p(self._workspace(B)), dev.stream_ptr())
        check(rc, 'gp_cg_solve')
        self.last_cg_iterations = it.value
        if rc == 2:
            raise numpy.linalg.LinAlgError(
                'K + eta*I (eta=%g) is not positive definite (CG met p^T A p <= 0). The thresholded Matern matrix is '
                'indefinite; use a larger eta (reference: _generate_sparse_correlation.pyx:516-523).' % eta)
        if rc == 1:
            raise numpy.linalg.LinAlgError('CG did not converge in %d iterations (eta=%g)' % (it.value, eta))
        return X

    def _max_block(self):
        """widest column block handed to the SpMM: the 16 x 1 row-blocked kernel is tuned for B <= 16"""
        return 16 if self.blocked is not None and self.R == 16 else 32

    def solve(self, eta, Y):
        """mixed_correlation.py:280-299 for sparse K: CG per column with rtol 1e-6 (batched on the device)."""
        torch = dev.torch
        Y = numpy.asarray(Y, dtype=numpy.float64)
        vec = (Y.ndim == 1)
        Y2 = Y.reshape(self.n, -1)
        out = numpy.empty_like(Y2)
        k = Y2.shape[1]
        W = self._max_block()
        for c0 in range(0, k, W):
            blk = Y2[:, c0:c0 + W]
            B = 1
            while B < blk.shape[1]:
                B *= 2
            R = torch.zeros((self.n, B), dtype=torch.float64, device='cuda')
            R[:, :blk.shape[1]].copy_(torch.from_numpy(numpy.ascontiguousarray(blk)))
            X = self.from_op(self.solve_dev(eta, self.to_op(R)))
            out[:, c0:c0 + blk.shape[1]] = X[:, :blk.shape[1]].cpu().numpy()
        return out[:, 0] if vec else out

    def matmul(self, X):
        torch = dev.torch
        X = numpy.asarray(X, dtype=numpy.float64)
        vec = (X.ndim == 1)
        X2 = X.reshape(self.n, -1)
        k = X2.shape[1]
        B = 1
        while B < k:
            B *= 2
        W = self._max_block()
        if B > W:
            return numpy.hstack([self.matmul(X2[:, c:c + W]) for c in range(0, k, W)])
        Xd = torch.zeros((self.n, B), dtype=torch.float64, device='cuda')
        Xd[:, :k].copy_(torch.from_numpy(numpy.ascontiguousarray(X2)))
        res = self.from_op(self.spmm(0.0, self.to_op(Xd)))[:, :k].cpu().numpy()
        return res[:, 0] if vec else res

    # ---- fused likelihood ingredients -------------------------------------------------------------------------
    def gram(self, X_dev, Y_dev):
        """X^T Y (B x B, host) of two n x B device blocks."""
        torch = dev.torch
        B = X_dev.shape[1]
        if '_gram' not in self._ws:
            self._ws['_gram'] = torch.empty(lib.gp_gram_workspace_bytes(16) // 8, dtype=torch.float64, device='cuda')
        out = torch.empty(B * B, dtype=torch.float64, device='cuda')
        check(lib.gp_gram_skinny(_p(X_dev), _p(Y_dev), self.n, B, _p(out), _p(self._ws['_gram']), dev.stream_ptr()),
              'gp_gram_skinny')
        return out.cpu().numpy().reshape(B, B)

    def _rhs_block(self, X, z, key=None):
        """[X z] as an operator-space device block, zero padded to a power-of-two width. The device copy in the original
        row order is cached per (X, z) objects across engines (an optimiser builds a new operator for every rho but
        keeps X and z); the operator-space permutation of it is cached per engine."""
        torch = dev.torch
        if key is None:
            key = (dev.host_key(X), dev.host_key(z))
        if getattr(self, '_rhs_cache', None) is not None and self._rhs_cache[0] == key:
            return self._rhs_cache[1]
        Rd = None
        for ent in _RHS_DEVICE_CACHE:
            if ent[0] == key and ent[1].shape[0] == self.n:
                Rd = ent[1]
        if Rd is None:
            X2 = numpy.asarray(X, dtype=float)
            p = X2.shape[1] + 1
            B = 1
            while B < p:
                B *= 2
            if B > 16:
                raise ValueError('the sparse likelihood evaluation supports at most 15 basis functions.')
            R = numpy.zeros((self.n, B))
            R[:, :p - 1] = X2
            R[:, p - 1] = numpy.asarray(z, dtype=float)
            Rd = torch.from_numpy(R).cuda()
            _RHS_DEVICE_CACHE.append((key, Rd, X, z))       # X, z kept alive so the ids stay unique
            del _RHS_DEVICE_CACHE[:-2]
        Rop = self.to_op(Rd)
        self._rhs_cache = (key, Rop, X, z)
        return Rop

    def fused(self, eta, X, z, traceinv=True, drho=True, cubic=False):
        """Everything log-likelihood + gradient need at one eta, in the layout of the dense evaluator's out[]
        (csrc/gp_loglik.cu): [logdet Kn, tr Kn^-1, tr Kn^-2, tr(Kn^-1 dK/drho), info, -, -, -, G, H, Q] with R = [X z],
        S = Kn^-1 R (batched CG, the reference's tol 1e-6), G = R^T S, H = S^T S, Q = S^T dK S; the traces are the
        stochastic estimates (SLQ / Hutchinson) of this engine. Reference formulas: _direct_likelihood.py:113-150,
        _profile_likelihood.py:104-130."""
        n, m = X.shape
        p = m + 1
        key = (dev.host_key(X), dev.host_key(z))          # ONE content fingerprint per evaluation
        Rd = self._rhs_block(X, z, key)
        if bool(self.opt.get('overlap', True)) and (self.method == 'slq' or drho):
            self.prefetch_slq(eta)
        S = self.solve_rhs_block(eta, Rd, key, refs=(X, z))
        out = numpy.zeros(8 + 4 * p * p)
        out[8:8 + p * p] = self.gram(Rd, S)[:p, :p].ravel()
        out[8 + p * p:8 + 2 * p * p] = self.gram(S, S)[:p, :p].ravel()
        if drho:
            if self.K.ddata is None:
                raise ValueError('d/d(correlation_scale) needs a DeviceCSR generated with with_derivative=True')
            D = self.spmm(0.0, S, derivative=True)
            out[8 + 2 * p * p:8 + 3 * p * p] = self.gram(S, D)[:p, :p].ravel()
        if cubic:
            # third moments T3 = S^T Kn^-1 S (Hessian / second eta-derivative): one more batched CG on S
            V2 = self.solve_dev(eta, S.clone())
            out[8 + 3 * p * p:8 + 4 * p * p] = self.gram(S, V2)[:p, :p].ravel()
        out[0] = self.logdet(eta)
        if traceinv or drho:
            out[1] = self.traceinv(eta)
            out[2] = self.traceinv(eta, exponent=2)
        if drho:
            out[3] = self.traceinv_dK(eta)
        return out

    def trace_K(self):
        """(tr K, tr K^2 = sum of squared entries) of the stored symmetric matrix (rare path: host reduction)."""
        Kh = self.K.to_scipy()
        return float(Kh.diagonal().sum()), float((Kh.data ** 2).sum())
