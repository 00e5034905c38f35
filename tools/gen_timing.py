"""Developer timing: host/device breakdown of the sparse generator + row-blocked build at n = 2^20."""
import ctypes, os, sys, time
import numpy
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'gaussian-process-param-estimation_b200'))
import torch
from gaussian_proc import _device as dev
from gaussian_proc import _sparse as S
lib = dev.lib
_p = S._p

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2 ** 20
numpy.random.seed(0)
points = numpy.random.rand(n, 2)
scale = dev.host_f64(numpy.array([0.005, 0.005]))
nu, density = 0.5, 1e-3
d = 2

def run():
    T = {}
    def tick(name, t0):
        torch.cuda.synchronize()
        T[name] = (time.perf_counter() - t0) * 1e3
        return time.perf_counter()
    t = time.perf_counter()
    tau = S.estimate_kernel_threshold(n, d, density, scale, nu)
    dpts = torch.from_numpy(points).cuda()
    ws = torch.empty(lib.gp_sparse_workspace_bytes(n, d) // 8 + 8, dtype=torch.float64, device='cuda')
    indptr = torch.empty(n + 1, dtype=torch.int32, device='cuda')
    nnz = ctypes.c_int64()
    s = dev.stream_ptr()
    t = tick('setup+h2d', t)
    lib.gp_matern_sparse_count(_p(dpts), dev.host_ptr(points), n, d, dev.host_ptr(scale), float(nu), tau, _p(ws), _p(indptr), ctypes.byref(nnz), s)
    t = tick('count', t)
    lo, hi = dev.host_f64(dpts.amin(dim=0).cpu().numpy()), dev.host_f64(dpts.amax(dim=0).cpu().numpy())
    keys = torch.empty(n, dtype=torch.int64, device='cuda')
    lib.gp_spatial_keys(_p(dpts), n, d, dev.host_ptr(lo), dev.host_ptr(hi), _p(keys), s)
    order = torch.sort(keys, stable=True)[1].to(torch.int32)
    t = tick('order', t)
    indices = torch.empty(nnz.value, dtype=torch.int32, device='cuda')
    data = torch.empty(nnz.value, dtype=torch.float64, device='cuda')
    ddata = torch.empty(nnz.value, dtype=torch.float64, device='cuda')
    t = tick('alloc', t)
    lib.gp_matern_sparse_fill(_p(dpts), dev.host_ptr(points), n, d, dev.host_ptr(scale), float(nu), tau, _p(ws), _p(indptr), _p(indices), _p(data), _p(ddata), 0, s)
    t = tick('fill', t)
    K = S.DeviceCSR(n, indptr, indices, data, ddata, kernel_threshold=tau, order=order, sorted_rows=False)
    e = S.SparseEngine(K, 'slq', {})
    t = tick('blocked_build', t)
    return T

run()
print(run())
print(run())
