"""Developer timing of the dense Matern generator (csrc/gp_matern.cu) at the headline size: K only and K + dK/drho, CUDA
events; the ncu capture of profiles/r02_matern_dense_ncu_summary.md runs this script."""
import ctypes, json, os, sys
import numpy
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'gaussian-process-param-estimation_b200'))
import torch
from gaussian_proc import _device as dev
lib = dev.lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
nu = float(sys.argv[2]) if len(sys.argv) > 2 else 2.5
numpy.random.seed(0)
pts = torch.from_numpy(numpy.random.rand(n, 2)).cuda()
npad = dev.padded_size(n)
K = torch.empty((npad, npad), dtype=torch.float64, device='cuda')
dK = torch.empty((npad, npad), dtype=torch.float64, device='cuda')
scale = dev.host_f64(numpy.array([0.1, 0.1]))
P = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
out = {'n': n, 'npad': npad, 'nu': nu}
for name, dk in (('K', None), ('K_and_dK', dK)):
    best = 1e30
    for rep in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        rc = lib.gp_matern_dense(P(pts), n, 2, dev.host_ptr(scale), nu, P(K), npad, P(dk) if dk is not None else None, dev.stream_ptr())
        b.record(); torch.cuda.synchronize()
        assert rc == 0
        best = min(best, a.elapsed_time(b))
    nbytes = 8.0 * npad * npad * (2 if dk is not None else 1)
    out[name] = {'ms': best, 'GBps_written': nbytes / best * 1e-6}
print(json.dumps(out))
