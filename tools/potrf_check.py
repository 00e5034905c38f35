"""Developer check: gp_potrf_f64 / trtri / lauum with the TMA GEMM against the cp.async GEMM, element by element."""
import ctypes, os, sys
import numpy
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'gaussian-process-param-estimation_b200'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch
import bench
from gaussian_proc import _device as dev, generate_correlation
lib = dev.lib
P = lambda t: ctypes.c_void_p(t.data_ptr())
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
pts, z, X = bench.make_inputs(n)
K = generate_correlation(pts, 0.104, 2.5, device=True)
npad = K.npad
f64 = torch.float64
pws = torch.empty(lib.gp_potrf_workspace_bytes(npad) // 8, dtype=f64, device='cuda')
tws = torch.empty(lib.gp_potri_workspace_bytes(npad) // 8 + 8, dtype=f64, device='cuda')
info = torch.zeros(1, dtype=torch.int32, device='cuda')

def run(impl, stage):
    lib.gp_gemm_set_impl(impl)
    A = torch.empty((npad, npad), dtype=f64, device='cuda')
    s = dev.stream_ptr()
    lib.gp_shift_copy(P(K.data), n, npad, 0.01, P(A), s)
    lib.gp_potrf_f64(P(A), n, npad, P(info), P(pws), s)
    res = [torch.tril(A).clone(), int(info.item())]
    if stage >= 1:
        W = torch.zeros((npad, npad), dtype=f64, device='cuda')
        lib.gp_trtri_f64(P(A), P(W), npad, P(pws), P(tws), s)
        res.append(W)
    torch.cuda.synchronize()
    return res

def where(a, b):
    d = (a - b).abs()
    m = float(d.max())
    idx = int(d.argmax())
    return m, divmod(idx, a.shape[1])

ref = run(0, 1)
print('cpasync info', ref[1])
for rep in range(3):
    got = run(1, 1)
    m, (r, c) = where(got[0], ref[0])
    bad = ((got[0] - ref[0]).abs() > 1e-6).nonzero()
    print('rep', rep, 'tma info', got[1], 'potrf max|diff|', m, 'at', (r, c), 'n bad', bad.shape[0],
          'first bad', bad[0].tolist() if bad.shape[0] else None, 'bad row range',
          (int(bad[:, 0].min()), int(bad[:, 0].max()), int(bad[:, 1].min()), int(bad[:, 1].max())) if bad.shape[0] else None)
    m2, rc2 = where(got[2], ref[2])
    print('        trtri max|diff|', m2, 'at', rc2)
    del got
