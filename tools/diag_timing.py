import ctypes, sys, os, numpy, torch
lib = ctypes.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'bin', 'libgpgp_diagtiming.so'))
lib.gp_potrf_workspace_bytes.restype = ctypes.c_int64
P = lambda t: ctypes.c_void_p(t.data_ptr())
numpy.random.seed(1)
pts = numpy.random.rand(128, 2)
d = numpy.sqrt(((pts[:, None] - pts[None]) ** 2).sum(-1)) / 0.1
K = torch.from_numpy((1 + numpy.sqrt(5) * d + 5 / 3 * d ** 2) * numpy.exp(-numpy.sqrt(5) * d) + 0.1 * numpy.eye(128)).cuda()
info = torch.zeros(1, dtype=torch.int32, device='cuda')
ws = torch.empty(128 * 128, dtype=torch.float64, device='cuda')
for _ in range(5):
    A = K.clone()
    lib.gp_potrf_f64(P(A), ctypes.c_int64(128), ctypes.c_int64(128), P(info), P(ws), None)
    torch.cuda.synchronize()
out = (ctypes.c_longlong * 32)()
lib.gp_diag_timing_read(out)
t = list(out)
names = {0: 'start', 1: 'loaded'}
print('load', t[1] - t[0])
for J in range(4):
    b = 2 + 4 * J
    print('J=%d factor32 %d' % (J, t[b + 1] - t[b]), ('solve %d update %d' % (t[b + 2] - t[b + 1], (t[b + 4] if J < 3 else 0) - t[b + 2])) if J < 3 else '')
print('inv diag blocks', t[19] - t[18], 'offdiag', t[20] - t[19], 'store', t[21] - t[20], 'total', t[21] - t[0])
