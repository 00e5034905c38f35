"""Developer check (GPU): dense generator, potrf, potrs, potri against torch/numpy, plus phase timings.
Run on the GPU box:  python tools/gpu_check_dense.py [n ...]
"""
import ctypes
import json
import os
import sys
import time

import numpy

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..',
                                'gaussian-process-param-estimation_b200'))
import torch  # noqa: E402
from gaussian_proc import _device as dev  # noqa: E402

lib = dev.lib


def P(t):
    return ctypes.c_void_p(t.data_ptr())


def matern_np(pts, rho, nu):
    d = numpy.sqrt(((pts[:, None, :] - pts[None, :, :]) ** 2).sum(-1)) / rho
    if nu == 0.5:
        return numpy.exp(-d)
    if nu == 1.5:
        return (1 + numpy.sqrt(3) * d) * numpy.exp(-numpy.sqrt(3) * d)
    if nu == 2.5:
        return (1 + numpy.sqrt(5) * d + 5.0 / 3.0 * d ** 2) * numpy.exp(-numpy.sqrt(5) * d)
    return numpy.exp(-0.5 * d ** 2)


def timed(fn, reps=1):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def run(n, nu=2.5, rho=0.1, eta=0.1, verify=True):
    numpy.random.seed(0)
    pts = numpy.random.rand(n, 2)
    npad = dev.padded_size(n)
    s = dev.stream_ptr()
    dpts = torch.from_numpy(pts).cuda()
    K = torch.empty((npad, npad), dtype=torch.float64, device='cuda')
    dK = torch.empty((npad, npad), dtype=torch.float64, device='cuda')
    A = torch.empty((npad, npad), dtype=torch.float64, device='cuda')
    W = torch.zeros((npad, npad), dtype=torch.float64, device='cuda')
    scale = numpy.array([rho, rho])
    info = torch.zeros(1, dtype=torch.int32, device='cuda')
    ws_potrf = torch.empty(lib.gp_potrf_workspace_bytes(npad) // 8, dtype=torch.float64, device='cuda')
    ws_potri = torch.empty(lib.gp_potri_workspace_bytes(npad) // 8 + 1, dtype=torch.float64, device='cuda')
    out = {'n': n, 'npad': npad, 'nu': nu}

    def gen():
        rc = lib.gp_matern_dense(P(dpts), n, 2, dev.host_ptr(scale), nu, P(K), npad, P(dK), s)
        assert rc == 0, rc

    def gen_nodk():
        rc = lib.gp_matern_dense(P(dpts), n, 2, dev.host_ptr(scale), nu, P(K), npad, None, s)
        assert rc == 0, rc
    gen()
    out['t_matern_dK_ms'] = timed(gen, 3)
    out['t_matern_ms'] = timed(gen_nodk, 3)
    out['matern_GBs'] = 8.0 * npad * npad / out['t_matern_ms'] * 1e-6

    def shift():
        assert lib.gp_shift_copy(P(K), n, npad, eta, P(A), s) == 0

    def potrf():
        rc = lib.gp_potrf_f64(P(A), n, npad, P(info), P(ws_potrf), s)
        assert rc == 0, rc
    shift(); potrf()
    shift()
    out['t_shift_ms'] = timed(shift)
    out['t_potrf_ms'] = timed(potrf)
    out['potrf_tflops'] = (npad ** 3 / 3.0) / out['t_potrf_ms'] * 1e-9
    out['info'] = int(info.item())

    if verify:
        Kh = K[:n, :n].cpu().numpy()
        Kref = matern_np(pts, rho, nu)
        out['matern_max_abs_err'] = float(numpy.abs(Kh - Kref).max())
        out['matern_sym'] = bool((Kh == Kh.T).all())
        out['matern_diag1'] = bool((numpy.diag(Kh) == 1.0).all())
        # dK by finite differences
        h = 1e-6
        dref = (matern_np(pts, rho + h, nu) - matern_np(pts, rho - h, nu)) / (2 * h)
        out['dK_max_abs_err_fd'] = float(numpy.abs(dK[:n, :n].cpu().numpy() - dref).max())
        if npad > n:
            pad = K[n:, :].cpu().numpy()
            eye = numpy.zeros_like(pad)
            eye[numpy.arange(npad - n), n + numpy.arange(npad - n)] = 1.0
            out['pad_identity'] = bool((pad == eye).all())
        L = torch.tril(A)[:n, :n]
        Kn = K[:n, :n] + eta * torch.eye(n, dtype=torch.float64, device='cuda')
        Lref = torch.linalg.cholesky(Kn)
        out['potrf_rel_err'] = float((L - Lref).abs().max() / Lref.abs().max())
        out['potrf_resid'] = float((L @ L.T - Kn).abs().max())

    # logdet
    ld = torch.zeros(1, dtype=torch.float64, device='cuda')
    assert lib.gp_logdet_from_chol(P(A), n, npad, P(ld), s) == 0
    out['logdet'] = float(ld.item())
    if verify:
        out['logdet_ref'] = float(torch.linalg.slogdet(Kn)[1].item())

    # potrs
    nrhs = 7
    B = torch.zeros((npad, nrhs), dtype=torch.float64, device='cuda')
    B[:n] = torch.from_numpy(numpy.random.rand(n, nrhs)).cuda()
    B0 = B.clone()

    def potrs():
        assert lib.gp_potrs_f64(P(A), npad, P(ws_potrf), P(B), nrhs, nrhs, s) == 0
    potrs()
    if verify:
        Xref = torch.cholesky_solve(B0[:n], Lref)
        out['potrs_rel_err'] = float((B[:n] - Xref).abs().max() / Xref.abs().max())
    B.copy_(B0)
    out['t_potrs_ms'] = timed(potrs)

    # potri (keeps a copy of L since lauum overwrites A)
    Lsave = A.clone()

    def trtri():
        assert lib.gp_trtri_f64(P(A), P(W), npad, P(ws_potrf), P(ws_potri), s) == 0

    def lauum():
        assert lib.gp_lauum_f64(P(W), P(A), npad, s) == 0
    out['t_trtri_ms'] = timed(trtri)
    out['t_lauum_ms'] = timed(lauum)
    out['trtri_tflops'] = (npad ** 3 / 3.0) / out['t_trtri_ms'] * 1e-9
    out['lauum_tflops'] = (npad ** 3 / 3.0) / out['t_lauum_ms'] * 1e-9
    if verify:
        Winv = torch.tril(W)[:n, :n]
        Lt = torch.tril(Lsave)[:n, :n]
        out['trtri_resid'] = float((Winv @ Lt - torch.eye(n, dtype=torch.float64, device='cuda')).abs().max())
        Ainv = torch.tril(A)[:n, :n]
        Aref = torch.cholesky_inverse(Lref)
        out['potri_rel_err'] = float((Ainv - torch.tril(Aref)).abs().max() / Aref.abs().max())
    # torch / cuSOLVER comparison
    Kn2 = (K + eta * torch.eye(npad, dtype=torch.float64, device='cuda'))
    torch.linalg.cholesky(Kn2)
    out['t_torch_cholesky_ms'] = timed(lambda: torch.linalg.cholesky(Kn2))
    out['t_total_eval_ms'] = out['t_shift_ms'] + out['t_potrf_ms'] + out['t_potrs_ms'] + out['t_trtri_ms'] + out['t_lauum_ms']
    print(json.dumps(out))


def time_diag_kernel():
    """potrf of a single 128x128 block = one chol_diag_block_kernel launch (+ a 4-byte memset)"""
    numpy.random.seed(1)
    pts = numpy.random.rand(128, 2)
    K = torch.from_numpy(matern_np(pts, 0.1, 2.5) + 0.1 * numpy.eye(128)).cuda()
    A = K.clone()
    info = torch.zeros(1, dtype=torch.int32, device='cuda')
    ws = torch.empty(lib.gp_potrf_workspace_bytes(128) // 8, dtype=torch.float64, device='cuda')
    s = dev.stream_ptr()

    def f():
        lib.gp_potrf_f64(P(A), 128, 128, P(info), P(ws), s)
    for _ in range(3):
        A.copy_(K); f()
    ts = []
    for _ in range(20):
        A.copy_(K)
        ts.append(timed(f))
    L = torch.tril(A)
    print(json.dumps({'diag_block_kernel_us_median': float(numpy.median(ts)) * 1e3,
                      'resid': float((L @ L.T - K).abs().max())}))


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'diag':
        time_diag_kernel()
        sys.exit(0)
    sizes = [int(a) for a in sys.argv[1:]] or [1000]
    for n in sizes:
        run(n, verify=(n <= 6000))
