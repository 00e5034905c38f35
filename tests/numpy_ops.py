"""TEST INFRASTRUCTURE: NumPy implementation of the `ops` interface of gaussian_proc/_blockcyclic.py so that the
distributed block-cyclic algorithm (indexing, broadcasts, reductions) can be exercised on CPU with the gloo backend.
Storage = torch CPU float64 tensors (gloo collectives work on them)."""

import numpy
import torch


class NumpyOps(object):
    def empty(self, shape):
        return torch.zeros(shape, dtype=torch.float64)

    zeros = empty

    def from_host(self, a, dtype=None):
        return torch.from_numpy(numpy.ascontiguousarray(a).copy())

    def to_host(self, t):
        return t.numpy().copy()

    def generate(self, prow, pcol, rg, cg, n, scale, nu, eta, out):
        pr, pc = prow.numpy(), pcol.numpy()
        x = numpy.sqrt((((pr[:, None, :] - pc[None, :, :]) / scale) ** 2).sum(-1))
        if nu == 0.5:
            K = numpy.exp(-x)
        elif nu == 1.5:
            K = (1 + numpy.sqrt(3) * x) * numpy.exp(-numpy.sqrt(3) * x)
        else:
            K = (1 + numpy.sqrt(5) * x + 5.0 / 3.0 * x ** 2) * numpy.exp(-numpy.sqrt(5) * x)
        gi, gj = rg.numpy()[:, None], cg.numpy()[None, :]
        K = numpy.where(gi == gj, 1.0 + eta, K)
        pad = (gi >= n) | (gj >= n)
        K = numpy.where(pad, (gi == gj).astype(float), K)
        out.copy_(torch.from_numpy(K))

    def potrf_inv(self, D, nvalid):
        try:
            L = numpy.linalg.cholesky(D.numpy())
        except numpy.linalg.LinAlgError:
            return torch.zeros_like(D), 1
        D.copy_(torch.from_numpy(L))
        return torch.from_numpy(numpy.linalg.inv(L)), 0

    def gemm_nt(self, C, A, B, alpha, beta):
        C.copy_(beta * C + alpha * (A @ B.T))

    def logdet_chol(self, D):
        return float(2.0 * numpy.sum(numpy.log(numpy.diag(D.numpy()))))

    def matmul(self, A, X):
        return A @ X

    def matmul_t(self, A, X):
        return A.T @ X
