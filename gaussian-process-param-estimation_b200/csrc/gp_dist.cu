// Helper kernels of the distributed (multi-GPU) dense path (gaussian_proc/_blockcyclic.py): skinny products with the
// rectangular row slab of inv(L) a rank owns, and the weighted Frobenius inner product that turns a block column of
// inv(L)^T inv(L) and the same block column of dK/drho into their share of tr(Kn^-1 dK). All HBM-bound, one pass over
// the big operand, fixed-order two-stage reductions (bit-reproducible).
#include "../../include/gpgp.h"
#include "gp_common.cuh"
#include "gp_internal.h"

namespace gp {

constexpr int RP = 16;        // max skinny width
constexpr int RCH = 2048;     // rows per partial of the transposed product

// Y[r][c] = sum_k X[r][k] R[k][c]: one warp per row, lanes stride over k (coalesced), p <= 16 accumulators per lane
__global__ void __launch_bounds__(256)
rect_apply_kernel(const double* __restrict__ X, int64_t M, int64_t N, int64_t ldx, const double* __restrict__ R, int p,
                  int64_t ldr, double* __restrict__ Y, int64_t ldy, double alpha, double beta) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t r = (int64_t)blockIdx.x * 8 + warp; r < M; r += (int64_t)gridDim.x * 8) {
        const double* xr = X + r * ldx;
        double acc[RP];
#pragma unroll
        for (int c = 0; c < RP; ++c) acc[c] = 0.0;
        for (int64_t k = lane; k < N; k += 32) {
            const double x = xr[k];
            const double* rr = R + k * ldr;
#pragma unroll
            for (int c = 0; c < RP; ++c)
                if (c < p) acc[c] += x * rr[c];
        }
#pragma unroll
        for (int c = 0; c < RP; ++c)
            if (c < p) {
                double v = warp_sum(acc[c]);
                if (lane == 0) Y[r * ldy + c] = (beta != 0.0) ? alpha * v + beta * Y[r * ldy + c] : alpha * v;
            }
    }
}

// partial[chunk][k][c] = sum_{r in chunk} X[r][k] Y[r][c]: thread per column k (coalesced over k), rows of the chunk in order
__global__ void __launch_bounds__(256)
rect_apply_t_partial_kernel(const double* __restrict__ X, int64_t M, int64_t N, int64_t ldx, const double* __restrict__ Y,
                            int p, int64_t ldy, double* __restrict__ partial) {
    __shared__ double ys[64 * RP];
    const int64_t k = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const int64_t r0 = (int64_t)blockIdx.y * RCH, r1 = min(M, r0 + RCH);
    double acc[RP];
#pragma unroll
    for (int c = 0; c < RP; ++c) acc[c] = 0.0;
    for (int64_t base = r0; base < r1; base += 64) {
        const int cnt = (int)min((int64_t)64, r1 - base);
        __syncthreads();
        for (int idx = threadIdx.x; idx < cnt * p; idx += 256) ys[(idx / p) * RP + idx % p] = Y[(base + idx / p) * ldy + idx % p];
        __syncthreads();
        if (k < N)
            for (int r = 0; r < cnt; ++r) {
                const double x = X[(base + r) * ldx + k];
#pragma unroll
                for (int c = 0; c < RP; ++c)
                    if (c < p) acc[c] += x * ys[r * RP + c];
            }
    }
    if (k < N)
        for (int c = 0; c < p; ++c) partial[((int64_t)blockIdx.y * N + k) * p + c] = acc[c];
}

__global__ void rect_apply_t_reduce_kernel(const double* __restrict__ partial, int64_t total, int nchunks, double* __restrict__ S,
                                           int p, int64_t lds, double alpha, double beta) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    double s = 0.0;
    for (int i = 0; i < nchunks; ++i) s += partial[(int64_t)i * total + t];
    double* dst = S + (t / p) * lds + t % p;
    *dst = (beta != 0.0) ? alpha * s + beta * *dst : alpha * s;
}

// partial[b] = sum over this block's rows of w(r) * sum_c A[r][c] B[r][c]; w = 1 for r < rows_w1, w_rest after
__global__ void __launch_bounds__(256)
pair_dot_partial_kernel(const double* __restrict__ A, int64_t lda, const double* __restrict__ B, int64_t ldb, int64_t rows,
                        int64_t cols, int64_t rows_w1, double w_rest, double* __restrict__ partial) {
    __shared__ double red[32];
    double acc = 0.0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t r = (int64_t)blockIdx.x * 8 + warp; r < rows; r += (int64_t)gridDim.x * 8) {
        const double* a = A + r * lda;
        const double* b = B + r * ldb;
        double s = 0.0;
        for (int64_t c = lane; c < cols; c += 32) s += a[c] * b[c];
        acc += (r < rows_w1) ? s : w_rest * s;
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}

__global__ void pair_dot_final_kernel(const double* __restrict__ partial, int n, double* out) {
    __shared__ double red[32];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += partial[i];      // fixed assignment -> fixed order
    s = block_sum(s, red);
    if (threadIdx.x == 0) out[0] += s;
}

constexpr int PAIR_BLOCKS = 592;

}  // namespace gp

using namespace gp;

extern "C" {

int64_t gp_rect_workspace_bytes(int64_t M, int64_t N, int64_t p) {
    int64_t chunks = (M + RCH - 1) / RCH;
    int64_t a = chunks * N * p * (int64_t)sizeof(double);
    int64_t b = (int64_t)PAIR_BLOCKS * sizeof(double);
    return (a > b ? a : b) + 256;
}

int gp_rect_apply(const double* X, int64_t M, int64_t N, int64_t ldx, const double* R, int64_t p, int64_t ldr, double* Y,
                  int64_t ldy, double alpha, double beta, void* stream) {
    if (!X || !R || !Y || M <= 0 || N <= 0 || p <= 0 || p > RP || ldx < N || ldr < p || ldy < p) return -1;
    int64_t blocks = (M + 7) / 8;
    if (blocks > 148 * 16) blocks = 148 * 16;
    rect_apply_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(X, M, N, ldx, R, (int)p, ldr, Y, ldy, alpha, beta);
    GP_COUNT(1);
    GP_LAUNCH_CHECK();
    return 0;
}

int gp_rect_apply_t(const double* X, int64_t M, int64_t N, int64_t ldx, const double* Y, int64_t p, int64_t ldy, double* S,
                    int64_t lds, double alpha, double beta, void* ws, void* stream) {
    if (!X || !Y || !S || !ws || M <= 0 || N <= 0 || p <= 0 || p > RP || ldx < N || ldy < p || lds < p) return -1;
    cudaStream_t s = (cudaStream_t)stream;
    int nchunks = (int)((M + RCH - 1) / RCH);
    dim3 grid((unsigned)((N + 255) / 256), (unsigned)nchunks);
    rect_apply_t_partial_kernel<<<grid, 256, 0, s>>>(X, M, N, ldx, Y, (int)p, ldy, (double*)ws);
    int64_t total = N * p;
    rect_apply_t_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>((const double*)ws, total, nchunks, S, (int)p, lds, alpha, beta);
    GP_COUNT(2);
    GP_LAUNCH_CHECK();
    return 0;
}

int gp_pair_dot(const double* A, int64_t lda, const double* B, int64_t ldb, int64_t rows, int64_t cols, int64_t rows_w1,
                double w_rest, double* accum_dev, void* ws, void* stream) {
    if (!A || !B || !accum_dev || !ws || rows <= 0 || cols <= 0 || lda < cols || ldb < cols) return -1;
    cudaStream_t s = (cudaStream_t)stream;
    pair_dot_partial_kernel<<<PAIR_BLOCKS, 256, 0, s>>>(A, lda, B, ldb, rows, cols, rows_w1, w_rest, (double*)ws);
    pair_dot_final_kernel<<<1, 256, 0, s>>>((const double*)ws, PAIR_BLOCKS, accum_dev);
    GP_COUNT(2);
    GP_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
