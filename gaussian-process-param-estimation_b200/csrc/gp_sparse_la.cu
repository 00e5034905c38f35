// Sparse (CSR) operator K + eta*I on blocks of probe / right-hand-side columns: SpMM, batched Lanczos (stochastic
// Lanczos quadrature for logdet and trace of the inverse) and batched CG (solves, Hutchinson tr(Kn^-1 dK)).
// Replaces what the reference reaches through imate's 'slq' / 'hutchinson' methods and scipy.sparse.linalg.cg
// (gaussian_proc/_mixed_correlation/mixed_correlation.py:193-209,263-268; _linear_solver.py:49-68, tol = 1e-6).
// All column blocks are n x B row-major with B in {1,2,4,8,16,32}; HBM-bound: 12 nnz + 4 (n+1) + 16 n B bytes / SpMM.
#include "../../include/gpgp.h"
#include "gp_common.cuh"

namespace gp {

constexpr int RED_PARTS = 592;  // CTAs of the column reductions (4 per SM)

// ---- Y = (K + eta I) X, one warp per row; lane = (q, c): q-th nonzero of the current group, column c ---------
template <int B>
__global__ void __launch_bounds__(256)
csr_spmm_kernel(const int* __restrict__ indptr, const int* __restrict__ indices, const double* __restrict__ data, int n,
                double eta, const double* __restrict__ X, double* __restrict__ Y) {
    constexpr int NQ = 32 / B;
    const int lane = threadIdx.x & 31;
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= n) return;
    const int q = lane / B, c = lane % B;
    const int s0 = indptr[row], s1 = indptr[row + 1];
    double acc = 0.0;
    for (int p = s0 + q; p < s1; p += NQ) acc += data[p] * X[(int64_t)indices[p] * B + c];
#pragma unroll
    for (int o = B; o < 32; o <<= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (q == 0) Y[(int64_t)row * B + c] = acc + eta * X[(int64_t)row * B + c];
}

// ---- column reductions: partial[cta][c] = sum over the CTA's elements of column c --------------------------------
// mode 0: x*y   mode 1: (lanczos) w -= a q + b qprev, accumulate w*w   mode 2: (cg) x += a p, r -= a ap, accumulate r*r
template <int MODE>
__global__ void __launch_bounds__(256)
col_fused_kernel(int64_t total, int B, const double* __restrict__ X, const double* __restrict__ Y, double* W,
                 double* Z, const double* __restrict__ a, const double* __restrict__ b, double* partial) {
    __shared__ double red[256];
    const int64_t stride = (int64_t)gridDim.x * 256;  // multiple of 32 >= B: a thread stays in one column
    const int c = threadIdx.x % B;
    double acc = 0.0;
    double ac = (MODE != 0) ? a[c] : 0.0, bc = (MODE == 1) ? b[c] : 0.0;
    for (int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x; idx < total; idx += stride) {
        if (MODE == 0) {
            acc += X[idx] * Y[idx];
        } else if (MODE == 1) {
            double w = W[idx] - ac * X[idx] - bc * Y[idx];   // X = q_j, Y = q_{j-1}
            W[idx] = w;
            acc += w * w;
        } else {
            Z[idx] += ac * X[idx];                            // Z = solution, X = p
            double r = W[idx] - ac * Y[idx];                  // W = residual, Y = A p
            W[idx] = r;
            acc += r * r;
        }
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    // threads t, t+B, t+2B, ... share column t % B (256 % B == 0)
    if (threadIdx.x < B) {
        double s = 0.0;
        for (int t = threadIdx.x; t < 256; t += B) s += red[t];
        partial[(int64_t)blockIdx.x * B + threadIdx.x] = s;
    }
}

// final stage + the scalar recurrences, one thread per column (fixed order -> reproducible)
// op 0: out = sum
// op 1 (lanczos alpha): alpha[c] = sum; store into coef row
// op 2 (lanczos beta):  beta[c] = sqrt(sum); store; inv[c] = beta > tiny ? 1/beta : 0
// op 3 (cg pAp):        alpha[c] = active ? rr/sum : 0
// op 4 (cg rr_new):     beta[c] = active ? sum/rr : 0; rr = sum; active &= rr > tol2*bb
__global__ void col_final_kernel(const double* partial, int nparts, int B, int op, double* out, double* s1, double* s2,
                                 double* s3, double tol2) {
    int c = threadIdx.x;
    if (c >= B) return;
    double s = 0.0;
    for (int i = 0; i < nparts; ++i) s += partial[(int64_t)i * B + c];
    if (op == 0) {
        out[c] = s;
    } else if (op == 1) {
        out[c] = s;
        s1[c] = s;
    } else if (op == 2) {
        double bt = sqrt(s);
        out[c] = bt;
        s1[c] = bt;
        s2[c] = (bt > 1e-300) ? 1.0 / bt : 0.0;
    } else if (op == 3) {
        // s1 = rr, s2 = active flag (1/0), s3 = breakdown flag, out = alpha
        if (s2[c] != 0.0 && !(s > 0.0)) { s3[0] = 1.0; s2[c] = 0.0; }   // p^T A p <= 0: A is not positive definite
        out[c] = (s2[c] != 0.0) ? s1[c] / s : 0.0;
    } else {
        // s1 = rr (updated), s2 = active, s3 = bb, out = beta
        double rr_old = s1[c];
        out[c] = (s2[c] != 0.0 && rr_old != 0.0) ? s / rr_old : 0.0;
        s1[c] = s;
        if (!(s > tol2 * s3[c])) s2[c] = 0.0;
    }
}

// q_next = w * inv_beta (per column); also used for the initial normalisation
__global__ void col_scale_kernel(int64_t total, int B, const double* __restrict__ W, const double* __restrict__ inv, double* Q) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    Q[idx] = W[idx] * inv[idx % B];
}

// p = r + beta p
__global__ void cg_direction_kernel(int64_t total, int B, const double* __restrict__ R, const double* __restrict__ beta, double* P) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    P[idx] = R[idx] + beta[idx % B] * P[idx];
}

// Rademacher probes from a counter-based hash of (seed, probe id, row): independent of batching and rank count
__global__ void rademacher_kernel(int64_t n, int B, uint64_t seed, int64_t probe0, const int* __restrict__ row_map, double* V) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * B) return;
    int64_t r = idx / B;
    uint64_t row = (uint64_t)(row_map ? row_map[r] : r), pid = (uint64_t)(probe0 + idx % B);
    uint64_t z = seed * 0x9E3779B97F4A7C15ull + pid * 0xBF58476D1CE4E5B9ull + row * 0x94D049BB133111EBull + 0x2545F4914F6CDD1Dull;
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull;
    z ^= z >> 27; z *= 0x94D049BB133111EBull;
    z ^= z >> 31;
    V[idx] = (z & 1ull) ? 1.0 : -1.0;
}

static int spmm(const int* indptr, const int* indices, const double* data, int n, double eta, const double* X, int B, double* Y,
                cudaStream_t s) {
    int blocks = (int)(((int64_t)n * 32 + 255) / 256);
    switch (B) {
        case 1: csr_spmm_kernel<1><<<blocks, 256, 0, s>>>(indptr, indices, data, n, eta, X, Y); break;
        case 2: csr_spmm_kernel<2><<<blocks, 256, 0, s>>>(indptr, indices, data, n, eta, X, Y); break;
        case 4: csr_spmm_kernel<4><<<blocks, 256, 0, s>>>(indptr, indices, data, n, eta, X, Y); break;
        case 8: csr_spmm_kernel<8><<<blocks, 256, 0, s>>>(indptr, indices, data, n, eta, X, Y); break;
        case 16: csr_spmm_kernel<16><<<blocks, 256, 0, s>>>(indptr, indices, data, n, eta, X, Y); break;
        case 32: csr_spmm_kernel<32><<<blocks, 256, 0, s>>>(indptr, indices, data, n, eta, X, Y); break;
        default: return -2;
    }
    GP_COUNT(1);
    GP_LAUNCH_CHECK();
    return 0;
}

// new row r = old row order[r]; columns mapped through inv_order; one warp per new row
__global__ void __launch_bounds__(256)
csr_permute_kernel(int n, const int* __restrict__ order, const int* __restrict__ inv_order, const int* __restrict__ indptr,
                   const int* __restrict__ indices, const double* __restrict__ data, const double* __restrict__ ddata,
                   const int* __restrict__ new_indptr, int* new_indices, double* new_data, double* new_ddata) {
    const int lane = threadIdx.x & 31;
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= n) return;
    const int o = order[r];
    const int s0 = indptr[o], len = indptr[o + 1] - s0, d0 = new_indptr[r];
    for (int t = lane; t < len; t += 32) {
        new_indices[d0 + t] = inv_order[indices[s0 + t]];
        new_data[d0 + t] = data[s0 + t];
        if (ddata) new_ddata[d0 + t] = ddata[s0 + t];
    }
}

static size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace gp

using namespace gp;

extern "C" {

int gp_csr_spmm(const int* indptr, const int* indices, const double* data, int64_t n, double eta, const double* X, int64_t B,
                double* Y, void* stream) {
    if (!indptr || !indices || !data || !X || !Y || n <= 0 || n > INT32_MAX) return -1;
    return spmm(indptr, indices, data, (int)n, eta, X, (int)B, Y, (cudaStream_t)stream);
}

int gp_csr_permute(int64_t n, const int* order, const int* inv_order, const int* indptr, const int* indices,
                   const double* data, const double* ddata, const int* new_indptr, int* new_indices, double* new_data,
                   double* new_ddata, void* stream) {
    if (!order || !inv_order || !indptr || !indices || !data || !new_indptr || !new_indices || !new_data || n <= 0 || n > INT32_MAX)
        return -1;
    csr_permute_kernel<<<(unsigned)((n * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        (int)n, order, inv_order, indptr, indices, data, ddata, new_indptr, new_indices, new_data, new_ddata);
    GP_COUNT(1);
    GP_LAUNCH_CHECK();
    return 0;
}

int gp_rademacher(double* V, int64_t n, int64_t B, uint64_t seed, int64_t probe_offset, const int* row_map, void* stream) {
    if (!V || n <= 0 || B <= 0) return -1;
    rademacher_kernel<<<(unsigned)((n * B + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, (int)B, seed, probe_offset, row_map, V);
    GP_COUNT(1);
    GP_LAUNCH_CHECK();
    return 0;
}

// out[c] = sum_i X[i][c] * Y[i][c]
int64_t gp_krylov_workspace_bytes(int64_t n, int64_t B) {
    return (int64_t)(4 * al256(sizeof(double) * n * B) + al256(sizeof(double) * RED_PARTS * 32) + 16 * al256(sizeof(double) * 32));
}

int gp_col_dot(const double* X, const double* Y, int64_t n, int64_t B, double* out_dev, void* ws, void* stream) {
    if (!X || !Y || !out_dev || !ws || B <= 0 || B > 32 || (32 % B)) return -1;
    cudaStream_t s = (cudaStream_t)stream;
    double* partial = (double*)ws;
    col_fused_kernel<0><<<RED_PARTS, 256, 0, s>>>(n * B, (int)B, X, Y, nullptr, nullptr, nullptr, nullptr, partial);
    col_final_kernel<<<1, 32, 0, s>>>(partial, RED_PARTS, (int)B, 0, out_dev, nullptr, nullptr, nullptr, 0.0);
    GP_COUNT(2);
    GP_LAUNCH_CHECK();
    return 0;
}

// m Lanczos steps of A = K + eta I started from the columns of V (normalised internally).
// alpha_dev, beta_dev: (m x B) row-major; beta[j] is the norm of the (j+1)-th unnormalised vector.
int gp_lanczos(const int* indptr, const int* indices, const double* data, int64_t n, double eta, const double* V, int64_t B,
               int64_t m, double* alpha_dev, double* beta_dev, void* ws, void* stream) {
    if (!indptr || !indices || !data || !V || !alpha_dev || !beta_dev || !ws || n <= 0 || n > INT32_MAX || m <= 0) return -1;
    if (B <= 0 || B > 32 || (32 % B)) return -2;
    cudaStream_t s = (cudaStream_t)stream;
    const int N = (int)n, Bc = (int)B;
    const int64_t total = n * B;
    char* base = (char*)ws;
    size_t vb = al256(sizeof(double) * total);
    double* Q0 = (double*)base;
    double* Q1 = (double*)(base + vb);
    double* W = (double*)(base + 2 * vb);
    double* partial = (double*)(base + 4 * vb);
    double* sc = (double*)(base + 4 * vb + al256(sizeof(double) * RED_PARTS * 32));
    double* a = sc;            // current alpha
    double* bprev = sc + 32;   // beta_{j-1}
    double* inv = sc + 64;     // 1 / beta
    double* tmp = sc + 96;
    const unsigned eb = (unsigned)((total + 255) / 256);
    // q_0 = v / ||v||
    col_fused_kernel<0><<<RED_PARTS, 256, 0, s>>>(total, Bc, V, V, nullptr, nullptr, nullptr, nullptr, partial);
    col_final_kernel<<<1, 32, 0, s>>>(partial, RED_PARTS, Bc, 2, tmp, tmp, inv, nullptr, 0.0);
    col_scale_kernel<<<eb, 256, 0, s>>>(total, Bc, V, inv, Q0);
    GP_CUDA_CHECK(cudaMemsetAsync(Q1, 0, sizeof(double) * total, s));
    GP_CUDA_CHECK(cudaMemsetAsync(bprev, 0, sizeof(double) * 32, s));
    GP_COUNT(3);
    double* q = Q0;
    double* qprev = Q1;
    for (int64_t j = 0; j < m; ++j) {
        int rc = spmm(indptr, indices, data, N, eta, q, Bc, W, s);
        if (rc) return rc;
        col_fused_kernel<0><<<RED_PARTS, 256, 0, s>>>(total, Bc, q, W, nullptr, nullptr, nullptr, nullptr, partial);
        col_final_kernel<<<1, 32, 0, s>>>(partial, RED_PARTS, Bc, 1, alpha_dev + j * B, a, nullptr, nullptr, 0.0);
        col_fused_kernel<1><<<RED_PARTS, 256, 0, s>>>(total, Bc, q, qprev, W, nullptr, a, bprev, partial);
        col_final_kernel<<<1, 32, 0, s>>>(partial, RED_PARTS, Bc, 2, beta_dev + j * B, bprev, inv, nullptr, 0.0);
        col_scale_kernel<<<eb, 256, 0, s>>>(total, Bc, W, inv, qprev);  // q_{j+1} overwrites q_{j-1}
        GP_COUNT(5);
        double* t = q; q = qprev; qprev = t;
    }
    GP_LAUNCH_CHECK();
    return 0;
}

// Batched CG for (K + eta I) X = R0, all B columns at once, stop per column at ||r|| <= tol ||b|| (the reference's
// scipy cg tol=1e-6, atol=0). X: in = initial guess is ignored (zero start), out = solution. R0 is overwritten.
// iters_host receives the number of iterations performed. Returns 0, or 1 if maxiter was hit before convergence.
int gp_cg_solve(const int* indptr, const int* indices, const double* data, int64_t n, double eta, double* R0, double* X,
                int64_t B, double tol, int64_t maxiter, int64_t* iters_host, void* ws, void* stream) {
    if (!indptr || !indices || !data || !R0 || !X || !ws || n <= 0 || n > INT32_MAX) return -1;
    if (B <= 0 || B > 32 || (32 % B)) return -2;
    cudaStream_t s = (cudaStream_t)stream;
    const int N = (int)n, Bc = (int)B;
    const int64_t total = n * B;
    char* base = (char*)ws;
    size_t vb = al256(sizeof(double) * total);
    double* Pd = (double*)base;
    double* AP = (double*)(base + vb);
    double* partial = (double*)(base + 4 * vb);
    double* sc = (double*)(base + 4 * vb + al256(sizeof(double) * RED_PARTS * 32));
    double *rr = sc, *active = sc + 32, *bb = sc + 64, *alpha = sc + 96, *beta = sc + 128, *flag = sc + 160;
    const unsigned eb = (unsigned)((total + 255) / 256);
    const double tol2 = tol * tol;
    GP_CUDA_CHECK(cudaMemsetAsync(X, 0, sizeof(double) * total, s));
    GP_CUDA_CHECK(cudaMemsetAsync(flag, 0, sizeof(double) * 32, s));
    GP_CUDA_CHECK(cudaMemcpyAsync(Pd, R0, sizeof(double) * total, cudaMemcpyDeviceToDevice, s));
    col_fused_kernel<0><<<RED_PARTS, 256, 0, s>>>(total, Bc, R0, R0, nullptr, nullptr, nullptr, nullptr, partial);
    col_final_kernel<<<1, 32, 0, s>>>(partial, RED_PARTS, Bc, 0, rr, nullptr, nullptr, nullptr, 0.0);
    GP_CUDA_CHECK(cudaMemcpyAsync(bb, rr, sizeof(double) * 32, cudaMemcpyDeviceToDevice, s));
    double ones[32], act[32];
    GP_CUDA_CHECK(cudaMemcpyAsync(ones, rr, sizeof(double) * B, cudaMemcpyDeviceToHost, s));
    GP_CUDA_CHECK(cudaStreamSynchronize(s));
    for (int c = 0; c < 32; ++c) act[c] = (c < B && ones[c] > 0.0) ? 1.0 : 0.0;
    GP_CUDA_CHECK(cudaMemcpyAsync(active, act, sizeof(double) * 32, cudaMemcpyHostToDevice, s));
    GP_COUNT(2);
    int64_t it = 0;
    bool converged = false;
    const int check_every = 8;
    while (it < maxiter) {
        int rc = spmm(indptr, indices, data, N, eta, Pd, Bc, AP, s);
        if (rc) return rc;
        col_fused_kernel<0><<<RED_PARTS, 256, 0, s>>>(total, Bc, Pd, AP, nullptr, nullptr, nullptr, nullptr, partial);
        col_final_kernel<<<1, 32, 0, s>>>(partial, RED_PARTS, Bc, 3, alpha, rr, active, flag, 0.0);
        col_fused_kernel<2><<<RED_PARTS, 256, 0, s>>>(total, Bc, Pd, AP, R0, X, alpha, nullptr, partial);
        col_final_kernel<<<1, 32, 0, s>>>(partial, RED_PARTS, Bc, 4, beta, rr, active, bb, tol2);
        cg_direction_kernel<<<eb, 256, 0, s>>>(total, Bc, R0, beta, Pd);
        GP_COUNT(5);
        ++it;
        if (it % check_every == 0 || it == maxiter) {
            double brk = 0.0;
            GP_CUDA_CHECK(cudaMemcpyAsync(act, active, sizeof(double) * 32, cudaMemcpyDeviceToHost, s));
            GP_CUDA_CHECK(cudaMemcpyAsync(&brk, flag, sizeof(double), cudaMemcpyDeviceToHost, s));
            GP_CUDA_CHECK(cudaStreamSynchronize(s));
            if (brk != 0.0) { if (iters_host) *iters_host = it; return 2; }
            bool any = false;
            for (int c = 0; c < B; ++c) any = any || (act[c] != 0.0);
            if (!any) { converged = true; break; }
        }
    }
    GP_LAUNCH_CHECK();
    if (iters_host) *iters_host = it;
    return converged ? 0 : 1;  // 2 (above): negative curvature met, K + eta I is not positive definite
}

}  // extern "C"
