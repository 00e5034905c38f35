// Symmetric eigenvalue path on own kernels: imate_method = 'eigenvalue', the method the reference's Likelihood hard-codes
// (gaussian_proc/_likelihood/likelihood.py:41; mixed_correlation.py:76-79 computes ALL eigenvalues of K once with
// scipy.linalg.eigh(eigvals_only=True), after which logdet / traceinv / trace are O(n) reductions, :127-136,172-181,239-248).
//
//   gp_sytrd_f64      K = Q T Q^T, T tridiagonal (d, e), Q = H_0 H_1 ... H_{n-3} kept as Householder vectors.
//                     One-stage reduction with the rank-2 update of step k-1 FUSED into the matrix-vector product of step k:
//                     the trailing matrix is read and written exactly once per column (16 B per element per column,
//                     HBM-bound: 16 n^3 / 3 bytes in total), two launches per column:
//                       sytd_vec_kernel  (one CTA)   w_{k-1} from p_{k-1}; column k after the pending update; reflector v_k
//                       sytd_pass_kernel (row strips) A <- A - v_{k-1} w_{k-1}^T - w_{k-1} v_{k-1}^T ;  p_k = A v_k
//                     The whole symmetric matrix is kept (both triangles), so the product needs no transposed partial sums
//                     and every reduction has a fixed order (bit-reproducible).
//   gp_stebz_f64      all eigenvalues of T by bisection on Sturm counts, one thread per eigenvalue (absolute accuracy
//                     ~ eps ||T||, as LAPACK dstebz).
//   gp_ormtr_skinny   R <- Q^T R or Q R for a skinny block (n x p, p <= 16): the reflectors applied in sequence by one CTA.
//   gp_tridiag_solve  (T + eta I) Y = A for p right-hand sides + log det (T + eta I) from the pivots (LDL^T, one warp).
//   gp_eig_reduce     sum log(lam + eta), sum 1 / (lam + eta), sum 1 / (lam + eta)^2 (fixed-order reduction).
#include "../../include/gpgp.h"
#include "gp_common.cuh"
#include "gp_internal.h"
#include <float.h>

namespace gp {

constexpr int EP = 16;            // max skinny width
constexpr int VEC_THREADS = 1024;

// block-wide sum broadcast to every thread (red: >= 33 doubles)
__device__ __forceinline__ double block_sum_all(double v, double* red) {
    v = block_sum(v, red);
    if (threadIdx.x == 0) red[32] = v;
    __syncthreads();
    v = red[32];
    __syncthreads();
    return v;
}

// One CTA. On entry (k >= 1): p = A^(k-1) v_{k-1} (unscaled), v = v_{k-1} (v[k] = 1, zero at indices < k), tau_prev.
// Produces w = w_{k-1}; the diagonal entry d[k] and the column k of A^(k) (from row k of the symmetric matrix);
// the reflector v_k (written to vnext and kept in row k of A right of the diagonal), tau[k], e[k].
// k == 0 has no pending update (first = 1).
__global__ void __launch_bounds__(VEC_THREADS)
sytd_vec_kernel(double* __restrict__ A, int n, int64_t lda, int k, int first, const double* __restrict__ p,
                const double* __restrict__ v, double* __restrict__ w, double* __restrict__ vnext, double* __restrict__ d,
                double* __restrict__ e, double* __restrict__ tau) {
    __shared__ double red[40];
    const int tid = threadIdx.x;
    double tprev = 0.0, vk = 0.0, wk = 0.0;
    if (!first) {
        tprev = tau[k - 1];
        // w = tau p - (tau^2 / 2) (p^T v) v   on indices >= k
        double s = 0.0;
        for (int i = k + tid; i < n; i += VEC_THREADS) s += p[i] * v[i];
        s = block_sum_all(s, red);
        const double alpha = 0.5 * tprev * tprev * s;
        for (int i = k + tid; i < n; i += VEC_THREADS) w[i] = tprev * p[i] - alpha * v[i];
        __syncthreads();
        vk = v[k];
        wk = w[k];
    }
    // column k of the updated trailing matrix (read as row k: the matrix is kept symmetric), its diagonal entry first
    double* rowk = A + (int64_t)k * lda;
    if (tid == 0) d[k] = first ? rowk[k] : rowk[k] - 2.0 * vk * wk;
    double ss = 0.0;
    for (int i = k + 1 + tid; i < n; i += VEC_THREADS) {
        double c = rowk[i];
        if (!first) c -= v[i] * wk + w[i] * vk;
        vnext[i] = c;                       // temporarily the column itself
        if (i > k + 1) ss += c * c;
    }
    ss = block_sum_all(ss, red);            // sum of squares below the first sub-diagonal entry
    __syncthreads();
    const double x0 = vnext[k + 1];
    double beta, t, scale;
    if (ss == 0.0) {                        // already tridiagonal in this column: H = I
        beta = x0; t = 0.0; scale = 0.0;
    } else {
        const double nrm = sqrt(x0 * x0 + ss);
        beta = (x0 >= 0.0) ? -nrm : nrm;
        t = (beta - x0) / beta;
        scale = 1.0 / (x0 - beta);
    }
    __syncthreads();
    for (int i = k + 1 + tid; i < n; i += VEC_THREADS) {
        double val = (i == k + 1) ? 1.0 : vnext[i] * scale;
        vnext[i] = val;
        rowk[i] = val;                      // the reflector is kept in row k of A (LAPACK keeps it in the column)
    }
    if (tid == 0) {
        vnext[k] = 0.0;
        e[k] = beta;
        tau[k] = t;
    }
}

// Row strips: one warp per row i > k. A[i][j] (j > k) receives the pending rank-2 update and is multiplied by vnext.
constexpr int PASS_ROWS = 8;
__global__ void __launch_bounds__(PASS_ROWS * 32)
sytd_pass_kernel(double* __restrict__ A, int n, int64_t lda, int k, int first, const double* __restrict__ v,
                 const double* __restrict__ w, const double* __restrict__ vnext, double* __restrict__ p) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i = k + 1 + blockIdx.x * PASS_ROWS + warp;
    if (i >= n) return;
    double* row = A + (int64_t)i * lda;
    const double vi = first ? 0.0 : v[i], wi = first ? 0.0 : w[i];
    double acc = 0.0;
    // columns k+1 .. n-1, two per lane per step (16-byte accesses once j is even)
    int j = k + 1;
    if (j & 1) {                            // peel to an even column
        if (lane == 0) {
            double a = row[j];
            if (!first) { a -= vi * w[j] + wi * v[j]; row[j] = a; }
            acc += a * vnext[j];
        }
        ++j;
    }
#pragma unroll 4
    for (int jj = j + 2 * lane; jj < n; jj += 64) {
        if (jj + 1 < n) {
            double2 a = *reinterpret_cast<const double2*>(row + jj);
            if (!first) {
                const double2 wj = *reinterpret_cast<const double2*>(w + jj);
                const double2 vj = *reinterpret_cast<const double2*>(v + jj);
                a.x -= vi * wj.x + wi * vj.x;
                a.y -= vi * wj.y + wi * vj.y;
                *reinterpret_cast<double2*>(row + jj) = a;
            }
            const double2 vn = *reinterpret_cast<const double2*>(vnext + jj);
            acc += a.x * vn.x + a.y * vn.y;
        } else {
            double a = row[jj];
            if (!first) { a -= vi * w[jj] + wi * v[jj]; row[jj] = a; }
            acc += a * vnext[jj];
        }
    }
    acc = warp_sum(acc);
    if (lane == 0) p[i] = acc;
}

// ---- eigenvalues of the tridiagonal matrix by bisection ----------------------------------------------------------------
// number of eigenvalues < x (Sturm sequence of the LDL^T pivots with the usual safeguard against a zero pivot)
__device__ __forceinline__ int sturm_count(const double* __restrict__ d, const double* __restrict__ e2, int n, double x, double pivmin) {
    int cnt = 0;
    double q = d[0] - x;
    if (fabs(q) < pivmin) q = -pivmin;
    cnt += (q < 0.0);
    for (int i = 1; i < n; ++i) {
        q = d[i] - x - e2[i - 1] / q;
        if (fabs(q) < pivmin) q = -pivmin;
        cnt += (q < 0.0);
    }
    return cnt;
}

__global__ void square_offdiag_kernel(const double* __restrict__ e, int n, double* __restrict__ e2) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n - 1) e2[i] = e[i] * e[i];
}

__global__ void __launch_bounds__(128)
stebz_kernel(const double* __restrict__ d, const double* __restrict__ e2, int n, double gl, double gu, double pivmin,
             double* __restrict__ lam) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;      // the j-th smallest eigenvalue
    if (j >= n) return;
    double lo = gl, hi = gu;
    for (int it = 0; it < 110; ++it) {
        const double mid = 0.5 * (lo + hi);
        if (mid <= lo || mid >= hi) break;                      // interval at machine resolution
        if (sturm_count(d, e2, n, mid, pivmin) > j) hi = mid; else lo = mid;
    }
    lam[j] = 0.5 * (lo + hi);
}

// ---- Householder vectors applied to a skinny block ------------------------------------------------------------------------
// R (n x p, ldr) <- Q^T R (trans = 1: k = 0 .. n-3) or Q R (trans = 0: k = n-3 .. 0); reflector k lives in A[k][k+1 ..]
__global__ void __launch_bounds__(1024)
ormtr_skinny_kernel(const double* __restrict__ A, int n, int64_t lda, const double* __restrict__ tau, int trans,
                    double* __restrict__ R, int p, int64_t ldr) {
    __shared__ double red[32 * EP];
    __shared__ double dots[EP];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int step = 0; step < n - 2; ++step) {
        const int k = trans ? step : (n - 3 - step);
        const double t = tau[k];
        if (t == 0.0) continue;                                 // uniform over the CTA
        const double* v = A + (int64_t)k * lda;
        double acc[EP];
#pragma unroll
        for (int c = 0; c < EP; ++c) acc[c] = 0.0;
        for (int i = k + 1 + tid; i < n; i += 1024) {
            const double vi = v[i];
            const double* r = R + (int64_t)i * ldr;
#pragma unroll
            for (int c = 0; c < EP; ++c)
                if (c < p) acc[c] += vi * r[c];
        }
#pragma unroll
        for (int c = 0; c < EP; ++c)
            if (c < p) {
                double s = warp_sum(acc[c]);
                if (lane == 0) red[warp * EP + c] = s;
            }
        __syncthreads();
        if (tid < p) {
            double s = 0.0;
            for (int wq = 0; wq < 32; ++wq) s += red[wq * EP + tid];
            dots[tid] = t * s;
        }
        __syncthreads();
        for (int i = k + 1 + tid; i < n; i += 1024) {
            const double vi = v[i];
            double* r = R + (int64_t)i * ldr;
#pragma unroll
            for (int c = 0; c < EP; ++c)
                if (c < p) r[c] -= vi * dots[c];
        }
        __syncthreads();
    }
}

// ---- (T + eta I) Y = B, p right-hand sides ------------------------------------------------------------------------------
// One CTA. The recurrences are sequential in i, so the only way to make them fast is to keep every operand of the inner
// loop in shared memory: the CTA stages chunks of TCH rows (d, e, the right-hand sides) cooperatively, warp 0 runs the
// recurrence on the chunk (lane c < p owns column c, the pivots are recomputed by every lane), and the chunk is written back
// cooperatively. log det (T + eta I) = sum log |delta_i| is summed in parallel from the stored pivots afterwards.
constexpr int TCH = 256;
__global__ void __launch_bounds__(256)
tridiag_solve_kernel(const double* __restrict__ d, const double* __restrict__ e, int n, double eta, const double* __restrict__ B,
                     int p, int64_t ldb, double* __restrict__ Y, int64_t ldy, double* __restrict__ piv, double* __restrict__ out) {
    __shared__ double sd[TCH], se[TCH], sp[TCH];
    __shared__ double sy[TCH * EP];
    __shared__ double red[40];
    __shared__ double carry[EP + 2];          // [0..p): last y / x of the previous chunk, [EP]: last pivot, [EP + 1]: e before the chunk
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // ---- forward: delta_i = d_i + eta - e_{i-1}^2 / delta_{i-1};  y_i = b_i - (e_{i-1} / delta_{i-1}) y_{i-1}
    for (int c0 = 0; c0 < n; c0 += TCH) {
        const int cnt = min(TCH, n - c0);
        for (int r = tid; r < cnt; r += 256) {
            sd[r] = d[c0 + r] + eta;
            se[r] = (c0 + r > 0) ? e[c0 + r - 1] : 0.0;       // the off-diagonal entry ABOVE row c0 + r
        }
        for (int idx = tid; idx < cnt * p; idx += 256) sy[(idx / p) * EP + idx % p] = B[(int64_t)(c0 + idx / p) * ldb + idx % p];
        __syncthreads();
        if (warp == 0) {
            double delta = (c0 > 0) ? carry[EP] : 1.0;
            double y = (c0 > 0 && lane < p) ? carry[lane] : 0.0;
            for (int r = 0; r < cnt; ++r) {
                const double l = se[r] / delta;                // zero for the very first row
                delta = sd[r] - l * se[r];
                y = ((lane < p) ? sy[r * EP + lane] : 0.0) - l * y;
                if (lane == 0) sp[r] = delta;
                if (lane < p) sy[r * EP + lane] = y;
            }
            if (lane < p) carry[lane] = y;
            if (lane == 0) carry[EP] = delta;
        }
        __syncthreads();
        for (int r = tid; r < cnt; r += 256) piv[c0 + r] = sp[r];
        for (int idx = tid; idx < cnt * p; idx += 256) Y[(int64_t)(c0 + idx / p) * ldy + idx % p] = sy[(idx / p) * EP + idx % p];
        __syncthreads();
    }
    // ---- backward: x_i = (y_i - e_i x_{i+1}) / delta_i
    const int nchunks = (n + TCH - 1) / TCH;
    for (int ch = nchunks - 1; ch >= 0; --ch) {
        const int c0 = ch * TCH, cnt = min(TCH, n - c0);
        for (int r = tid; r < cnt; r += 256) {
            sp[r] = piv[c0 + r];
            se[r] = (c0 + r + 1 < n) ? e[c0 + r] : 0.0;         // the off-diagonal entry BELOW row c0 + r
        }
        for (int idx = tid; idx < cnt * p; idx += 256) sy[(idx / p) * EP + idx % p] = Y[(int64_t)(c0 + idx / p) * ldy + idx % p];
        __syncthreads();
        if (warp == 0) {
            double x = (ch < nchunks - 1 && lane < p) ? carry[lane] : 0.0;
            for (int r = cnt - 1; r >= 0; --r) {
                x = (((lane < p) ? sy[r * EP + lane] : 0.0) - se[r] * x) / sp[r];
                if (lane < p) sy[r * EP + lane] = x;
            }
            if (lane < p) carry[lane] = x;
        }
        __syncthreads();
        for (int idx = tid; idx < cnt * p; idx += 256) Y[(int64_t)(c0 + idx / p) * ldy + idx % p] = sy[(idx / p) * EP + idx % p];
        __syncthreads();
    }
    // ---- log det and the number of non-positive pivots
    double ld = 0.0, neg = 0.0;
    for (int i = tid; i < n; i += 256) {
        const double dl = piv[i];
        ld += log(fabs(dl));
        neg += (dl <= 0.0) ? 1.0 : 0.0;
    }
    ld = block_sum_all(ld, red);
    neg = block_sum_all(neg, red);
    if (tid == 0) { out[0] = ld; out[1] = neg; }
}

__global__ void __launch_bounds__(1024)
eig_reduce_kernel(const double* __restrict__ lam, int n, double eta, double* __restrict__ out) {
    __shared__ double red[40];
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, bad = 0.0;
    for (int i = threadIdx.x; i < n; i += 1024) {
        const double x = lam[i] + eta;
        if (!(x > 0.0)) bad += 1.0;
        const double r = 1.0 / x;
        s0 += log(fabs(x));
        s1 += r;
        s2 += r * r;
    }
    s0 = block_sum_all(s0, red);
    s1 = block_sum_all(s1, red);
    s2 = block_sum_all(s2, red);
    bad = block_sum_all(bad, red);
    if (threadIdx.x == 0) { out[0] = s0; out[1] = s1; out[2] = s2; out[3] = bad; }
}

__global__ void gershgorin_kernel(const double* __restrict__ d, const double* __restrict__ e, int n, double* out) {
    __shared__ double lo[256], hi[256], mx[256];
    double l = DBL_MAX, h = -DBL_MAX, m = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) {
        double r = ((i > 0) ? fabs(e[i - 1]) : 0.0) + ((i + 1 < n) ? fabs(e[i]) : 0.0);
        l = fmin(l, d[i] - r);
        h = fmax(h, d[i] + r);
        if (i + 1 < n) m = fmax(m, e[i] * e[i]);
    }
    lo[threadIdx.x] = l; hi[threadIdx.x] = h; mx[threadIdx.x] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < 256; ++i) { l = fmin(l, lo[i]); h = fmax(h, hi[i]); m = fmax(m, mx[i]); }
        out[0] = l; out[1] = h; out[2] = m;
    }
}

}  // namespace gp

using namespace gp;

extern "C" {

// A (n x n, lda, symmetric, BOTH triangles valid) is overwritten: diagonal -> d, row k right of the diagonal -> reflector k.
// d (n), e (n), tau (n), ws: 4 n doubles. Everything stream-ordered, no host synchronisation.
int64_t gp_sytrd_workspace_bytes(int64_t n) { return 4 * ((n + 1) & ~(int64_t)1) * (int64_t)sizeof(double) + 256; }

int gp_sytrd_f64(double* A, int64_t n, int64_t lda, double* d, double* e, double* tau, void* ws, void* stream) {
    if (!A || !d || !e || !tau || !ws || n <= 0 || n > INT32_MAX || lda < n || (lda & 1) || ((uintptr_t)A & 15)) return -1;
    cudaStream_t s = (cudaStream_t)stream;
    if ((uintptr_t)ws & 15) return -2;
    const int64_t nn = (n + 1) & ~(int64_t)1;      // even stride: every vector stays 16-byte aligned
    double* p = (double*)ws;
    double* va = p + nn;         // the two reflector buffers alternate between "previous" and "next"
    double* vb = va + nn;
    double* w = vb + nn;
    GP_CUDA_CHECK(cudaMemsetAsync(ws, 0, 4 * nn * sizeof(double), s));
    GP_CUDA_CHECK(cudaMemsetAsync(e, 0, n * sizeof(double), s));
    GP_CUDA_CHECK(cudaMemsetAsync(tau, 0, n * sizeof(double), s));
    const int N = (int)n;
    for (int k = 0; k + 2 < N; ++k) {
        double* vprev = (k & 1) ? vb : va;
        double* vnext = (k & 1) ? va : vb;
        sytd_vec_kernel<<<1, VEC_THREADS, 0, s>>>(A, N, lda, k, k == 0, p, vprev, w, vnext, d, e, tau);
        const int rows = N - k - 1;
        sytd_pass_kernel<<<(rows + PASS_ROWS - 1) / PASS_ROWS, PASS_ROWS * 32, 0, s>>>(A, N, lda, k, k == 0, vprev, w, vnext, p);
    }
    GP_COUNT(2 * (N > 2 ? N - 2 : 0));
    // the last 2 x 2 block: apply the pending update of step n-3 and read d[n-2], d[n-1], e[n-2]
    if (N >= 3) {
        const int k = N - 2;
        double* vprev = (k & 1) ? vb : va;
        double* vnext = (k & 1) ? va : vb;
        sytd_vec_kernel<<<1, VEC_THREADS, 0, s>>>(A, N, lda, k, 0, p, vprev, w, vnext, d, e, tau);   // d[n-2], e[n-2] (= the entry), tau = 0
        sytd_pass_kernel<<<1, PASS_ROWS * 32, 0, s>>>(A, N, lda, k, 0, vprev, w, vnext, p);          // updates A[n-1][n-1]
        GP_CUDA_CHECK(cudaMemcpyAsync(d + (N - 1), A + (int64_t)(N - 1) * lda + (N - 1), sizeof(double), cudaMemcpyDeviceToDevice, s));
        GP_COUNT(2);
    } else if (N == 2) {
        GP_CUDA_CHECK(cudaMemcpyAsync(d, A, sizeof(double), cudaMemcpyDeviceToDevice, s));
        GP_CUDA_CHECK(cudaMemcpyAsync(d + 1, A + lda + 1, sizeof(double), cudaMemcpyDeviceToDevice, s));
        GP_CUDA_CHECK(cudaMemcpyAsync(e, A + 1, sizeof(double), cudaMemcpyDeviceToDevice, s));
    } else {
        GP_CUDA_CHECK(cudaMemcpyAsync(d, A, sizeof(double), cudaMemcpyDeviceToDevice, s));
    }
    GP_LAUNCH_CHECK();
    return 0;
}

// all eigenvalues (ascending) of the symmetric tridiagonal (d, e); ws: n + 8 doubles
int gp_stebz_f64(const double* d, const double* e, int64_t n, double* lam, void* ws, void* stream) {
    if (!d || !e || !lam || !ws || n <= 0 || n > INT32_MAX) return -1;
    cudaStream_t s = (cudaStream_t)stream;
    double* e2 = (double*)ws;
    double* bounds = e2 + n;
    if (n > 1) square_offdiag_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(e, (int)n, e2);
    gershgorin_kernel<<<1, 256, 0, s>>>(d, e, (int)n, bounds);
    double hb[3];
    GP_CUDA_CHECK(cudaMemcpyAsync(hb, bounds, sizeof(hb), cudaMemcpyDeviceToHost, s));
    GP_CUDA_CHECK(cudaStreamSynchronize(s));
    const double tnorm = fmax(fabs(hb[0]), fabs(hb[1]));
    const double gl = hb[0] - 2.0 * DBL_EPSILON * tnorm * (double)n - 2.0 * DBL_MIN;
    const double gu = hb[1] + 2.0 * DBL_EPSILON * tnorm * (double)n + 2.0 * DBL_MIN;
    const double pivmin = fmax(DBL_MIN, DBL_MIN * hb[2]) * 4.0;
    stebz_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(d, e2, (int)n, gl, gu, pivmin, lam);
    GP_COUNT(3);
    GP_LAUNCH_CHECK();
    return 0;
}

int gp_ormtr_skinny(const double* A, int64_t n, int64_t lda, const double* tau, int trans, double* R, int64_t p, int64_t ldr,
                    void* stream) {
    if (!A || !tau || !R || n <= 0 || p <= 0 || p > EP || ldr < p) return -1;
    if (n >= 3) {
        ormtr_skinny_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(A, (int)n, lda, tau, trans ? 1 : 0, R, (int)p, ldr);
        GP_COUNT(1);
        GP_LAUNCH_CHECK();
    }
    return 0;
}

// Y (n x p) <- (T + eta I)^-1 B; out (device, 2 doubles): log det (T + eta I), number of non-positive pivots; ws: n doubles
int gp_tridiag_solve(const double* d, const double* e, int64_t n, double eta, const double* B, int64_t p, int64_t ldb, double* Y,
                     int64_t ldy, void* ws, double* out, void* stream) {
    if (!d || !e || !B || !Y || !ws || !out || n <= 0 || p <= 0 || p > EP || ldb < p || ldy < p) return -1;
    tridiag_solve_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(d, e, (int)n, eta, B, (int)p, ldb, Y, ldy, (double*)ws, out);
    GP_COUNT(1);
    GP_LAUNCH_CHECK();
    return 0;
}

// out (device, 4 doubles): sum log(lam + eta), sum 1/(lam + eta), sum 1/(lam + eta)^2, count of lam + eta <= 0
int gp_eig_reduce(const double* lam, int64_t n, double eta, double* out, void* stream) {
    if (!lam || !out || n <= 0) return -1;
    eig_reduce_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(lam, (int)n, eta, out);
    GP_COUNT(1);
    GP_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
