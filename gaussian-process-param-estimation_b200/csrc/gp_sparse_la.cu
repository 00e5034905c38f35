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

// ---- row-blocked operator: R consecutive rows share one column list (R x 1 blocks, zero filled) -----------------
// Rows that are neighbours in a spatially sorted order have almost the same pattern, so one gathered row of X serves
// R rows of K: gather traffic (the L1/L2-bound part of a multi-column SpMM) drops ~R-fold and the column index is
// amortised over R values. Block-columns of a row block: [all columns of row 0][columns of row 1 not in row 0]...
// (each segment in the source row's order); any order is valid for the product.
__device__ __forceinline__ int find_col(const int* __restrict__ row, int len, int c) {
    int lo = 0, hi = len;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (row[mid] < c) lo = mid + 1; else hi = mid;
    }
    return (lo < len && row[lo] == c) ? lo : -1;
}

// one warp per row block. PASS 0: nblk[rb] = number of distinct columns; PASS 1: fill (bidx, bvals[, bdvals]).
// new row r = old row order[r]; new column = inv_order[old column]; the source CSR must have sorted rows.
template <int R, int PASS>
__global__ void __launch_bounds__(256)
bcsr_build_kernel(int n, const int* __restrict__ order, const int* __restrict__ inv_order, const int* __restrict__ indptr,
                  const int* __restrict__ indices, const double* __restrict__ data, const double* __restrict__ ddata,
                  int* nblk, const int64_t* __restrict__ bptr, int* bidx, double* bvals, double* bdvals) {
    const int lane = threadIdx.x & 31;
    const int rb = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (rb * R >= n) return;
    int s[R], len[R];
#pragma unroll
    for (int k = 0; k < R; ++k) {
        int r = rb * R + k;
        if (r < n) {
            int o = order ? order[r] : r;
            s[k] = indptr[o];
            len[k] = indptr[o + 1] - s[k];
        } else {
            s[k] = 0;
            len[k] = 0;
        }
    }
    const int64_t base = (PASS == 1) ? bptr[rb] : 0;
    int running = 0;
#pragma unroll
    for (int k = 0; k < R; ++k) {
        for (int t0 = 0; t0 < len[k]; t0 += 32) {
            const int t = t0 + lane;
            const bool valid = t < len[k];
            const int c = valid ? indices[s[k] + t] : -1;
            bool owned = valid;
#pragma unroll
            for (int kk = 0; kk < k; ++kk)
                if (owned && find_col(indices + s[kk], len[kk], c) >= 0) owned = false;
            const unsigned m = __ballot_sync(0xffffffffu, owned);
            if (PASS == 1 && owned) {
                const int64_t slot = base + running + __popc(m & ((1u << lane) - 1u));
                bidx[slot] = inv_order ? inv_order[c] : c;
                double v[R], dv[R];
#pragma unroll
                for (int kk = 0; kk < R; ++kk) { v[kk] = 0.0; dv[kk] = 0.0; }
                v[k] = data[s[k] + t];
                if (ddata) dv[k] = ddata[s[k] + t];
#pragma unroll
                for (int kk = k + 1; kk < R; ++kk) {
                    int pos = find_col(indices + s[kk], len[kk], c);
                    if (pos >= 0) {
                        v[kk] = data[s[kk] + pos];
                        if (ddata) dv[kk] = ddata[s[kk] + pos];
                    }
                }
#pragma unroll
                for (int kk = 0; kk < R; ++kk) bvals[slot * R + kk] = v[kk];
                if (ddata) {
#pragma unroll
                    for (int kk = 0; kk < R; ++kk) bdvals[slot * R + kk] = dv[kk];
                }
            }
            running += __popc(m);
        }
    }
    if (PASS == 0 && lane == 0) nblk[rb] = running;
}

// Y = (K + eta I) X on the row-blocked operator. One warp per row block. The (index, values) stream of the row block
// is moved global -> shared with cp.async in chunks of 32 block-columns (one per lane, fully coalesced, double
// buffered: the next chunk is in flight while the current one is consumed), which keeps enough bytes in flight for
// HBM whatever B is. In the consume loop a lane owns CPL (= 2, or 1 for B = 1) columns of X and every NQ-th
// block-column of the chunk: it reads index + R values from shared memory (broadcast within the lane group) and
// gathers ONE row of X (16-byte loads, served by L1/L2: rows are in Z-order) for R rows of K.
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gmem_src));
}

template <int B, int R>
__global__ void __launch_bounds__(256)
bcsr_spmm_kernel(const int64_t* __restrict__ bptr, const int* __restrict__ bidx, const double* __restrict__ bvals, int n,
                 double eta, const double* __restrict__ X, double* __restrict__ Y) {
    constexpr int CPL = (B >= 2) ? 2 : 1;
    constexpr int LPB = B / CPL;       // lanes per block-column
    constexpr int NQ = 32 / LPB;       // block-columns per warp step
    constexpr int NST = 2;             // cp.async stages
    __shared__ int s_idx[8][NST][32];
    __shared__ __align__(16) double s_val[8][NST][32 * R];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int rb = blockIdx.x * 8 + w;
    if (rb * R >= n) return;           // whole warps leave; only __syncwarp below
    const int q = lane / LPB, c = (lane % LPB) * CPL;
    const int64_t p0 = bptr[rb], p1 = bptr[rb + 1];
    const int nch = (int)((p1 - p0 + 31) >> 5);
    double acc[R][CPL];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int k = 0; k < CPL; ++k) acc[r][k] = 0.0;

    auto issue = [&](int ch) {
        const int st = ch % NST;
        const int64_t p = p0 + (int64_t)ch * 32 + lane;
        if (ch < nch && p < p1) {
            cp_async4(&s_idx[w][st][lane], bidx + p);
#pragma unroll
            for (int r = 0; r < R; r += 2) cp_async16(&s_val[w][st][lane * R + r], bvals + p * R + r);
        }
        cp_async_commit();
    };
    issue(0);
    for (int ch = 0; ch < nch; ++ch) {
        issue(ch + 1);
        cp_async_wait<1>();
        __syncwarp();
        const int st = ch % NST;
        const int cnt = (int)min((int64_t)32, p1 - (p0 + (int64_t)ch * 32));
#pragma unroll
        for (int j = 0; j < 32 / NQ; ++j) {
            const int e = j * NQ + q;
            const bool ok = e < cnt;
            const int col = ok ? s_idx[w][st][e] : 0;
            double v[R], x[CPL];
#pragma unroll
            for (int r = 0; r < R; r += 2) {
                double2 t = *reinterpret_cast<const double2*>(&s_val[w][st][e * R + r]);
                v[r] = ok ? t.x : 0.0;
                v[r + 1] = ok ? t.y : 0.0;
            }
            if (CPL == 2) {
                double2 t = *reinterpret_cast<const double2*>(X + (int64_t)col * B + c);
                x[0] = t.x;
                x[CPL - 1] = t.y;
            } else {
                x[0] = X[(int64_t)col * B + c];
            }
#pragma unroll
            for (int r = 0; r < R; ++r)
#pragma unroll
                for (int k = 0; k < CPL; ++k) acc[r][k] += v[r] * x[k];
        }
        __syncwarp();
    }
#pragma unroll
    for (int o = LPB; o < 32; o <<= 1)
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int k = 0; k < CPL; ++k) acc[r][k] += __shfl_xor_sync(0xffffffffu, acc[r][k], o);
    if (q == 0) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int row = rb * R + r;
            if (row < n) {
#pragma unroll
                for (int k = 0; k < CPL; ++k)
                    Y[(int64_t)row * B + c + k] = acc[r][k] + eta * X[(int64_t)row * B + c + k];
            }
        }
    }
}

// the operator a Krylov routine works on: plain CSR (R = 1) or the row-blocked form (R = 2, 4, 8)
struct SparseOp {
    int R;                 // 1: CSR (ptr32, idx, val); > 1: row-blocked (ptr64, idx, val)
    const int* ptr32;
    const int64_t* ptr64;
    const int* idx;
    const double* val;
    int n;
};

template <int R>
static int bcsr_spmm_launch(const SparseOp& A, double eta, const double* X, int B, double* Y, cudaStream_t s) {
    const int nrb = (A.n + R - 1) / R;
    const int blocks = (nrb + 7) / 8;
    switch (B) {
        case 1: bcsr_spmm_kernel<1, R><<<blocks, 256, 0, s>>>(A.ptr64, A.idx, A.val, A.n, eta, X, Y); break;
        case 2: bcsr_spmm_kernel<2, R><<<blocks, 256, 0, s>>>(A.ptr64, A.idx, A.val, A.n, eta, X, Y); break;
        case 4: bcsr_spmm_kernel<4, R><<<blocks, 256, 0, s>>>(A.ptr64, A.idx, A.val, A.n, eta, X, Y); break;
        case 8: bcsr_spmm_kernel<8, R><<<blocks, 256, 0, s>>>(A.ptr64, A.idx, A.val, A.n, eta, X, Y); break;
        case 16: bcsr_spmm_kernel<16, R><<<blocks, 256, 0, s>>>(A.ptr64, A.idx, A.val, A.n, eta, X, Y); break;
        case 32: bcsr_spmm_kernel<32, R><<<blocks, 256, 0, s>>>(A.ptr64, A.idx, A.val, A.n, eta, X, Y); break;
        default: return -2;
    }
    return 0;
}

static int spmm(const SparseOp& A, double eta, const double* X, int B, double* Y, cudaStream_t s) {
    const int n = A.n;
    int rc = 0;
    if (A.R == 1) {
        int blocks = (int)(((int64_t)n * 32 + 255) / 256);
        switch (B) {
            case 1: csr_spmm_kernel<1><<<blocks, 256, 0, s>>>(A.ptr32, A.idx, A.val, n, eta, X, Y); break;
            case 2: csr_spmm_kernel<2><<<blocks, 256, 0, s>>>(A.ptr32, A.idx, A.val, n, eta, X, Y); break;
            case 4: csr_spmm_kernel<4><<<blocks, 256, 0, s>>>(A.ptr32, A.idx, A.val, n, eta, X, Y); break;
            case 8: csr_spmm_kernel<8><<<blocks, 256, 0, s>>>(A.ptr32, A.idx, A.val, n, eta, X, Y); break;
            case 16: csr_spmm_kernel<16><<<blocks, 256, 0, s>>>(A.ptr32, A.idx, A.val, n, eta, X, Y); break;
            case 32: csr_spmm_kernel<32><<<blocks, 256, 0, s>>>(A.ptr32, A.idx, A.val, n, eta, X, Y); break;
            default: return -2;
        }
    } else if (A.R == 2) {
        rc = bcsr_spmm_launch<2>(A, eta, X, B, Y, s);
    } else if (A.R == 4) {
        rc = bcsr_spmm_launch<4>(A, eta, X, B, Y, s);
    } else if (A.R == 8) {
        rc = bcsr_spmm_launch<8>(A, eta, X, B, Y, s);
    } else {
        return -3;
    }
    if (rc) return rc;
    GP_COUNT(1);
    GP_LAUNCH_CHECK();
    return 0;
}

static SparseOp csr_op(const int* indptr, const int* indices, const double* data, int64_t n) {
    SparseOp A = {1, indptr, nullptr, indices, data, (int)n};
    return A;
}
static SparseOp bcsr_op(int64_t R, const int64_t* bptr, const int* bidx, const double* bvals, int64_t n) {
    SparseOp A = {(int)R, nullptr, bptr, bidx, bvals, (int)n};
    return A;
}

static size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace gp

using namespace gp;

extern "C" {

int gp_csr_spmm(const int* indptr, const int* indices, const double* data, int64_t n, double eta, const double* X, int64_t B,
                double* Y, void* stream) {
    if (!indptr || !indices || !data || !X || !Y || n <= 0 || n > INT32_MAX) return -1;
    return spmm(csr_op(indptr, indices, data, n), eta, X, (int)B, Y, (cudaStream_t)stream);
}

int gp_bcsr_spmm(int64_t R, const int64_t* bptr, const int* bidx, const double* bvals, int64_t n, double eta, const double* X,
                 int64_t B, double* Y, void* stream) {
    if (!bptr || !bidx || !bvals || !X || !Y || n <= 0 || n > INT32_MAX) return -1;
    return spmm(bcsr_op(R, bptr, bidx, bvals, n), eta, X, (int)B, Y, (cudaStream_t)stream);
}

}  // extern "C"

template <int R>
static int bcsr_build(int pass, int n, const int* order, const int* inv_order, const int* indptr, const int* indices,
                      const double* data, const double* ddata, int* nblk, const int64_t* bptr, int* bidx, double* bvals,
                      double* bdvals, cudaStream_t s) {
    const int nrb = (n + R - 1) / R;
    const unsigned blocks = (unsigned)(((int64_t)nrb * 32 + 255) / 256);
    if (pass == 0)
        bcsr_build_kernel<R, 0><<<blocks, 256, 0, s>>>(n, order, inv_order, indptr, indices, data, ddata, nblk, bptr, bidx, bvals, bdvals);
    else
        bcsr_build_kernel<R, 1><<<blocks, 256, 0, s>>>(n, order, inv_order, indptr, indices, data, ddata, nblk, bptr, bidx, bvals, bdvals);
    GP_COUNT(1);
    GP_LAUNCH_CHECK();
    return 0;
}

static int bcsr_build_any(int64_t R, int pass, int64_t n, const int* order, const int* inv_order, const int* indptr,
                          const int* indices, const double* data, const double* ddata, int* nblk, const int64_t* bptr,
                          int* bidx, double* bvals, double* bdvals, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (!indptr || !indices || n <= 0 || n > INT32_MAX || ((order == nullptr) != (inv_order == nullptr))) return -1;
    switch (R) {
        case 2: return bcsr_build<2>(pass, (int)n, order, inv_order, indptr, indices, data, ddata, nblk, bptr, bidx, bvals, bdvals, s);
        case 4: return bcsr_build<4>(pass, (int)n, order, inv_order, indptr, indices, data, ddata, nblk, bptr, bidx, bvals, bdvals, s);
        case 8: return bcsr_build<8>(pass, (int)n, order, inv_order, indptr, indices, data, ddata, nblk, bptr, bidx, bvals, bdvals, s);
        default: return -3;
    }
}

extern "C" {

int gp_bcsr_count(int64_t R, int64_t n, const int* order, const int* inv_order, const int* indptr, const int* indices,
                  int* nblk, void* stream) {
    if (!nblk) return -1;
    return bcsr_build_any(R, 0, n, order, inv_order, indptr, indices, nullptr, nullptr, nblk, nullptr, nullptr, nullptr, nullptr,
                          stream);
}

int gp_bcsr_fill(int64_t R, int64_t n, const int* order, const int* inv_order, const int* indptr, const int* indices,
                 const double* data, const double* ddata, const int64_t* bptr, int* bidx, double* bvals, double* bdvals,
                 void* stream) {
    if (!data || !bptr || !bidx || !bvals || (ddata && !bdvals)) return -1;
    return bcsr_build_any(R, 1, n, order, inv_order, indptr, indices, data, ddata, nullptr, bptr, bidx, bvals, bdvals, stream);
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

}  // extern "C"

// m Lanczos steps of A = K + eta I started from the columns of V (normalised internally).
// alpha_dev, beta_dev: (m x B) row-major; beta[j] is the norm of the (j+1)-th unnormalised vector.
static int lanczos_run(const SparseOp& A, double eta, const double* V, int64_t B, int64_t m, double* alpha_dev,
                       double* beta_dev, void* ws, void* stream) {
    const int64_t n = A.n;
    if (!A.idx || !A.val || !V || !alpha_dev || !beta_dev || !ws || n <= 0 || m <= 0) return -1;
    if (B <= 0 || B > 32 || (32 % B)) return -2;
    cudaStream_t s = (cudaStream_t)stream;
    const int Bc = (int)B;
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
        int rc = spmm(A, eta, q, Bc, W, s);
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
static int cg_run(const SparseOp& A, double eta, double* R0, double* X, int64_t B, double tol, int64_t maxiter,
                  int64_t* iters_host, void* ws, void* stream) {
    const int64_t n = A.n;
    if (!A.idx || !A.val || !R0 || !X || !ws || n <= 0) return -1;
    if (B <= 0 || B > 32 || (32 % B)) return -2;
    cudaStream_t s = (cudaStream_t)stream;
    const int Bc = (int)B;
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
        int rc = spmm(A, eta, Pd, Bc, AP, s);
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

extern "C" {

int gp_lanczos(const int* indptr, const int* indices, const double* data, int64_t n, double eta, const double* V, int64_t B,
               int64_t m, double* alpha_dev, double* beta_dev, void* ws, void* stream) {
    if (!indptr || n <= 0 || n > INT32_MAX) return -1;
    return lanczos_run(csr_op(indptr, indices, data, n), eta, V, B, m, alpha_dev, beta_dev, ws, stream);
}

int gp_bcsr_lanczos(int64_t R, const int64_t* bptr, const int* bidx, const double* bvals, int64_t n, double eta,
                    const double* V, int64_t B, int64_t m, double* alpha_dev, double* beta_dev, void* ws, void* stream) {
    if (!bptr || n <= 0 || n > INT32_MAX || (R != 2 && R != 4 && R != 8)) return -1;
    return lanczos_run(bcsr_op(R, bptr, bidx, bvals, n), eta, V, B, m, alpha_dev, beta_dev, ws, stream);
}

int gp_cg_solve(const int* indptr, const int* indices, const double* data, int64_t n, double eta, double* R0, double* X,
                int64_t B, double tol, int64_t maxiter, int64_t* iters_host, void* ws, void* stream) {
    if (!indptr || n <= 0 || n > INT32_MAX) return -1;
    return cg_run(csr_op(indptr, indices, data, n), eta, R0, X, B, tol, maxiter, iters_host, ws, stream);
}

int gp_bcsr_cg_solve(int64_t R, const int64_t* bptr, const int* bidx, const double* bvals, int64_t n, double eta, double* R0,
                     double* X, int64_t B, double tol, int64_t maxiter, int64_t* iters_host, void* ws, void* stream) {
    if (!bptr || n <= 0 || n > INT32_MAX || (R != 2 && R != 4 && R != 8)) return -1;
    return cg_run(bcsr_op(R, bptr, bidx, bvals, n), eta, R0, X, B, tol, maxiter, iters_host, ws, stream);
}

}  // extern "C"
