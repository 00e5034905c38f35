// Internal (non-ABI) declarations shared between the gpgp translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gp {

// k-range restriction of one 128x128 output tile at (m0, n0); exploits triangular operands
enum KRange {
    KR_FULL = 0,      // k in [0, K)
    KR_A_LOWER = 1,   // A(m,k) lower triangular: k in [0, m0 + 128)
    KR_B_LOWER = 2,   // B(k,n) lower triangular: k in [n0, K)
    KR_TN_LOWER = 3,  // A(m,k)=W[k][m], B(k,n)=W[k][n], W lower: k in [max(m0,n0), K)
};
enum TileMask {
    TM_ALL = 0,
    TM_LOWER = 1,  // only tiles with n0 <= m0; diagonal tiles store col <= row only
};

// C[M x N] (row-major, ldc) = beta*C + alpha * op(A) * op(B)
//   at == 0: A(m,k) = A[m*lda + k]   at == 1: A(m,k) = A[k*lda + m]
//   bt == 0: B(k,n) = B[n*ldb + k]   bt == 1: B(k,n) = B[k*ldb + n]
// M, N multiples of 128; K multiple of 32; all ld even; pointers 16-byte aligned.
int launch_dgemm(int at, int bt, double* C, int64_t ldc, const double* A, int64_t lda,
                 const double* B, int64_t ldb, int M, int N, int K, double alpha, double beta,
                 int krange, int tmask, cudaStream_t stream);

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (device, kernel); thread-safe
int configure_once(const void* func, int smem_bytes);

int launch_dgemm_ktab(int at, int bt, double* C, int64_t ldc, const double* A, int64_t lda, const double* B, int64_t ldb,
                      int M, int N, int K, double alpha, double beta, const int* kbeg_tab, const int* kend_tab,
                      cudaStream_t stream);

int set_gemm_impl(int impl);   // 1 = TMA + mbarrier kernel (default), 0 = cp.async kernel (A-B measurements)
int profile_enable(int on);
int profile_read(double* ms_sum, double* ms_union, double* flops, long long* launches);

}  // namespace gp
