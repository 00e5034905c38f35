// Developer experiment: DMMA GEMM efficiency on the exact shapes potrf / trtri / lauum use.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include "../include/gpgp.h"
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
static float run(int at, int bt, double* C, long ldc, const double* A, long lda, const double* B, long ldb, long M, long N, long K, double al, double be, int kr, int tm, int reps) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    gp_dgemm_f64(at, bt, C, ldc, A, lda, B, ldb, M, N, K, al, be, kr, tm, 0); CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    for (int r = 0; r < reps; ++r) gp_dgemm_f64(at, bt, C, ldc, A, lda, B, ldb, M, N, K, al, be, kr, tm, 0);
    cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms / reps;
}
int main() {
    const long n = 19072;  // 149 tiles
    double *C, *A; CK(cudaMalloc(&C, n * n * 8)); CK(cudaMalloc(&A, n * n * 8));
    CK(cudaMemset(C, 0, n * n * 8)); CK(cudaMemset(A, 0, n * n * 8));
    long T = n / 128;
    double full = 2.0 * n * n * 512, low = 2.0 * 128 * 128 * 512 * (T * (T + 1) / 2);
    float t;
    t = run(0, 0, C, n, A, n, A, n, n, n, 512, 1, 0, 0, 0, 3); printf("full  K=512 beta=0      : %.3f ms %.2f TF\n", t, full / t * 1e-9);
    t = run(0, 0, C, n, A, n, A, n, n, n, 512, -1, 1, 0, 0, 3); printf("full  K=512 beta=1      : %.3f ms %.2f TF\n", t, full / t * 1e-9);
    t = run(0, 0, C, n, A, n, A, n, n, n, 512, -1, 1, 0, 1, 3); printf("lower K=512 beta=1      : %.3f ms %.2f TF\n", t, low / t * 1e-9);
    t = run(0, 0, C, n, A, n, A, n, n, n, 1024, -1, 1, 0, 1, 3); printf("lower K=1024 beta=1     : %.3f ms %.2f TF\n", t, 2 * low / t * 1e-9);
    t = run(0, 0, C, n, A, n, A, n, n, n, 2048, -1, 1, 0, 1, 3); printf("lower K=2048 beta=1     : %.3f ms %.2f TF\n", t, 4 * low / t * 1e-9);
    t = run(0, 0, C, n, A, n, A, n, n, 128, 128, 1, 0, 0, 0, 10); printf("trsm  N=128 K=128       : %.3f ms %.2f TF\n", t, 2.0 * n * 128 * 128 / t * 1e-9);
    t = run(0, 0, C, n, A, n, A, n, n, 384, 128, -1, 1, 0, 1, 10); printf("inner N=384 K=128 lower : %.3f ms\n", t);
    // lauum-like and trtri-like
    double tn = 0; for (long tm = 0; tm < T; ++tm) tn += (double)(tm + 1) * (n - tm * 128);
    t = run(1, 1, C, n, A, n, A, n, n, n, n, 1, 0, 3, 1, 1); printf("lauum TN lower          : %.3f ms %.2f TF(tile)\n", t, 2.0 * 128 * 128 * tn / t * 1e-9);
    long h = n / 2 / 128 * 128;
    t = run(0, 1, C, h, A, n, A, n, h, h, h, 1, 0, 2, 0, 1); printf("trtri NN B-lower half   : %.3f ms %.2f TF(alg)\n", t, 1.0 * h * h * h / t * 1e-9);
    t = run(0, 1, C, n, A, n, A, h, h, h, h, -1, 0, 1, 0, 1); printf("trtri NN A-lower half   : %.3f ms %.2f TF(alg)\n", t, 1.0 * h * h * h / t * 1e-9);
    return 0;
}
