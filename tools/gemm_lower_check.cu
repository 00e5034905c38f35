// Developer check: the trailing update C -= P P^T (lower tiles, beta = 1) with the TMA kernel against the cp.async kernel.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "../include/gpgp.h"
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
__global__ void fill(double* p, size_t n, unsigned seed) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) { unsigned x = (unsigned)(i * 2654435761u) ^ seed; x ^= x >> 13; x *= 0x5bd1e995; x ^= x >> 15; p[i] = (x & 0xffff) / 65536.0 - 0.5; }
}
__global__ void diffk(const double* a, const double* b, size_t n, double tol, unsigned long long* cnt, double* mx) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) { double d = fabs(a[i] - b[i]); if (!(d <= tol)) { atomicAdd(cnt, 1ull); } if (d > *mx) *mx = d; }
}
int main(int argc, char** argv) {
    const long n = 19072, K = argc > 1 ? atol(argv[1]) : 512;
    int tmask = argc > 2 ? atoi(argv[2]) : 1;
    double beta = argc > 3 ? atof(argv[3]) : 1.0;
    double alpha = argc > 4 ? atof(argv[4]) : -1.0;
    double *C, *C0, *Cref, *P, *mx; unsigned long long* cnt;
    CK(cudaMalloc(&C, n * n * 8)); CK(cudaMalloc(&C0, n * n * 8)); CK(cudaMalloc(&Cref, n * n * 8)); CK(cudaMalloc(&P, n * K * 8));
    CK(cudaMalloc(&mx, 8)); CK(cudaMalloc(&cnt, 8));
    fill<<<(unsigned)((n * n + 255) / 256), 256>>>(C0, n * n, 1);
    fill<<<(unsigned)((n * K + 255) / 256), 256>>>(P, n * K, 7);
    CK(cudaDeviceSynchronize());
    gp_gemm_set_impl(0);
    CK(cudaMemcpy(Cref, C0, n * n * 8, cudaMemcpyDeviceToDevice));
    int rc = gp_dgemm_f64(0, 0, Cref, n, P, K, P, K, n, n, K, alpha, beta, 0, tmask, 0);
    CK(cudaDeviceSynchronize());
    gp_gemm_set_impl(1);
    for (int rep = 0; rep < 5; ++rep) {
        CK(cudaMemcpy(C, C0, n * n * 8, cudaMemcpyDeviceToDevice));
        rc |= gp_dgemm_f64(0, 0, C, n, P, K, P, K, n, n, K, alpha, beta, 0, tmask, 0);
        CK(cudaMemset(cnt, 0, 8)); CK(cudaMemset(mx, 0, 8));
        diffk<<<(unsigned)((n * n + 255) / 256), 256>>>(C, Cref, n * n, 1e-9, cnt, mx);
        unsigned long long h; double hm; CK(cudaMemcpy(&h, cnt, 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&hm, mx, 8, cudaMemcpyDeviceToHost));
        printf("K=%ld tmask=%d beta=%g rep %d rc=%d: entries off by > 1e-9: %llu (max |diff| ~ %.3e)\n", K, tmask, beta, rep, rc, h, hm);
    }
    return 0;
}
