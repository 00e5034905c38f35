// Developer check: the DMMA GEMM launched concurrently on two streams (and next to a kernel with an odd shared-memory
// footprint) must return exactly what it returns alone.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "../include/gpgp.h"
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__global__ void smem_hog(double* out, int iters) {
    extern __shared__ double sh[];
    sh[threadIdx.x] = threadIdx.x;
    __syncthreads();
    double a = 0;
    for (int i = 0; i < iters; ++i) a += sh[(threadIdx.x + i) & 127] * 1e-9;
    if (a == 12345.678) out[0] = a;
}

static double maxdiff(const double* a, const double* b, size_t n) {
    std::vector<double> ha(n), hb(n);
    CK(cudaMemcpy(ha.data(), a, n * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hb.data(), b, n * 8, cudaMemcpyDeviceToHost));
    double m = 0;
    for (size_t i = 0; i < n; ++i) { double d = fabs(ha[i] - hb[i]); if (!(d <= m)) m = d; }
    return m;
}

int main() {
    const long n = 4096;
    size_t bytes = n * n * 8;
    double *A, *B, *C1, *C2, *R1, *R2, *dummy;
    CK(cudaMalloc(&A, bytes)); CK(cudaMalloc(&B, bytes)); CK(cudaMalloc(&C1, bytes)); CK(cudaMalloc(&C2, bytes));
    CK(cudaMalloc(&R1, bytes)); CK(cudaMalloc(&R2, bytes)); CK(cudaMalloc(&dummy, 1024));
    std::vector<double> h(n * n);
    srand(3);
    for (auto& v : h) v = rand() / (double)RAND_MAX - 0.5;
    CK(cudaMemcpy(A, h.data(), bytes, cudaMemcpyHostToDevice));
    for (auto& v : h) v = rand() / (double)RAND_MAX - 0.5;
    CK(cudaMemcpy(B, h.data(), bytes, cudaMemcpyHostToDevice));
    cudaStream_t s1, s2, s3;
    CK(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&s3, cudaStreamNonBlocking));
    int combos[4][2] = {{0, 0}, {0, 1}, {1, 1}, {1, 0}};
    for (int c1 = 0; c1 < 4; ++c1) for (int c2 = 0; c2 < 4; ++c2) {
        // references, one at a time
        gp_dgemm_f64(combos[c1][0], combos[c1][1], R1, n, A, n, B, n, n, n, n, 1.0, 0.0, 0, 0, s1); CK(cudaDeviceSynchronize());
        gp_dgemm_f64(combos[c2][0], combos[c2][1], R2, n, B, n, A, n, n, n, n, 1.0, 0.0, 0, 0, s1); CK(cudaDeviceSynchronize());
        double worst1 = 0, worst2 = 0;
        for (int rep = 0; rep < 3; ++rep) {
            CK(cudaMemsetAsync(C1, 0, bytes, s1)); CK(cudaMemsetAsync(C2, 0, bytes, s2));
            CK(cudaDeviceSynchronize());
            smem_hog<<<148, 128, (size_t)(13 * 1024 + 128 * (rep + 1)), s3>>>(dummy, 200000);
            gp_dgemm_f64(combos[c1][0], combos[c1][1], C1, n, A, n, B, n, n, n, n, 1.0, 0.0, 0, 0, s1);
            gp_dgemm_f64(combos[c2][0], combos[c2][1], C2, n, B, n, A, n, n, n, n, 1.0, 0.0, 0, 0, s2);
            CK(cudaDeviceSynchronize());
            worst1 = fmax(worst1, maxdiff(C1, R1, n * n)); worst2 = fmax(worst2, maxdiff(C2, R2, n * n));
        }
        printf("concurrent (%d,%d) || (%d,%d): max |diff| vs alone = %.3e , %.3e\n", combos[c1][0], combos[c1][1], combos[c2][0], combos[c2][1], worst1, worst2);
    }
    return 0;
}
