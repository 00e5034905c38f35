// Round-1 microbenchmarks that fix the FP64 roofline denominators on the B200 in use:
//   (1) raw DMMA.8x8x4 issue rate, (2) raw DFMA rate, (3) cuBLAS DGEMM (the FP64 "library peak"),
//   (4) gpgp's DMMA GEMM kernel on the shapes the Cholesky uses, checked against cuBLAS.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu -L<pkg> -lgpgp -lcublas
#include <cublas_v2.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "../include/gpgp.h"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__global__ void dmma_rate(double* out, int iters) {
    double acc[16][2];
    for (int i = 0; i < 16; ++i) acc[i][0] = acc[i][1] = 0.0;
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(acc[i][0]), "+d"(acc[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
    for (int i = 0; i < 16; ++i) s += acc[i][0] + acc[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void dfma_rate(double* out, int iters) {
    double acc[16];
    for (int i = 0; i < 16; ++i) acc[i] = i;
    double a = 1.0 + threadIdx.x * 1e-9, b = threadIdx.x * 1e-4;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
    for (int i = 0; i < 16; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

static float time_ms(cudaEvent_t a, cudaEvent_t b) { float ms; cudaEventElapsedTime(&ms, a, b); return ms; }

int main(int argc, char** argv) {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", prop.name, sms, prop.clockRate);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 1024));

    for (int warps = 4; warps <= 32; warps *= 2) {
        int iters = 20000;
        dmma_rate<<<sms, warps * 32>>>(out, 100); CK(cudaDeviceSynchronize());
        cudaEventRecord(e0); dmma_rate<<<sms, warps * 32>>>(out, iters); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
        double fl = (double)sms * warps * iters * 16 * 512.0;
        printf("{\"bench\": \"dmma_rate\", \"warps_per_sm\": %d, \"tflops\": %.2f}\n", warps, fl / time_ms(e0, e1) * 1e-9);
        dfma_rate<<<sms, warps * 32>>>(out, 100); CK(cudaDeviceSynchronize());
        cudaEventRecord(e0); dfma_rate<<<sms, warps * 32>>>(out, iters); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
        fl = (double)sms * warps * 32 * iters * 16 * 2.0;
        printf("{\"bench\": \"dfma_rate\", \"warps_per_sm\": %d, \"tflops\": %.2f}\n", warps, fl / time_ms(e0, e1) * 1e-9);
    }

    // ---- cuBLAS DGEMM and gpgp DGEMM --------------------------------------------------------
    cublasHandle_t h; cublasCreate(&h);
    const int NMAX = 8192;
    size_t bytes = sizeof(double) * NMAX * NMAX;
    double *A, *B, *C, *Cref; CK(cudaMalloc(&A, bytes)); CK(cudaMalloc(&B, bytes)); CK(cudaMalloc(&C, bytes)); CK(cudaMalloc(&Cref, bytes));
    std::vector<double> hA((size_t)NMAX * NMAX), hB((size_t)NMAX * NMAX);
    srand(1);
    for (size_t i = 0; i < hA.size(); ++i) { hA[i] = rand() / (double)RAND_MAX - 0.5; hB[i] = rand() / (double)RAND_MAX - 0.5; }
    CK(cudaMemcpy(A, hA.data(), bytes, cudaMemcpyHostToDevice)); CK(cudaMemcpy(B, hB.data(), bytes, cudaMemcpyHostToDevice));
    double one = 1.0, zero = 0.0;

    struct Shape { int M, N, K; };
    Shape shapes[] = {{8192, 8192, 8192}, {8192, 8192, 512}, {8192, 8192, 128}, {8192, 128, 128}, {4096, 4096, 4096}};
    for (auto s : shapes) {
        // row-major C = A * B^T  (NT)  == column-major C^T = B * A^T -> cublas(T, N) on swapped operands
        auto cublas_nt = [&]() { cublasDgemm(h, CUBLAS_OP_T, CUBLAS_OP_N, s.N, s.M, s.K, &one, B, NMAX, A, NMAX, &zero, Cref, NMAX); };
        cublas_nt(); CK(cudaDeviceSynchronize());
        cudaEventRecord(e0); for (int r = 0; r < 3; ++r) cublas_nt(); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
        double fl = 2.0 * s.M * s.N * s.K * 3;
        printf("{\"bench\": \"cublas_dgemm_nt\", \"M\": %d, \"N\": %d, \"K\": %d, \"tflops\": %.2f}\n", s.M, s.N, s.K, fl / time_ms(e0, e1) * 1e-9);
        int rc = gp_dgemm_f64(0, 0, C, NMAX, A, NMAX, B, NMAX, s.M, s.N, s.K, 1.0, 0.0, 0, 0, 0);
        CK(cudaDeviceSynchronize());
        cudaEventRecord(e0); for (int r = 0; r < 3; ++r) gp_dgemm_f64(0, 0, C, NMAX, A, NMAX, B, NMAX, s.M, s.N, s.K, 1.0, 0.0, 0, 0, 0); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
        printf("{\"bench\": \"gpgp_dgemm_nt\", \"rc\": %d, \"M\": %d, \"N\": %d, \"K\": %d, \"tflops\": %.2f}\n", rc, s.M, s.N, s.K, fl / time_ms(e0, e1) * 1e-9);
    }

    // ---- correctness of all operand layouts vs cuBLAS at 1024 x 768 x 512 ----------------------------
    {
        int M = 1024, N = 768, K = 512;
        std::vector<double> c1((size_t)M * NMAX), c2((size_t)M * NMAX);
        for (int at = 0; at < 2; ++at) for (int bt = 0; bt < 2; ++bt) {
            // reference through cuBLAS in column-major terms: C^T(NxM) = opB' * opA'
            cublasOperation_t ta = at == 0 ? CUBLAS_OP_T : CUBLAS_OP_N;  // A(m,k): at==0 stored [m][k] -> col-major (k x m) -> A^T needs OP_T to be (m x k)... see below
            // column-major view: stored A[m][k] (row-major, ld) == col-major matrix Ac of shape (k x m). C^T = Bop * Aop with
            //   Aop (K x M) = Ac if at==0 else Ac^T (stored [k][m] == col-major (m x k))
            //   Bop (N x K) = Bc^T if bt==0 (stored [n][k] == col-major (k x n)) else Bc (stored [k][n] == col-major (n x k))
            cublasOperation_t opB = bt == 0 ? CUBLAS_OP_T : CUBLAS_OP_N;
            cublasOperation_t opA = at == 0 ? CUBLAS_OP_N : CUBLAS_OP_T;
            (void)ta;
            cublasDgemm(h, opB, opA, N, M, K, &one, B, NMAX, A, NMAX, &zero, Cref, NMAX);
            int rc = gp_dgemm_f64(at, bt, C, NMAX, A, NMAX, B, NMAX, M, N, K, 1.0, 0.0, 0, 0, 0);
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(c1.data(), C, sizeof(double) * M * NMAX, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(c2.data(), Cref, sizeof(double) * M * NMAX, cudaMemcpyDeviceToHost));
            double md = 0, mx = 0;
            for (int i = 0; i < M; ++i) for (int j = 0; j < N; ++j) { md = fmax(md, fabs(c1[(size_t)i * NMAX + j] - c2[(size_t)i * NMAX + j])); mx = fmax(mx, fabs(c2[(size_t)i * NMAX + j])); }
            printf("{\"check\": \"gemm_vs_cublas\", \"at\": %d, \"bt\": %d, \"rc\": %d, \"max_abs_diff\": %.3e, \"max_abs\": %.3e}\n", at, bt, rc, md, mx);
        }
    }
    return 0;
}
