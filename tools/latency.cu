// Single-warp dependent-chain latencies on the FP64 path (cycles per op), to size the serial diagonal-block kernel.
#include <cuda_runtime.h>
#include <stdio.h>
__global__ void lat(double* out, long long* clk, double x0) {
    __shared__ double sm[64];
    int lane = threadIdx.x;
    double x = x0 + lane * 1e-9, y = 1.0000001;
    long long t0, t1;
    // DFMA chain
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 256; ++i) x = fma(x, y, 1e-9);
    t1 = clock64(); if (lane == 0) clk[0] = t1 - t0;
    // DMUL chain
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 256; ++i) x = x * y;
    t1 = clock64(); if (lane == 0) clk[1] = t1 - t0;
    // rsqrt chain
    x = fabs(x) + 1.0;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i) x = rsqrt(x) + 1.0;
    t1 = clock64(); if (lane == 0) clk[2] = t1 - t0;
    // shfl chain (double)
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i) x = __shfl_sync(0xffffffffu, x, (i * 7) & 31) + 1e-9;
    t1 = clock64(); if (lane == 0) clk[3] = t1 - t0;
    // STS -> syncwarp -> LDS chain
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i) { sm[lane] = x; __syncwarp(); x = sm[(lane + 1) & 31] + 1e-9; __syncwarp(); }
    t1 = clock64(); if (lane == 0) clk[4] = t1 - t0;
    // sqrt and div chains
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i) x = sqrt(x) + 1.0;
    t1 = clock64(); if (lane == 0) clk[5] = t1 - t0;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i) x = 1.0 / x + 1.0;
    t1 = clock64(); if (lane == 0) clk[6] = t1 - t0;
    // 32 independent DFMAs issued back to back (throughput of one warp)
    double a[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) a[i] = x + i;
    t0 = clock64();
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int i = 0; i < 32; ++i) a[i] = fma(a[i], y, x);
    t1 = clock64(); if (lane == 0) clk[7] = t1 - t0;
    double s = x;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += a[i];
    out[lane] = s;
}
int main() {
    double* out; long long* clk; cudaMalloc(&out, 256); cudaMalloc(&clk, 64 * 8);
    for (int rep = 0; rep < 2; ++rep) { lat<<<1, 32>>>(out, clk, 1.0); cudaDeviceSynchronize(); }
    long long h[8]; cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
    printf("{\"dfma_dep\": %.1f, \"dmul_dep\": %.1f, \"rsqrt_plus_add_dep\": %.1f, \"shfl64_plus_add_dep\": %.1f, \"sts_sync_lds_add_dep\": %.1f, \"sqrt_plus_add_dep\": %.1f, \"div_plus_add_dep\": %.1f, \"dfma_indep_per_op\": %.2f}\n",
           h[0] / 256.0, h[1] / 256.0, h[2] / 64.0, h[3] / 64.0, h[4] / 64.0, h[5] / 64.0, h[6] / 64.0, h[7] / 256.0);
    return 0;
}
