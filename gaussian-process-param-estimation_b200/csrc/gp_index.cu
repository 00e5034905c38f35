// Index plumbing of the sparse operator build on own kernels (no library sort / scan / gather): bounding box of the points,
// stable LSD radix sort of the space-filling-curve keys (-> the deterministic spatial order of the row-blocked operator),
// inverse permutation, exclusive scan of the block counts, row gathers between original and operator order.
// (Round 1 used torch.sort / cumsum / amin / amax / fancy indexing here.)
#include "../../include/gpgp.h"
#include "gp_common.cuh"
#include "gp_internal.h"
#include <float.h>

namespace gp {

// ---- bounding box -----------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
minmax_partial_kernel(const double* __restrict__ pts, int64_t n, int d, double* __restrict__ part) {
    __shared__ double slo[256], shi[256];
    for (int k = 0; k < d; ++k) {
        double lo = DBL_MAX, hi = -DBL_MAX;
        for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
            const double v = pts[i * d + k];
            lo = fmin(lo, v);
            hi = fmax(hi, v);
        }
        slo[threadIdx.x] = lo;
        shi[threadIdx.x] = hi;
        __syncthreads();
        for (int s = 128; s > 0; s >>= 1) {
            if (threadIdx.x < s) {
                slo[threadIdx.x] = fmin(slo[threadIdx.x], slo[threadIdx.x + s]);
                shi[threadIdx.x] = fmax(shi[threadIdx.x], shi[threadIdx.x + s]);
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            part[((int64_t)blockIdx.x * d + k) * 2] = slo[0];
            part[((int64_t)blockIdx.x * d + k) * 2 + 1] = shi[0];
        }
        __syncthreads();
    }
}

__global__ void minmax_final_kernel(const double* __restrict__ part, int nparts, int d, double* __restrict__ out) {
    const int k = threadIdx.x;
    if (k >= d) return;
    double lo = DBL_MAX, hi = -DBL_MAX;
    for (int b = 0; b < nparts; ++b) {
        lo = fmin(lo, part[((int64_t)b * d + k) * 2]);
        hi = fmax(hi, part[((int64_t)b * d + k) * 2 + 1]);
    }
    out[k] = lo;
    out[d + k] = hi;
}

// ---- stable LSD radix sort of (64-bit key, 32-bit value), 8 bits per pass, one warp per tile of RTILE elements --------------
constexpr int RTILE = 1024;
constexpr int RWARPS = 4;

__global__ void __launch_bounds__(RWARPS * 32)
radix_hist_kernel(const unsigned long long* __restrict__ keys, int64_t n, int shift, int ntiles, unsigned* __restrict__ hist) {
    __shared__ unsigned cnt[RWARPS][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tile = blockIdx.x * RWARPS + warp;
    for (int i = lane; i < 256; i += 32) cnt[warp][i] = 0;
    __syncwarp();
    if (tile < ntiles) {
        const int64_t i0 = (int64_t)tile * RTILE;
        for (int off = 0; off < RTILE; off += 32) {
            const int64_t i = i0 + off + lane;
            const bool ok = i < n;
            const unsigned dg = ok ? (unsigned)((keys[i] >> shift) & 255ull) : 256u + lane;    // inactive lanes: unique dummy digits
            const unsigned peers = __match_any_sync(0xffffffffu, dg);
            if (ok && (peers & ((1u << lane) - 1u)) == 0) cnt[warp][dg] += __popc(peers);
            __syncwarp();
        }
        for (int i = lane; i < 256; i += 32) hist[(int64_t)i * ntiles + tile] = cnt[warp][i];
    }
}

// Exclusive scan by one CTA of 32 warps, coalesced: warp w owns a contiguous segment; pass 1 sums the segments (lane-strided
// loads + shuffle reduction), thread 0 scans the 32 segment sums, pass 2 re-reads each segment 32 elements at a time with a
// warp-inclusive scan. Fixed order: bit-reproducible. OUT may alias IN element-wise (same index read before written).
template <typename TIN, typename TOUT, bool SHIFTED>
__device__ __forceinline__ void cta_exclusive_scan(const TIN* in, int64_t total, TOUT* out) {
    __shared__ unsigned long long seg[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t per = ((total + 31) / 32 + 31) / 32 * 32;       // segment length, a multiple of 32
    const int64_t b = (int64_t)warp * per, e = (b + per < total) ? b + per : total;
    unsigned long long s = 0;
    for (int64_t i = b + lane; i < e; i += 32) s += (unsigned long long)in[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) seg[warp] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long run = 0;
        for (int i = 0; i < 32; ++i) { unsigned long long v = seg[i]; seg[i] = run; run += v; }
        if (SHIFTED) out[0] = (TOUT)0;
    }
    __syncthreads();
    unsigned long long run = seg[warp];
    for (int64_t i0 = b; i0 < e; i0 += 32) {
        const int64_t i = i0 + lane;
        const unsigned long long v = (i < e) ? (unsigned long long)in[i] : 0ull;
        unsigned long long inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (i < e) {
            if (SHIFTED) out[i + 1] = (TOUT)(run + inc);          // offsets[i + 1] = inclusive sum
            else out[i] = (TOUT)(run + inc - v);                   // exclusive, in place
        }
        run += __shfl_sync(0xffffffffu, inc, 31);
    }
}

__global__ void __launch_bounds__(1024)
scan_u32_kernel(unsigned* __restrict__ a, int64_t total) { cta_exclusive_scan<unsigned, unsigned, false>(a, total, a); }

__global__ void __launch_bounds__(RWARPS * 32)
radix_scatter_kernel(const unsigned long long* __restrict__ keys, const int* __restrict__ vals, int64_t n, int shift, int ntiles,
                     const unsigned* __restrict__ offs, unsigned long long* __restrict__ keys_out, int* __restrict__ vals_out) {
    __shared__ unsigned base[RWARPS][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tile = blockIdx.x * RWARPS + warp;
    if (tile >= ntiles) return;
    for (int i = lane; i < 256; i += 32) base[warp][i] = offs[(int64_t)i * ntiles + tile];
    __syncwarp();
    const int64_t i0 = (int64_t)tile * RTILE;
    for (int off = 0; off < RTILE; off += 32) {
        const int64_t i = i0 + off + lane;
        const bool ok = i < n;
        const unsigned long long key = ok ? keys[i] : 0ull;
        const unsigned dg = ok ? (unsigned)((key >> shift) & 255ull) : 256u + lane;
        const unsigned peers = __match_any_sync(0xffffffffu, dg);
        const unsigned before = __popc(peers & ((1u << lane) - 1u));      // stable: earlier lanes with the same digit first
        unsigned dst = 0;
        if (ok) dst = base[warp][dg] + before;
        __syncwarp();
        if (ok && before == 0) base[warp][dg] += __popc(peers);
        __syncwarp();
        if (ok) {
            keys_out[dst] = key;
            vals_out[dst] = vals ? vals[i] : (int)i;
        }
    }
}

__global__ void inverse_permutation_kernel(const int* __restrict__ order, int64_t n, int* __restrict__ inv) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) inv[order[i]] = (int)i;
}

// out[0] = 0, out[i + 1] = sum_{j <= i} a[j] (int32 counts -> int64 offsets), one CTA, fixed order
__global__ void __launch_bounds__(1024)
scan_i32_to_i64_kernel(const int* __restrict__ a, int64_t n, long long* __restrict__ out) {
    cta_exclusive_scan<int, long long, true>(a, n, out);
}

__global__ void __launch_bounds__(1024)
scan_i32_to_i32_kernel(const int* __restrict__ a, int64_t n, int* __restrict__ out) {
    cta_exclusive_scan<int, int, true>(a, n, out);
}

// out[i][:] = in[map[i]][:] (rows of B doubles)
__global__ void gather_rows_kernel(const double* __restrict__ in, const int* __restrict__ map, int64_t n, int B, double* __restrict__ out) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * B) return;
    const int64_t i = t / B;
    const int c = (int)(t - i * B);
    out[t] = in[(int64_t)map[i] * B + c];
}

}  // namespace gp

using namespace gp;

extern "C" {

// out_host[0 .. d) = column minima, out_host[d .. 2 d) = column maxima of the device points (n x d); synchronises the stream
int gp_points_bbox(const double* points, int64_t n, int64_t d, double* out_host, void* ws, void* stream) {
    if (!points || !out_host || !ws || n <= 0 || d <= 0 || d > 8) return -1;
    cudaStream_t s = (cudaStream_t)stream;
    const int nparts = 148;
    double* part = (double*)ws;
    double* out = part + (int64_t)nparts * d * 2;
    minmax_partial_kernel<<<nparts, 256, 0, s>>>(points, n, (int)d, part);
    minmax_final_kernel<<<1, 32, 0, s>>>(part, nparts, (int)d, out);
    GP_COUNT(2);
    GP_LAUNCH_CHECK();
    GP_CUDA_CHECK(cudaMemcpyAsync(out_host, out, sizeof(double) * 2 * d, cudaMemcpyDeviceToHost, s));
    GP_CUDA_CHECK(cudaStreamSynchronize(s));
    return 0;
}

int64_t gp_sort_workspace_bytes(int64_t n) {
    const int64_t ntiles = (n + RTILE - 1) / RTILE;
    return n * 8 + n * 4 + 256 * ntiles * 4 + 4096;       // second key buffer, second value buffer, histograms
}

// order_out[i] = index of the i-th smallest key (stable). keys_dev is used as scratch (overwritten). ws: gp_sort_workspace_bytes(n).
int gp_sort_keys_u64(unsigned long long* keys_dev, int64_t n, int key_bits, int* order_out, void* ws, void* stream) {
    if (!keys_dev || !order_out || !ws || n <= 0 || n > INT32_MAX || key_bits <= 0 || key_bits > 64) return -1;
    cudaStream_t s = (cudaStream_t)stream;
    const int ntiles = (int)((n + RTILE - 1) / RTILE);
    unsigned long long* kb = (unsigned long long*)ws;
    int* vb = (int*)(kb + n);
    unsigned* hist = (unsigned*)(((uintptr_t)(vb + n) + 255) & ~(uintptr_t)255);
    const int passes = (key_bits + 7) / 8;
    // ping-pong between the pairs P0 = (keys_dev, order_out) and P1 = (kb, vb); pass p writes the pair it does not read. The
    // last pass has to land in P0: with an even pass count the first source is P0, with an odd one the keys are first
    // copied to P1.
    bool src_is_p0 = (passes & 1) == 0;
    if (!src_is_p0) GP_CUDA_CHECK(cudaMemcpyAsync(kb, keys_dev, n * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, s));
    const int blocks = (ntiles + RWARPS - 1) / RWARPS;
    for (int p = 0; p < passes; ++p) {
        const int shift = 8 * p;
        const unsigned long long* ksrc = src_is_p0 ? keys_dev : kb;
        const int* vsrc = (p == 0) ? nullptr : (src_is_p0 ? order_out : vb);      // first pass: the values are the identity
        unsigned long long* kdst = src_is_p0 ? kb : keys_dev;
        int* vdst = src_is_p0 ? vb : order_out;
        radix_hist_kernel<<<blocks, RWARPS * 32, 0, s>>>(ksrc, n, shift, ntiles, hist);
        scan_u32_kernel<<<1, 1024, 0, s>>>(hist, (int64_t)256 * ntiles);
        radix_scatter_kernel<<<blocks, RWARPS * 32, 0, s>>>(ksrc, vsrc, n, shift, ntiles, hist, kdst, vdst);
        src_is_p0 = !src_is_p0;
    }
    GP_COUNT(3 * passes);
    GP_LAUNCH_CHECK();
    return 0;
}

int gp_inverse_permutation(const int* order, int64_t n, int* inv, void* stream) {
    if (!order || !inv || n <= 0) return -1;
    inverse_permutation_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(order, n, inv);
    GP_COUNT(1);
    GP_LAUNCH_CHECK();
    return 0;
}

int gp_scan_counts(const int* counts, int64_t n, int64_t* offsets, void* stream) {
    if (!counts || !offsets || n <= 0) return -1;
    scan_i32_to_i64_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(counts, n, (long long*)offsets);
    GP_COUNT(1);
    GP_LAUNCH_CHECK();
    return 0;
}

// the same with 32-bit offsets (CSR row pointers); the caller guarantees that the total fits
int gp_scan_counts_i32(const int* counts, int64_t n, int* offsets, void* stream) {
    if (!counts || !offsets || n <= 0) return -1;
    scan_i32_to_i32_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(counts, n, offsets);
    GP_COUNT(1);
    GP_LAUNCH_CHECK();
    return 0;
}

int gp_gather_rows(const double* in, const int* map, int64_t n, int64_t B, double* out, void* stream) {
    if (!in || !map || !out || n <= 0 || B <= 0 || in == out) return -1;
    const int64_t total = n * B;
    gather_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(in, map, n, (int)B, out);
    GP_COUNT(1);
    GP_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
