// FP64 tensor-core (DMMA) GEMM for the Cholesky / inverse hot path.
//
// One 128x128 output tile per CTA, 8 warps (2x4), warp tile 64x32 built from m8n8k4 DMMA fragments,
// BK=32 k-slabs moved global->shared with cp.async through a 3-stage ring (measured: 2-6 % over BK=16 x 4 stages,
// the per-slab CTA barrier is the main loss against the raw DMMA issue rate). Shared tiles are XOR-swizzled
// so that both the 16-byte cp.async stores and the 8-byte fragment loads are bank-conflict free for
// K-major ([mn][k]) as well as MN-major ([k][mn]) operands; this lets one kernel serve
//   NT: trailing SYRK/GEMM update and the panel TRSM-by-inverse   (potrf)
//   NN: triangular products of the recursive inverse               (trtri)
//   TN: W^T W                                                      (lauum)
// Triangular operands are exploited by restricting each tile's k-range, never by element masks.
#include "gp_common.cuh"
#include "gp_internal.h"
#include <vector>

namespace gp {

unsigned long long g_launch_count = 0;

// optional per-launch timing of the DMMA GEMM (bench.py's roofline leg): events around every launch
struct GemmProfile {
    bool on = false;
    std::vector<cudaEvent_t> pool;
    size_t used = 0;
    double flops = 0.0;       // executed tile flops (2 * 128 * 128 * k per computed tile)
    long long launches = 0;
};
static GemmProfile g_prof;

static double tile_flops(int tiles_m, int tiles_n, int K, int krange, int tmask) {
    double kt = 0.0;  // sum over computed tiles of their k extent
    for (int tm = 0; tm < tiles_m; ++tm) {
        int ncols = (tmask == TM_LOWER) ? (tm + 1 < tiles_n ? tm + 1 : tiles_n) : tiles_n;
        if (krange == KR_FULL) kt += (double)ncols * K;
        else if (krange == KR_A_LOWER) kt += (double)ncols * ((tm + 1) * 128 < K ? (tm + 1) * 128 : K);
        else if (krange == KR_TN_LOWER) kt += (double)ncols * (K - (tm * 128 < K ? tm * 128 : K));
        else for (int tn = 0; tn < ncols; ++tn) kt += (double)(K - (tn * 128 < K ? tn * 128 : K));
    }
    return 2.0 * 128.0 * 128.0 * kt;
}

#ifndef GP_BK
#define GP_BK 32
#endif
#ifndef GP_STAGES
#define GP_STAGES 3
#endif
#ifndef GP_WM
#define GP_WM 64
#endif
#ifndef GP_WN
#define GP_WN 32
#endif
constexpr int BM = 128, BN = 128, BK = GP_BK, STAGES = GP_STAGES;
constexpr int WM = GP_WM, WN = GP_WN;
constexpr int WARPS_N = BN / WN, GEMM_THREADS = (BM / WM) * (BN / WN) * 32;
constexpr int MI = WM / 8, NI = WN / 8;
constexpr int TILE_ELEMS = 128 * BK;  // doubles per operand tile per stage
constexpr int CHUNKS_PER_THREAD = TILE_ELEMS / 2 / GEMM_THREADS;
constexpr int GEMM_SMEM = STAGES * 2 * TILE_ELEMS * (int)sizeof(double);
constexpr int RASTER_GROUP = 8;

// shared-memory offset (in doubles) of logical element (mn, k) of an operand tile
template <int T>
__device__ __forceinline__ int soff(int mn, int k) {
    if (T == 0) return mn * BK + (((k >> 2) ^ (mn & 3)) << 2) + (k & 3);  // [128][BK], 4-double chunks swizzled by row
    return k * 128 + (mn ^ ((k & 3) << 2));                                // [16][128], mn bits 2..3 swizzled by k
}

// copy one operand tile (128 mn x 16 k) into shared memory; g points at logical element (mn0, k0)
template <int T>
__device__ __forceinline__ void load_tile(double* tile, const double* g, int64_t ld, int tid) {
#pragma unroll
    for (int i = 0; i < CHUNKS_PER_THREAD; ++i) {
        int idx = tid + i * GEMM_THREADS;
        if (T == 0) {
            int r = idx / (BK / 2), c = idx % (BK / 2);  // row r, 16-byte chunk c (k = 2c, 2c+1)
            cp_async16(tile + r * BK + ((c >> 1) ^ (r & 3)) * 4 + ((c & 1) << 1), g + (int64_t)r * ld + 2 * c);
        } else {
            int kr = idx >> 6, c = idx & 63;  // k-row kr, chunk c (mn = 2c, 2c+1)
            cp_async16(tile + kr * 128 + ((2 * c) ^ ((kr & 3) << 2)), g + (int64_t)kr * ld + 2 * c);
        }
    }
}

template <int AT, int BT>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
dgemm_dmma_kernel(double* C, int64_t ldc, const double* A, int64_t lda, const double* B, int64_t ldb,
                  int tiles_m, int tiles_n, int K, double alpha, double beta, int krange, int tmask) {
    extern __shared__ __align__(16) double smem[];

    // grouped rasterisation: RASTER_GROUP tile-rows share each B tile while it is hot in L2
    int bid = blockIdx.x;
    int group_sz = RASTER_GROUP * tiles_n;
    int grp = bid / group_sz;
    int first_m = grp * RASTER_GROUP;
    int rows_in_grp = min(RASTER_GROUP, tiles_m - first_m);
    int rem = bid - grp * group_sz;
    int tm = first_m + rem % rows_in_grp;
    int tn = rem / rows_in_grp;
    if (tmask == TM_LOWER && tn > tm) return;

    const int m0 = tm * BM, n0 = tn * BN;
    int kbeg = 0, kend = K;
    if (krange == KR_A_LOWER) kend = min(K, m0 + BM);
    else if (krange == KR_B_LOWER) kbeg = min(n0, K);
    else if (krange == KR_TN_LOWER) kbeg = min(max(m0, n0), K);
    const int KT = (kend - kbeg) / BK;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = (warp / WARPS_N) * WM, wn = (warp % WARPS_N) * WN;

    const double* Ag = (AT == 0) ? A + (int64_t)m0 * lda + kbeg : A + (int64_t)kbeg * lda + m0;
    const double* Bg = (BT == 0) ? B + (int64_t)n0 * ldb + kbeg : B + (int64_t)kbeg * ldb + n0;
    const int64_t a_step = (AT == 0) ? BK : (int64_t)BK * lda;
    const int64_t b_step = (BT == 0) ? BK : (int64_t)BK * ldb;

    double acc[MI][NI][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    if (beta != 0.0) {
        // the epilogue read-modify-writes a 128 KB tile of C that streams from HBM: pull it into L2 now so those
        // loads cost an L2 hit instead of a DRAM round trip each (1024 lines of 128 B, 4 per thread)
#pragma unroll
        for (int i = 0; i < 1024 / GEMM_THREADS; ++i) {
            int line = tid + i * GEMM_THREADS;  // row = line >> 3, 128-byte segment = line & 7
            const double* pc = C + (int64_t)(m0 + (line >> 3)) * ldc + n0 + ((line & 7) << 4);
            asm volatile("prefetch.global.L2 [%0];" ::"l"(pc));
        }
    }

    // prologue: fill STAGES-1 slabs
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < KT) {
            load_tile<AT>(smem + s * 2 * TILE_ELEMS, Ag + s * a_step, lda, tid);
            load_tile<BT>(smem + s * 2 * TILE_ELEMS + TILE_ELEMS, Bg + s * b_step, ldb, tid);
        }
        cp_async_commit();
    }

    for (int kt = 0; kt < KT; ++kt) {
#ifndef GP_EXP_NOSYNC
        cp_async_wait<STAGES - 2>();
        __syncthreads();
#endif
        {   // refill the slot consumed in the previous iteration
            int nk = kt + STAGES - 1;
            if (nk < KT) {
                int s = nk % STAGES;
                load_tile<AT>(smem + s * 2 * TILE_ELEMS, Ag + nk * a_step, lda, tid);
                load_tile<BT>(smem + s * 2 * TILE_ELEMS + TILE_ELEMS, Bg + nk * b_step, ldb, tid);
            }
            cp_async_commit();
        }
        const double* As = smem + (kt % STAGES) * 2 * TILE_ELEMS;
        const double* Bs = As + TILE_ELEMS;
#pragma unroll
        for (int kk = 0; kk < BK / 4; ++kk) {
            double a[MI], b[NI];
#pragma unroll
            for (int i = 0; i < MI; ++i) a[i] = As[soff<AT>(wm + i * 8 + g, kk * 4 + t)];
#pragma unroll
            for (int j = 0; j < NI; ++j) b[j] = Bs[soff<BT>(wn + j * 8 + g, kk * 4 + t)];
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < NI; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }
    cp_async_wait<0>();

    // epilogue: each thread owns (row, 2 consecutive cols) of every 8x8 fragment -> 16-byte accesses
    const bool diag = (tmask == TM_LOWER) && (tm == tn);
    if (beta != 0.0 && !diag) {
        // full tile with accumulate: issue all loads of a row group before using them (memory-level parallelism)
#pragma unroll
        for (int i = 0; i < MI; i += 2) {
            double2 o[2][NI];
#pragma unroll
            for (int ii = 0; ii < 2; ++ii)
#pragma unroll
                for (int j = 0; j < NI; ++j)
                    o[ii][j] = *reinterpret_cast<const double2*>(C + (int64_t)(m0 + wm + (i + ii) * 8 + g) * ldc + n0 + wn + j * 8 + 2 * t);
#pragma unroll
            for (int ii = 0; ii < 2; ++ii)
#pragma unroll
                for (int j = 0; j < NI; ++j) {
                    double2 v;
                    v.x = alpha * acc[i + ii][j][0] + beta * o[ii][j].x;
                    v.y = alpha * acc[i + ii][j][1] + beta * o[ii][j].y;
                    *reinterpret_cast<double2*>(C + (int64_t)(m0 + wm + (i + ii) * 8 + g) * ldc + n0 + wn + j * 8 + 2 * t) = v;
                }
        }
        return;
    }
#pragma unroll
    for (int i = 0; i < MI; ++i) {
        int row = wm + i * 8 + g;
        double* crow = C + (int64_t)(m0 + row) * ldc + n0;
#pragma unroll
        for (int j = 0; j < NI; ++j) {
            int col = wn + j * 8 + 2 * t;
            if (diag && col > row) continue;
            double2 v;
            v.x = alpha * acc[i][j][0];
            v.y = alpha * acc[i][j][1];
            double2* p = reinterpret_cast<double2*>(crow + col);
            if (beta != 0.0) {
                double2 o = *p;
                v.x += beta * o.x;
                v.y += beta * o.y;
            }
            if (diag && col + 1 > row) {
                crow[col] = v.x;  // keep the strictly-upper neighbour untouched
            } else {
                *p = v;
            }
        }
    }
}

template <int AT, int BT>
static int launch_inst(double* C, int64_t ldc, const double* A, int64_t lda, const double* B, int64_t ldb,
                       int M, int N, int K, double alpha, double beta, int krange, int tmask, cudaStream_t stream) {
    static bool configured = false;
    if (!configured) {
        GP_CUDA_CHECK(cudaFuncSetAttribute(dgemm_dmma_kernel<AT, BT>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM));
        configured = true;
    }
    int tiles_m = M / BM, tiles_n = N / BN;
    if (tiles_m == 0 || tiles_n == 0) return 0;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (g_prof.on) {
        while (g_prof.pool.size() < g_prof.used + 2) {
            cudaEvent_t e;
            GP_CUDA_CHECK(cudaEventCreate(&e));
            g_prof.pool.push_back(e);
        }
        e0 = g_prof.pool[g_prof.used++];
        e1 = g_prof.pool[g_prof.used++];
        g_prof.flops += tile_flops(tiles_m, tiles_n, K, krange, tmask);
        g_prof.launches++;
        cudaEventRecord(e0, stream);
    }
    dgemm_dmma_kernel<AT, BT><<<tiles_m * tiles_n, GEMM_THREADS, GEMM_SMEM, stream>>>(
        C, ldc, A, lda, B, ldb, tiles_m, tiles_n, K, alpha, beta, krange, tmask);
    if (e1) cudaEventRecord(e1, stream);
    GP_COUNT(1);
    GP_LAUNCH_CHECK();
    return 0;
}

int launch_dgemm(int at, int bt, double* C, int64_t ldc, const double* A, int64_t lda, const double* B, int64_t ldb,
                 int M, int N, int K, double alpha, double beta, int krange, int tmask, cudaStream_t stream) {
    if (M < 0 || N < 0 || K < 0 || (M % BM) || (N % BN) || (K % BK)) return -1;
    if ((lda & 1) || (ldb & 1) || (ldc & 1)) return -2;
    if (((uintptr_t)A | (uintptr_t)B | (uintptr_t)C) & 15) return -3;
    if (at == 0 && bt == 0) return launch_inst<0, 0>(C, ldc, A, lda, B, ldb, M, N, K, alpha, beta, krange, tmask, stream);
    if (at == 0 && bt == 1) return launch_inst<0, 1>(C, ldc, A, lda, B, ldb, M, N, K, alpha, beta, krange, tmask, stream);
    if (at == 1 && bt == 1) return launch_inst<1, 1>(C, ldc, A, lda, B, ldb, M, N, K, alpha, beta, krange, tmask, stream);
    if (at == 1 && bt == 0) return launch_inst<1, 0>(C, ldc, A, lda, B, ldb, M, N, K, alpha, beta, krange, tmask, stream);
    return -4;
}

int profile_enable(int on) {
    g_prof.on = on != 0;
    g_prof.used = 0;
    g_prof.flops = 0.0;
    g_prof.launches = 0;
    return 0;
}

// synchronises the device; returns summed GEMM kernel milliseconds, executed flops and launch count since enable
int profile_read(double* ms, double* flops, long long* launches) {
    GP_CUDA_CHECK(cudaDeviceSynchronize());
    double total = 0.0;
    for (size_t i = 0; i + 1 < g_prof.used; i += 2) {
        float t = 0.f;
        GP_CUDA_CHECK(cudaEventElapsedTime(&t, g_prof.pool[i], g_prof.pool[i + 1]));
        total += t;
    }
    *ms = total;
    *flops = g_prof.flops;
    *launches = g_prof.launches;
    return 0;
}

}  // namespace gp
