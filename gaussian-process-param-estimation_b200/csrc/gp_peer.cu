// Peer-memory context of the row-slab sparse operator: one process per GPU, each rank allocates an ARENA (mailboxes of the
// in-kernel all-reduce + the exchange vectors the SpMM of the other ranks gathers from), exports it as a CUDA IPC handle
// and maps the arenas of its peers; kernels then load / store peer memory directly over NVLink (csrc/gp_peer.cuh,
// gp_sparse_la.cu). The handles travel through the caller's process group (gaussian_proc/_slab.py: torch.distributed
// all_gather) - plumbing only. Layout of an arena (identical on every rank, so an offset names the same object everywhere):
//   [ payload[SLOTS][MAX][PAYLOAD] doubles | flags[SLOTS][MAX] u64 | pad to 64 KB | vec 0 | vec 1 | vec 2 ]
// with vec k = nloc_max x 32 doubles (B <= 32 columns of the widest slab).
#include "../../include/gpgp.h"
#include "gp_common.cuh"
#include "gp_peer.cuh"
#include <string.h>

namespace gp {

struct PeerCtx {
    int rank, world;
    int64_t nloc_max;
    size_t mail_bytes, vec_bytes, arena_bytes;
    char* arena[PEER_MAX];
    bool opened[PEER_MAX];
    unsigned long long seq;
    int* err;
};

static size_t peer_mail_bytes() {
    size_t b = sizeof(double) * PEER_SLOTS * PEER_MAX * PEER_PAYLOAD + sizeof(unsigned long long) * PEER_SLOTS * PEER_MAX;
    return (b + 65535) & ~(size_t)65535;
}

PeerComm peer_next(PeerCtx* ctx) {
    PeerComm pc;
    memset(&pc, 0, sizeof(pc));
    if (!ctx) { pc.world = 1; return pc; }
    pc.rank = ctx->rank;
    pc.world = ctx->world;
    pc.seq = ++ctx->seq;
    for (int p = 0; p < ctx->world; ++p) pc.mail[p] = (double*)ctx->arena[p];
    pc.err = ctx->err;
    return pc;
}

PeerVec peer_vec(PeerCtx* ctx, int k) {
    PeerVec pv;
    memset(&pv, 0, sizeof(pv));
    for (int p = 0; p < ctx->world; ++p) pv.base[p] = (const double*)(ctx->arena[p] + ctx->mail_bytes + (size_t)k * ctx->vec_bytes);
    pv.rank = ctx->rank;
    return pv;
}

double* peer_local_vec(PeerCtx* ctx, int k) {
    return (double*)(ctx->arena[ctx->rank] + ctx->mail_bytes + (size_t)k * ctx->vec_bytes);
}

int peer_world(const PeerCtx* ctx) { return ctx ? ctx->world : 1; }

__global__ void __launch_bounds__(32) peer_barrier_kernel(PeerComm pc) { (void)peer_block_sum(pc, 0.0, 0); }

// values[0 .. count) <- sum over the ranks, in place (count <= PEER_PAYLOAD)
__global__ void __launch_bounds__(PEER_PAYLOAD) peer_allreduce_kernel(double* values, int count, PeerComm pc) {
    const double v = (threadIdx.x < count) ? values[threadIdx.x] : 0.0;
    const double s = peer_block_sum(pc, v, count);
    if (threadIdx.x < count) values[threadIdx.x] = s;
}

int peer_barrier_launch(PeerCtx* ctx, cudaStream_t s) {
    if (!ctx || ctx->world <= 1) return 0;
    peer_barrier_kernel<<<1, 32, 0, s>>>(peer_next(ctx));
    GP_COUNT(1);
    GP_LAUNCH_CHECK();
    return 0;
}

}  // namespace gp

using namespace gp;

extern "C" {

int64_t gp_peer_handle_bytes(void) { return (int64_t)sizeof(cudaIpcMemHandle_t); }

// Allocates this rank's arena (device memory of the current device, zeroed) for slabs of at most nloc_max rows.
void* gp_peer_create(int64_t rank, int64_t world, int64_t nloc_max) {
    if (rank < 0 || world < 1 || world > PEER_MAX || rank >= world || nloc_max <= 0) return nullptr;
    PeerCtx* ctx = new PeerCtx();
    memset(ctx, 0, sizeof(*ctx));
    ctx->rank = (int)rank;
    ctx->world = (int)world;
    ctx->nloc_max = nloc_max;
    ctx->mail_bytes = peer_mail_bytes();
    ctx->vec_bytes = ((size_t)nloc_max * 32 * sizeof(double) + 65535) & ~(size_t)65535;
    ctx->arena_bytes = ctx->mail_bytes + PEER_VECS * ctx->vec_bytes;
    void* a = nullptr;
    if (cudaMalloc(&a, ctx->arena_bytes) != cudaSuccess) { delete ctx; return nullptr; }
    if (cudaMemset(a, 0, ctx->mail_bytes) != cudaSuccess || cudaMalloc((void**)&ctx->err, sizeof(int)) != cudaSuccess) {
        cudaFree(a);
        delete ctx;
        return nullptr;
    }
    cudaMemset(ctx->err, 0, sizeof(int));
    cudaDeviceSynchronize();
    ctx->arena[rank] = (char*)a;
    return ctx;
}

int gp_peer_handle(void* ctx_, unsigned char* handle_out) {
    PeerCtx* ctx = (PeerCtx*)ctx_;
    if (!ctx || !handle_out) return -1;
    cudaIpcMemHandle_t h;
    GP_CUDA_CHECK(cudaIpcGetMemHandle(&h, ctx->arena[ctx->rank]));
    memcpy(handle_out, &h, sizeof(h));
    return 0;
}

// handles: world x gp_peer_handle_bytes(), rank-major (entry `rank` is ignored). Maps every peer's arena (peer access over
// NVLink is enabled by the IPC open). The caller places a process-group barrier after it (every mailbox is zeroed by then).
int gp_peer_connect(void* ctx_, const unsigned char* handles) {
    PeerCtx* ctx = (PeerCtx*)ctx_;
    if (!ctx || (ctx->world > 1 && !handles)) return -1;
    for (int p = 0; p < ctx->world; ++p) {
        if (p == ctx->rank || ctx->arena[p]) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + (size_t)p * sizeof(h), sizeof(h));
        void* ptr = nullptr;
        GP_CUDA_CHECK(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        ctx->arena[p] = (char*)ptr;
        ctx->opened[p] = true;
    }
    return 0;
}

int gp_peer_destroy(void* ctx_) {
    PeerCtx* ctx = (PeerCtx*)ctx_;
    if (!ctx) return -1;
    cudaDeviceSynchronize();
    for (int p = 0; p < ctx->world; ++p)
        if (ctx->opened[p]) cudaIpcCloseMemHandle(ctx->arena[p]);
    cudaFree(ctx->arena[ctx->rank]);
    cudaFree(ctx->err);
    delete ctx;
    return 0;
}

// this rank's copy of exchange vector k (nloc_max x 32 doubles)
double* gp_peer_vec(void* ctx_, int64_t k) {
    PeerCtx* ctx = (PeerCtx*)ctx_;
    if (!ctx || k < 0 || k >= PEER_VECS) return nullptr;
    return peer_local_vec(ctx, (int)k);
}

// cross-GPU barrier in stream order: everything the ranks enqueued before it is complete (and visible to peers) after it
int gp_peer_barrier(void* ctx_, void* stream) { return peer_barrier_launch((PeerCtx*)ctx_, (cudaStream_t)stream); }

// values_dev[0 .. count) <- sum over the ranks (rank order: bit-identical everywhere); count <= 256
int gp_peer_allreduce(void* ctx_, double* values_dev, int64_t count, void* stream) {
    PeerCtx* ctx = (PeerCtx*)ctx_;
    if (!ctx || !values_dev || count <= 0 || count > PEER_PAYLOAD) return -1;
    if (ctx->world <= 1) return 0;
    peer_allreduce_kernel<<<1, PEER_PAYLOAD, 0, (cudaStream_t)stream>>>(values_dev, (int)count, peer_next(ctx));
    GP_COUNT(1);
    GP_LAUNCH_CHECK();
    return 0;
}

// 1 if a wait of this rank timed out (a peer never arrived: the ranks diverged); synchronises the stream
int gp_peer_error(void* ctx_, void* stream) {
    PeerCtx* ctx = (PeerCtx*)ctx_;
    if (!ctx) return -1;
    int e = 0;
    GP_CUDA_CHECK(cudaMemcpyAsync(&e, ctx->err, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    GP_CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
    return e;
}

}  // extern "C"
