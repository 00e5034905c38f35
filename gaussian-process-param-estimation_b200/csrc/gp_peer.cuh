// Device side of the peer-memory exchanges of the row-slab sparse operator (sm_100a, NVLink / NVSwitch P2P).
#pragma once
#include "gp_internal.h"

namespace gp {

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;\n" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];\n" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_sys(double* p, double v) {
    asm volatile("st.relaxed.sys.global.f64 [%0], %1;\n" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ double ld_relaxed_sys(const double* p) {
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];\n" : "=d"(v) : "l"(p) : "memory");
    return v;
}

constexpr long long PEER_TIMEOUT_CYCLES = 40000000000ll;     // ~ 20 s: a rank that never arrives sets err instead of hanging

// Sum over the ranks of `count` (<= PEER_PAYLOAD) values, value i held by thread i of the CTA; EVERY thread of the CTA calls.
// Push model: each rank stores its values into slot (seq mod SLOTS) of EVERY rank's mailbox (its own included) over NVLink,
// fences, then raises its flag there to seq; each rank polls only its OWN mailbox and adds the world contributions in rank
// order - every rank obtains bit-identical sums, so the scalar recurrences of the Krylov iterations stay in lock step.
// A rank can be at most one exchange ahead of another (it needs that rank's contribution to finish the current one), so the
// ring never wraps onto a slot that is still being read. The flag a peer observes also publishes everything the sending
// GPU wrote BEFORE this kernel (stream order + fence.sys): the SpMM that follows may gather the sender's freshly written rows.
__device__ __forceinline__ double peer_block_sum(const PeerComm& pc, double v, int count) {
    if (pc.world <= 1) return v;
    const int tid = threadIdx.x;
    const int slot = (int)(pc.seq % PEER_SLOTS);
    const int64_t flags_off = (int64_t)PEER_SLOTS * PEER_MAX * PEER_PAYLOAD;
    if (tid < count) {
        for (int p = 0; p < pc.world; ++p)
            st_relaxed_sys(pc.mail[p] + ((int64_t)slot * PEER_MAX + pc.rank) * PEER_PAYLOAD + tid, v);
        __threadfence_system();
    }
    __syncthreads();
    if (tid < pc.world) {
        __threadfence_system();
        st_release_sys(reinterpret_cast<unsigned long long*>(pc.mail[tid] + flags_off) + slot * PEER_MAX + pc.rank, pc.seq);
        const unsigned long long* mine =
            reinterpret_cast<const unsigned long long*>(pc.mail[pc.rank] + flags_off) + slot * PEER_MAX + tid;
        const long long t0 = clock64();
        while (ld_acquire_sys(mine) < pc.seq) {
            if (clock64() - t0 > PEER_TIMEOUT_CYCLES) {
                atomicExch(pc.err, 1);
                break;
            }
        }
    }
    __syncthreads();
    double s = 0.0;
    if (tid < count)
        for (int p = 0; p < pc.world; ++p)
            s += ld_relaxed_sys(pc.mail[pc.rank] + ((int64_t)slot * PEER_MAX + p) * PEER_PAYLOAD + tid);
    return s;
}

}  // namespace gp
