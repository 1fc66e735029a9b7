// kernels_peer.cu -- epoch handshake of the fused peer-memory halo exchange.
//
// With the fused path the sweep kernel itself stores the slab's first/last valid plane into the
// neighbour GPU's ghost plane (NVLink peer stores, kernels_sweep_tma.cu).  Two tiny stream-ordered
// kernels keep neighbouring ranks within one sweep of each other:
//   signal: after sweep e, write e+1 into the neighbours' flag words (st.release.sys, after a
//           system fence; the preceding sweep kernel's peer stores are ordered before it by the
//           kernel boundary);
//   wait:   before sweep e, spin until both neighbours have signalled >= e: their sweep e-1 has
//           (a) delivered my ghost planes of the buffer I am about to read and (b) finished
//           reading the buffer whose ghost planes I am about to overwrite.
// The waiter only depends on progress of ANOTHER GPU, never on work queued behind it on its own
// device, so there is no co-residency requirement.  The spin is bounded: on timeout it records a
// status word instead of hanging the device.
#include "common.cuh"

namespace {

__global__ void peer_signal_kernel(unsigned long long* a, unsigned long long* b, unsigned long long value)
{
    __threadfence_system();
    if (a) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(a), "l"(value) : "memory");
    if (b) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(b), "l"(value) : "memory");
}

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__global__ void peer_wait_kernel(const unsigned long long* a, const unsigned long long* b, unsigned long long target,
                                 unsigned long long* status)
{
    const long long t0 = clock64();
    const long long limit = 20LL * 1000 * 1000 * 1000;      // ~10 s at 2 GHz: a dead neighbour, not a slow one
    for (;;) {
        const bool oa = !a || ld_acquire_sys(a) >= target;
        const bool ob = !b || ld_acquire_sys(b) >= target;
        if (oa && ob) break;
        if (clock64() - t0 > limit) {
            *status = target;                               // reported by the host at the next sync point
            break;
        }
        __nanosleep(200);
    }
    __threadfence_system();
}

}  // namespace

int launch_peer_signal(unsigned long long* peer_flag_a, unsigned long long* peer_flag_b, unsigned long long value, cudaStream_t s)
{
    if (!peer_flag_a && !peer_flag_b) return 0;
    peer_signal_kernel<<<1, 1, 0, s>>>(peer_flag_a, peer_flag_b, value);
    return 1;
}

int launch_peer_wait(const unsigned long long* flag_a, const unsigned long long* flag_b, unsigned long long target,
                     unsigned long long* status, cudaStream_t s)
{
    if (!flag_a && !flag_b) return 0;
    peer_wait_kernel<<<1, 1, 0, s>>>(flag_a, flag_b, target, status);
    return 1;
}
