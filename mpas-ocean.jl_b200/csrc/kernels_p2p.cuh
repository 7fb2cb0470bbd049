// kernels_p2p.cuh -- halo exchange by direct stores into the peers' memory over NVLink (no NCCL on the data path).
//
// The packed all-to-all path costs four launches per RK stage on the halo stream (boundary blocks, pack, the NCCL
// collective, unpack); on the latency-bound decompositions (Kelvin 1024x1024 over 8 GPUs: ~13 us of stage kernel per
// GPU against ~90 us per stage measured) that chain IS the step time.  Here the sender writes every halo value
// straight into the receiver's state arrays (peer pointers from CUDA IPC) at the receiver's own index for it, and
// raises a per-sender arrival counter in the receiver's memory; the receiver's next boundary launch is preceded by a
// one-warp kernel that waits for the counters of the peers it receives from.  Two short kernels per stage, nothing
// else: no message buffers, no unpack, no host involvement, and -- all protocol state living in device memory as
// monotonically increasing counters -- the whole schedule is capture/replay invariant.
//
// Ordering.  Sender: data stores -> block barrier -> one thread per block fences at system scope (cumulative over the
// stores the barrier ordered before it) and bumps a local done-counter; the block that sees the last ticket fences again
// and adds 1 (system scope) to its slot in each receiver's arrival array.  Receiver: thread q spins with acquire loads until arrival[q] has
// reached the count it expects (its own counter of completed waits + 1), then the kernel ends and the stream order
// makes the data visible to the stage kernel that follows.  Hazards: a rank can run at most one
// RK stage ahead of a peer (its stage t needs the peer's stage t-1 counter), consecutive stage outputs alternate
// between two provisional buffers and the time levels alternate per step, so a pushed value never lands in a slot
// whose previous content the receiver still has to read.
#pragma once
#include "common.cuh"

namespace mokab {
namespace p2p {

constexpr int kMaxPeers = 64;

#ifdef MOKAB_SIM   // the host build of the simulation tests: one address space, the global lock serialises everything
__device__ __forceinline__ void fence_system() {}
__device__ __forceinline__ unsigned long long add_system(unsigned long long *p, unsigned long long v) { unsigned long long o = *p; *p = o + v; return o; }
__device__ __forceinline__ unsigned int add_device(unsigned int *p, unsigned int v) { unsigned int o = *p; *p = o + v; return o; }
__device__ __forceinline__ unsigned long long load_acquire_system(const unsigned long long *p) { return *p; }
#else
__device__ __forceinline__ void fence_system() { __threadfence_system(); }
__device__ __forceinline__ unsigned long long add_system(unsigned long long *p, unsigned long long v) { return atomicAdd_system(p, v); }
__device__ __forceinline__ unsigned int add_device(unsigned int *p, unsigned int v) { return atomicAdd(p, v); }
__device__ __forceinline__ unsigned long long load_acquire_system(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
#endif

template <class R>
struct PushArgs {
    int n;                         // entries (all receivers concatenated, message order)
    int nC;                        // local cell count (entries < nC are cells, the rest edges shifted by nC)
    const int32_t *src;            // local entity of entry k in the combined index space [cells | edges]
    const int32_t *dst;            // where it lives on the receiver: c >= 0 -> cell c, d < 0 -> edge -d - 1
    const uint8_t *slot;           // which receiver (index into the tables below)
    const R *h, *u;                // local stage output
    R *const *peerH, *const *peerU;  // per receiver: the same stage output arrays in ITS memory
    unsigned int *done;            // local ticket counter (returns to 0 after every launch)
    unsigned long long *const *arrival;  // per receiver: this rank's slot in its arrival array
    int nrecv;                     // number of receivers
};

template <class R>
__global__ void __launch_bounds__(256) k_halo_push(const PushArgs<R> A)
{
    MOKAB_TRACE_BEGIN();
    const int k = blockIdx.x * 256 + threadIdx.x;
    if (k < A.n) {
        const int s = A.src[k];
        const R v = s < A.nC ? A.h[s] : A.u[s - A.nC];
        const int d = A.dst[k];
        const int p = A.slot[k];
        if (d >= 0) A.peerH[p][d] = v;
        else A.peerU[p][-d - 1] = v;
    }
    // ONE system-scope fence per block, by the thread that takes the ticket after the block barrier (the barrier orders the
    // other threads' stores before it, the fence is cumulative) -- not one per storing thread: r02l's timeline showed this
    // kernel lasting 16 us whatever its size, i.e. the time of its ~160 per-warp MEMBAR.SYS.
    __syncthreads();
    if (threadIdx.x == 0) {
        fence_system();
        const unsigned int ticket = add_device(A.done, 1u);
        if (ticket == gridDim.x - 1) {              // every block's stores are ordered before its ticket
            *A.done = 0u;
            fence_system();
            for (int p = 0; p < A.nrecv; ++p) add_system(A.arrival[p], 1ull);
        }
    }
    MOKAB_TRACE_END(100u);
}

// a rank with receivers but nothing to send them this stage still has to tick their counters
__global__ void __launch_bounds__(32) k_halo_signal(int nrecv, unsigned long long *const *arrival)
{
    if ((int)threadIdx.x < nrecv) add_system(arrival[threadIdx.x], 1ull);
}

// true once every sender's arrival counter has reached the next expected value; then bumps the expectations
__device__ __forceinline__ bool sender_ready(const unsigned long long *arrival, const unsigned long long *expect, int sender)
{
    return load_acquire_system(arrival + sender) >= expect[sender] + 1ull;
}

__global__ void __launch_bounds__(kMaxPeers)
k_halo_wait(int nsend, const int32_t *senders, const unsigned long long *arrival, unsigned long long *expect, int *error,
            long long timeout_cycles)
{
    MOKAB_TRACE_BEGIN();
    const int i = threadIdx.x;
    if (i < nsend) {
        const int q = senders[i];
#ifndef MOKAB_SIM
        const long long t0 = clock64();
        while (!sender_ready(arrival, expect, q)) {
            if (*(volatile int *)error || clock64() - t0 > timeout_cycles) {       // a peer died or the schedules diverged: never hang the GPU
                atomicExch(error, 1);
                break;
            }
            __nanosleep(64);
        }
#endif
        expect[q] += 1ull;
    }
    MOKAB_TRACE_END(101u);
}

// The variant for launches that carry the exchange themselves (fused::k_rk_stage<..., PUSH>): those count their own
// completions in `expect`, so "everything sent to me so far has arrived" is arrival >= expect, and nothing is bumped here.
__global__ void __launch_bounds__(kMaxPeers)
k_halo_wait_arrivals(int nsend, const int32_t *senders, const unsigned long long *arrival, const unsigned long long *expect, int *error,
                     long long timeout_cycles)
{
#ifndef MOKAB_SIM
    const int i = threadIdx.x;
    if (i >= nsend) return;
    const int q = senders[i];
    const long long t0 = clock64();
    while (load_acquire_system(arrival + q) < expect[q]) {
        if (*(volatile int *)error || clock64() - t0 > timeout_cycles) {
            atomicExch(error, 1);
            break;
        }
        __nanosleep(64);
    }
#endif
}

}  // namespace p2p
}  // namespace mokab
