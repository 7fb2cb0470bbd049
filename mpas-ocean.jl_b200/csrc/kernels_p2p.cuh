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

// ---- flag-in-data exchange (MOKAB_HALO_P2P_LL) --------------------------------------------------------------------------------
// r02m's timeline of a 131 k-cell part: wait 1 -> boundary launch 7 -> k_halo_push 9 us -> ..., and what is left in the push kernel
// is ordering, not data: one MEMBAR.SYS round trip per block, the ticket, a second fence and the tick of the peers' counters.
// Here nothing is ordered at all.  A value crosses NVLink as 8-byte PACKETS {32 bits of the value, 32-bit exchange number}
// (one packet per Float32, two per Float64): an aligned 8-byte store is single-copy atomic, so whoever reads a packet whose
// upper half is the number it waits for has its lower half too -- NCCL's "LL" idea.  The sender's kernel just stores
// (st.relaxed.sys) into a receive area in the neighbour's memory, slot = the entity's position in the neighbour's own receive
// list; the neighbour's wait kernel polls one thread per slot and scatters the values into the halo slots of its state.
// Exchange numbers live in device memory (one counter per side, bumped by the last block of every launch), so captured graphs
// replay unchanged.  Slots are double-buffered by the parity of the exchange number: a rank starts exchange s + 2 only after
// its wait of s + 1 has returned, which the neighbour feeds after ITS wait of s has consumed the packets of that parity.  That
// needs every pair that exchanges anything to wait for each other, so besides the data every rank sends each peer one CREDIT
// packet (number only) per exchange: a neighbour that receives without sending still holds its sender back.
__device__ __forceinline__ void st_relaxed_sys(unsigned long long *p, unsigned long long v)
{
#ifdef MOKAB_SIM
    *p = v;
#else
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
#endif
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long *p)
{
#ifdef MOKAB_SIM
    return *p;
#else
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
#endif
}
template <class R>
__host__ __device__ __forceinline__ unsigned long long ll_bits(R v)
{
    unsigned long long b = 0ull;
    memcpy(&b, &v, sizeof(R));           // (little-endian: a Float32 fills the lower half)
    return b;
}
template <class R>
__host__ __device__ __forceinline__ R ll_value(unsigned long long b)
{
    R v;
    memcpy(&v, &b, sizeof(R));
    return v;
}
// slot `pos` of a receive area holds, for either parity of the exchange number, two packets (the second unused in Float32)
__device__ __forceinline__ size_t ll_slot(int pos, unsigned int seq) { return ((size_t)pos * 2 + (seq & 1u)) * 2; }

template <class R>
struct LLPushArgs {
    int n, nReal, nC;                    // items = the send list + one credit per peer; send list length; local cell count
    const int32_t *src;                  // send list: local entity in the combined index space [cells | edges]
    const int32_t *llDst;                // per item: its slot in the receiver's area
    const uint8_t *slot;                 // per item: which receiver
    const R *h, *u;                      // local stage output
    unsigned long long *const *peerLL;   // per receiver: its receive area
    unsigned int *seq, *done;            // exchanges issued so far; ticket counter (returns to 0 after every launch)
};

template <class R>
__global__ void __launch_bounds__(256) k_halo_push_ll(const LLPushArgs<R> A)
{
    MOKAB_TRACE_BEGIN();
    const unsigned int s = *(volatile unsigned int *)A.seq + 1u;
    const int k = blockIdx.x * 256 + threadIdx.x;
    if (k < A.n) {
        unsigned long long bits = 0ull;
        if (k < A.nReal) {
            const int i = A.src[k];
            bits = ll_bits<R>(i < A.nC ? A.h[i] : A.u[i - A.nC]);
        }
        unsigned long long *q = A.peerLL[A.slot[k]] + ll_slot(A.llDst[k], s);
        const unsigned long long tag = (unsigned long long)s << 32;
        st_relaxed_sys(q, (bits & 0xffffffffull) | tag);
        if (sizeof(R) == 8) st_relaxed_sys(q + 1, (bits >> 32) | tag);
    }
    __syncthreads();                                       // every thread of the block has read the exchange number
    if (threadIdx.x == 0) {
        const unsigned int ticket = add_device(A.done, 1u);
        if (ticket == gridDim.x - 1) { *A.done = 0u; *A.seq = s; }
    }
    MOKAB_TRACE_END(102u);
}

template <class R>
struct LLWaitArgs {
    int n, nReal, nC;                    // slots = the receive list + one credit per peer; receive list length; local cell count
    const int32_t *idx;                  // receive list: where slot k goes in the combined index space [cells | edges]
    const unsigned long long *ll;        // this rank's receive area
    R *h, *u;                            // the stage output whose halo slots are filled
    unsigned int *seq, *done;
    int *error;
    long long timeout_cycles;
};

// `ready` (simulated runtime: a host-side predicate retried by the stream scheduler; hardware: the spin below)
template <class R>
__host__ __device__ __forceinline__ bool ll_arrived(const unsigned long long *q, unsigned int s, unsigned long long *bits)
{
#if defined(__CUDA_ARCH__) && !defined(MOKAB_SIM)
    const unsigned long long lo = ld_relaxed_sys(q), hi = sizeof(R) == 8 ? ld_relaxed_sys(q + 1) : ((unsigned long long)s << 32);
#else
    const unsigned long long lo = q[0], hi = sizeof(R) == 8 ? q[1] : ((unsigned long long)s << 32);
#endif
    if ((unsigned int)(lo >> 32) != s || (unsigned int)(hi >> 32) != s) return false;
    *bits = (lo & 0xffffffffull) | (hi << 32);
    return true;
}

template <class R>
__global__ void __launch_bounds__(256) k_halo_wait_ll(const LLWaitArgs<R> A)
{
    MOKAB_TRACE_BEGIN();
    const unsigned int s = *(volatile unsigned int *)A.seq + 1u;
    const int k = blockIdx.x * 256 + threadIdx.x;
    if (k < A.n) {
        const unsigned long long *q = A.ll + ll_slot(k, s);
        unsigned long long bits = 0ull;
        bool ok = ll_arrived<R>(q, s, &bits);
#ifndef MOKAB_SIM
        const long long t0 = clock64();
        while (!ok) {
            if (*(volatile int *)A.error || clock64() - t0 > A.timeout_cycles) {   // a peer died or the schedules diverged: never hang the GPU
                atomicExch(A.error, 1);
                break;
            }
            __nanosleep(32);
            ok = ll_arrived<R>(q, s, &bits);
        }
#endif
        if (ok && k < A.nReal) {
            const int i = A.idx[k];
            if (i < A.nC) A.h[i] = ll_value<R>(bits);
            else A.u[i - A.nC] = ll_value<R>(bits);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int ticket = add_device(A.done, 1u);
        if (ticket == gridDim.x - 1) { *A.done = 0u; *A.seq = s; }
    }
    MOKAB_TRACE_END(103u);
}

}  // namespace p2p
}  // namespace mokab
