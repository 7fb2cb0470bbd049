// common.cuh -- context, error plumbing, device buffers and small load/store helpers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <new>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/moka_b200.h"

// Cells per block (= threads per block) of the fused stage / step kernels and their adjoints: a block owns this many
// consecutive (space-filling-curve ordered) cells and the edges they own.  256 is what every measurement so far used; the
// resident-blocks hints of the kernels scale with it so that the register budget per thread stays the same.
#ifndef MOKAB_BLOCK_CELLS
#define MOKAB_BLOCK_CELLS 256
#endif
#define MOKAB_BLOCKS_SCALED(n) ((n) * 256 / MOKAB_BLOCK_CELLS)

namespace mokab {

extern thread_local std::string g_last_error;

struct Error : std::runtime_error {
    using std::runtime_error::runtime_error;
};

#define MOKAB_CUDA(expr)                                                                     \
    do {                                                                                     \
        cudaError_t e__ = (expr);                                                            \
        if (e__ != cudaSuccess)                                                              \
            throw ::mokab::Error(std::string(#expr) + " failed: " + cudaGetErrorString(e__) + \
                                 " (" __FILE__ ":" + std::to_string(__LINE__) + ")");        \
    } while (0)

#define MOKAB_REQUIRE(cond, msg)                       \
    do {                                               \
        if (!(cond)) throw ::mokab::Error(std::string(msg)); \
    } while (0)

// Run `body`, translate exceptions to the C convention (non-zero + thread-local message).
template <class F>
static inline int guarded(F &&body) noexcept
{
    try {
        body();
        return 0;
    } catch (const std::exception &ex) {
        g_last_error = ex.what();
        return 1;
    } catch (...) {
        g_last_error = "unknown error";
        return 2;
    }
}

template <class T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    bool owned = true;   // false: a window into somebody else's allocation (view)
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }
    void release()
    {
        if (p && owned) cudaFree(p);
        p = nullptr;
        n = 0;
        owned = true;
    }
    void view(T *ptr, size_t count)
    {
        release();
        p = ptr;
        n = count;
        owned = false;
    }
    void alloc(size_t count)
    {
        release();
        n = count;
        if (count) MOKAB_CUDA(cudaMalloc(&p, count * sizeof(T)));
    }
    void upload(const std::vector<T> &h, cudaStream_t s)
    {
        alloc(h.size());
        if (n) MOKAB_CUDA(cudaMemcpyAsync(p, h.data(), n * sizeof(T), cudaMemcpyHostToDevice, s));
    }
    void zero(cudaStream_t s)
    {
        if (n) MOKAB_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s));
    }
    size_t bytes() const { return n * sizeof(T); }
};

}  // namespace mokab

struct mokab_ctx {
    int device = 0;
    int num_sms = 148;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;  // the stream in use (own or caller's)
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int64_t launches = 0;
    void bind() const { MOKAB_CUDA(cudaSetDevice(device)); }
};

namespace mokab {

// ---- stage timeline (TRACE builds only: libmoka_b200_trace.so, -DMOKAB_TRACE) ---------------------------------------------
// Where does the stage time of a decomposed run go?  nsys is not in the image and ncu serialises kernels, so the trace build
// lets every block of the stage / halo kernels append one record -- which kernel, which block, %globaltimer at entry, after
// the gate (kernels that wait for peers) and at exit -- to a ring in device memory; tools/trace_stages.py turns the records
// of one step into a timeline per launch.  Nothing of this is compiled into libmoka_b200.so.
#ifdef MOKAB_TRACE
struct TraceRec { unsigned int kind, block, grid, pad; unsigned long long t0, t1, t2; };
__device__ TraceRec *g_trace_buf;
__device__ unsigned long long g_trace_count;
__device__ unsigned int g_trace_cap;
__device__ __forceinline__ unsigned long long trace_now()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void trace_emit(unsigned int kind, unsigned long long t0, unsigned long long t2)
{
    if (!g_trace_buf) return;
    const unsigned long long i = atomicAdd(&g_trace_count, 1ull) % g_trace_cap;
    TraceRec r;
    r.kind = kind; r.block = blockIdx.x; r.grid = gridDim.x; r.pad = 0; r.t0 = t0; r.t1 = trace_now(); r.t2 = t2;
    g_trace_buf[i] = r;
}
#define MOKAB_TRACE_BEGIN() unsigned long long tr0_ = 0, tr2_ = 0; if (threadIdx.x == 0) tr0_ = trace_now()
#define MOKAB_TRACE_MARK() do { if (threadIdx.x == 0) tr2_ = trace_now(); } while (0)
#define MOKAB_TRACE_END(kind) do { __syncthreads(); if (threadIdx.x == 0) trace_emit((kind), tr0_, tr2_); } while (0)
#else
#define MOKAB_TRACE_BEGIN() do { } while (0)
#define MOKAB_TRACE_MARK() do { } while (0)
#define MOKAB_TRACE_END(kind) do { } while (0)
#endif

// ---- device helpers -----------------------------------------------------------------------------
// Streaming (read-once) loads: keep them out of L1 so the gathered state stays resident there.
#ifdef MOKAB_SIM   // host build of the simulation tests (tests/sim): a plain load
template <class T>
__device__ __forceinline__ T ld_stream(const T *p) { return *p; }
__device__ __forceinline__ void prefetch_l2(const void *) {}
#else
// pull the line holding *p into L2 (no register, no fault on a bad address)
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
template <class T>
__device__ __forceinline__ T ld_stream(const T *p);
template <>
__device__ __forceinline__ int ld_stream<int>(const int *p)
{
    int v;
    asm("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
template <>
__device__ __forceinline__ unsigned short ld_stream<unsigned short>(const unsigned short *p)
{
    unsigned short v;
    asm("ld.global.nc.L1::no_allocate.u16 %0, [%1];" : "=h"(v) : "l"(p));
    return v;
}
template <>
__device__ __forceinline__ unsigned char ld_stream<unsigned char>(const unsigned char *p)
{
    unsigned int v;
    asm("ld.global.nc.L1::no_allocate.u8 %0, [%1];" : "=r"(v) : "l"(p));
    return (unsigned char)v;
}
template <>
__device__ __forceinline__ int2 ld_stream<int2>(const int2 *p)
{
    int2 v;
    asm("ld.global.nc.L1::no_allocate.v2.s32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}
template <>
__device__ __forceinline__ double ld_stream<double>(const double *p)
{
    double v;
    asm("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
template <>
__device__ __forceinline__ float ld_stream<float>(const float *p)
{
    float v;
    asm("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
#endif

}  // namespace mokab
