// cuda_runtime.h (SIMULATION SHIM) -- test infrastructure, never part of the product.
//
// tests/sim builds the library's own sources (mpas-ocean.jl_b200/csrc/*.cu*) for the HOST with g++ against this
// header instead of the CUDA toolkit's: kernels run on the CPU, one simulated thread at a time, and the runtime API
// below is backed by tests/sim/sim_runtime.cpp -- asynchronous streams with deferred, adversarially ordered
// execution, events, stream capture / graph replay with the captured dependency DAG only, poisoned allocations and
// in-stream collectives between emulated ranks.  The point is to check, in a container without a GPU, (1) the
// library's host logic and index arithmetic bit for bit against the oracle and (2) that every cross-stream
// dependency the code relies on is actually expressed (a missing one changes the result here deterministically,
// where on hardware it is a race that mostly wins).  It proves nothing about performance.
#pragma once
#include <stddef.h>
#include <stdint.h>

#ifndef MOKAB_SIM
#error "this header is the CPU simulation shim of tests/sim; the product builds with nvcc against the CUDA toolkit"
#endif

// ---- language shims -----------------------------------------------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __shared__ static   // blocks execute one after another on one host thread, so one instance per kernel is a block's

struct int2 { int x, y; };
static inline int2 make_int2(int x, int y) { return int2{x, y}; }
struct uint3 { unsigned x, y, z; };
struct dim3 { unsigned x = 1, y = 1, z = 1; };

namespace mokab_sim {
extern thread_local uint3 t_threadIdx, t_blockIdx;
extern thread_local dim3 t_blockDim, t_gridDim;
void sync_threads();
double shfl_down(double v, unsigned delta);
}  // namespace mokab_sim
#define threadIdx (::mokab_sim::t_threadIdx)
#define blockIdx (::mokab_sim::t_blockIdx)
#define blockDim (::mokab_sim::t_blockDim)
#define gridDim (::mokab_sim::t_gridDim)

template <class T> static inline T __ldg(const T *p) { return *p; }
// compiled with -ffp-contract=off and without -ffast-math: plain IEEE round-to-nearest operations, never fused
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline void __syncthreads() { ::mokab_sim::sync_threads(); }
static inline double __shfl_down_sync(unsigned, double v, unsigned delta) { return ::mokab_sim::shfl_down(v, delta); }

// ---- runtime API (the subset libmoka_b200 uses) ------------------------------------------------------------------
typedef enum cudaError {
    cudaSuccess = 0,
    cudaErrorInvalidValue = 1,
    cudaErrorMemoryAllocation = 2,
    cudaErrorNoDevice = 100,
    cudaErrorStreamCaptureUnsupported = 900,
    cudaErrorStreamCaptureInvalidated = 901,
    cudaErrorStreamCaptureUnjoined = 904,
    cudaErrorStreamCaptureIsolation = 905,
    cudaErrorLaunchFailure = 719,
    cudaErrorUnknown = 999
} cudaError_t;

typedef struct mokab_sim_stream *cudaStream_t;
typedef struct mokab_sim_event *cudaEvent_t;
typedef struct mokab_sim_graph *cudaGraph_t;
typedef struct mokab_sim_graph_exec *cudaGraphExec_t;

enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum cudaStreamCaptureMode { cudaStreamCaptureModeGlobal = 0, cudaStreamCaptureModeThreadLocal = 1, cudaStreamCaptureModeRelaxed = 2 };
#define cudaStreamNonBlocking 0x01
#define cudaEventDisableTiming 0x02
#define cudaHostAllocDefault 0x00

struct cudaIpcMemHandle_t { char reserved[64]; };
#define cudaIpcMemLazyEnablePeerAccess 0x01

struct cudaDeviceProp {
    char name[256];
    int major, minor, multiProcessorCount;
};

extern "C" {
const char *cudaGetErrorString(cudaError_t e);
cudaError_t cudaGetLastError(void);
cudaError_t cudaGetDeviceCount(int *n);
cudaError_t cudaSetDevice(int device);
cudaError_t cudaGetDeviceProperties(cudaDeviceProp *prop, int device);
cudaError_t cudaDeviceSynchronize(void);
cudaError_t cudaMalloc(void **p, size_t bytes);
cudaError_t cudaFree(void *p);
cudaError_t cudaHostAlloc(void **p, size_t bytes, unsigned flags);
cudaError_t cudaFreeHost(void *p);
cudaError_t cudaMemcpy(void *dst, const void *src, size_t bytes, cudaMemcpyKind kind);
cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t bytes, cudaMemcpyKind kind, cudaStream_t s);
cudaError_t cudaMemsetAsync(void *dst, int value, size_t bytes, cudaStream_t s);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned flags);
cudaError_t cudaStreamCreateWithPriority(cudaStream_t *s, unsigned flags, int priority);
cudaError_t cudaStreamDestroy(cudaStream_t s);
cudaError_t cudaStreamSynchronize(cudaStream_t s);
cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned flags);
cudaError_t cudaEventCreate(cudaEvent_t *e);
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned flags);
cudaError_t cudaEventDestroy(cudaEvent_t e);
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s);
cudaError_t cudaEventSynchronize(cudaEvent_t e);
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b);
cudaError_t cudaStreamBeginCapture(cudaStream_t s, cudaStreamCaptureMode mode);
cudaError_t cudaStreamEndCapture(cudaStream_t s, cudaGraph_t *g);
cudaError_t cudaGraphInstantiate(cudaGraphExec_t *ge, cudaGraph_t g, unsigned long long flags);
cudaError_t cudaGraphDestroy(cudaGraph_t g);
cudaError_t cudaGraphExecDestroy(cudaGraphExec_t ge);
cudaError_t cudaGraphLaunch(cudaGraphExec_t ge, cudaStream_t s);
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
cudaError_t cudaFuncSetAttribute(const void *func, enum cudaFuncAttribute attr, int value);
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *h, void *p);
cudaError_t cudaIpcOpenMemHandle(void **p, cudaIpcMemHandle_t h, unsigned flags);
cudaError_t cudaIpcCloseMemHandle(void *p);
}
template <class F> static inline cudaError_t cudaFuncSetAttribute(F *func, enum cudaFuncAttribute attr, int value)
{
    return cudaFuncSetAttribute((const void *)func, attr, value);
}
// the toolkit's C++ convenience overload
template <class T> static inline cudaError_t cudaMalloc(T **p, size_t bytes) { return cudaMalloc((void **)(void *)p, bytes); }

// ---- kernel launches: tests/sim/build.py rewrites  K<<<grid, block, smem, stream>>>(args...)  into
//      ::mokab_sim::launch(K, grid, block, stream, args...)  -- the arguments are converted to the kernel's parameter
//      types and stored at launch time (as the CUDA launch does), the body runs when the stream gets there ---------
#include <functional>
#include <tuple>
#include <type_traits>
#include <utility>

namespace mokab_sim {
// `coop`: the kernel uses __syncthreads / warp shuffles (its threads run as fibers); otherwise threads run to
// completion one after another and a call to either primitive aborts the launch with an error
void enqueue_kernel(cudaStream_t s, unsigned grid, unsigned block, size_t smem, bool coop, const char *name, std::function<void()> body);
// the dynamic shared memory of the block that is executing (size given at launch; one buffer per host thread, reused)
unsigned char *dynamic_smem();

static inline bool is_coop_name(const char *kernel)
{
    return __builtin_strstr(kernel, "reduce::k_") != nullptr || __builtin_strstr(kernel, "k_halo_push") != nullptr ||
           __builtin_strstr(kernel, "k_halo_wait_ll") != nullptr ||
           (__builtin_strstr(kernel, "k_rk_stage") != nullptr && __builtin_strstr(kernel, ", true>") != nullptr) ||   // the PUSH variant
           __builtin_strstr(kernel, "k_rk_stage_tma") != nullptr;                                                    // the TMA variant
}
// a device-side wait on memory that another stream (or rank) writes: retried by the scheduler until `ready` returns true
void enqueue_try(cudaStream_t s, const char *name, std::function<bool()> ready);

template <class... P, class... A>
static inline void launch_impl(bool coop, const char *name, void (*k)(P...), unsigned grid, unsigned block, size_t smem, cudaStream_t s,
                               A &&...a)
{
    std::tuple<std::decay_t<P>...> args(static_cast<P>(std::forward<A>(a))...);
    enqueue_kernel(s, grid, block, smem, coop, name, [k, args]() { std::apply(k, args); });
}
}  // namespace mokab_sim
