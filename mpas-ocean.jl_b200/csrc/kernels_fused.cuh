// kernels_fused.cuh -- the fused RungeKutta4 stage: tendencies + provisional state + accumulator in
// one pass over the mesh (the reference's intent at src/forward/time_integration.jl:112-137, which as
// written would be ~12 launches and 5 extra array round trips per stage).
//
// Per stage s, for every edge e and cell c (provisional state "old" = output of stage s-1):
//   kU[e] = -(g/dc[e]) * ((hOld[c2]-H[c2]) - (hOld[c1]-H[c1])) + sum_i wf[i,e] * uOld[eoe[i,e]]
//           (pressure_gradient.jl:63 and horizontal_advection_and_coriolis.jl:70-72, with
//            wf = weightsOnEdge * fEdge[eoe] folded at upload)
//   kH[c] = invArea[c] * sum_i sign[i,c] * dv[e_i] * uOld[e_i] * (hOld[c1(e_i)] + hOld[c2(e_i)])/2
//           (horizontal_advection.jl:64-65 with flux = u*hEdge, DiagnosticVars.jl:158-173, and
//            hEdge = interpolateCell2Edge of the SAME provisional state, Operators.jl:201-222)
//   stage 1:   out = cur + a*k ; acc  = cur + b*k        (acc never read)
//   stage 2,3: out = cur + a*k ; acc += b*k
//   stage 4:                      acc += b*k             (acc is the other time level: no copy-back)
// Neither tend*, thicknessFlux, layerThicknessEdge nor ssh ever touch HBM.
//
// DER = true (the default; MOKAB_MESH_EXPLICIT_EOE turns it off): edgesOnEdge is not read where it can be rebuilt.  MPAS orders edgesOnEdge[:, e] as "the other edges of cell 1
// in edgesOnCell order starting after e, then the same for cell 2"; mesh.cuh verifies that per edge and flags the
// blocks where it holds everywhere.  Those blocks read one position byte per edge and pick the ten indices out of the
// two edgesOnCell rows through L1 (the rows of a block's cells are 6 KB, resident after the first touch; the cell
// phase needs them anyway) -- 40 B of index traffic per edge (20 % of the Float64 stage's HBM bytes, 31 % in Float32) become 1 B.  The gather order, and with
// it every floating-point operation, is unchanged (bit-identical results, tested).  Measured on B200 at 2048x2048
// (profiles/README.md): +5.7 % (Float64, 2.81 vs 2.66 G cell-steps/s) and +10.6 % (Float32, 4.05 vs 3.66).  A first
// version that staged the rows in shared memory behind a block-wide barrier moved the same 16 % fewer DRAM bytes but
// ran SLOWER (2.11 G): the kernel is latency-bound, the barrier and the dependent shared-memory hop cost more than the
// bytes saved; reading the rows through L1 without a barrier is what made it pay.
//
// v1 data path: static connectivity / weights are slot-major and streamed with L1::no_allocate loads
// (read once per stage), the provisional state is gathered through L1/L2 -- after the Hilbert
// renumbering the 10+2 gathers of an edge and the 6x3 gathers of a cell land in lines its
// neighbours in the warp also touch.  A block owns TC consecutive cells and the edges they own,
// so the two halves of the block share their working set in L1.
#pragma once
#include "common.cuh"
#include "kernels_p2p.cuh"
#ifndef MOKAB_SIM
#include <cuda/barrier>
#endif

namespace mokab {
namespace fused {

constexpr int kTC = MOKAB_BLOCK_CELLS;       // cells per block
constexpr int kThreads = MOKAB_BLOCK_CELLS;

// Round-to-nearest multiply/add that the compiler may not contract into FMA: the fused kernel keeps the
// reference's operation order so that Float64 results are bit-identical to the oracle.
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }

// Float32 states carry the PERTURBATION of the layer thickness, h - restingThicknessSum (= ssh on this single-layer path), in
// their `h` arrays: around h ~ 1000 m a Float32 resolves 6e-5 m, which is 6e-5 of the 1 m wave signal and far from
// BASELINE.json's 1e-5 tolerance, while the perturbation itself keeps the full 24-bit significand for the pressure gradient
// and for the RK accumulation.  The whole thickness H + (h - H) is formed only where the thickness flux needs it.  Float64
// states keep the reference's variable (layerThickness) and operation order bit for bit.
template <class R> constexpr bool kPert = sizeof(R) == 4;

// Direct-store halo exchange folded into the BOUNDARY launch of a stage (PUSH = true; protocol and ordering argument in
// kernels_p2p.cuh, which holds the stand-alone variant): a block first waits until every sender's arrival counter has
// reached this rank's count of completed boundary launches, computes, stores each value a neighbour needs straight into
// that neighbour's array (CSR over the local entities: almost always empty), and the last block to finish bumps the
// expectations and ticks this rank's arrival counter on every receiver.  One launch per stage on the halo stream.
template <class R>
struct PushStage {
    const int32_t *startE, *startC;        // CSR over local edges / cells into dst / slot (size nE + 1 / nC + 1)
    const int32_t *dstE, *dstC;            // the entity's index in the receiver's arrays
    const uint8_t *slotE, *slotC;          // which receiver
    R *const *peerU, *const *peerH;        // per receiver: this stage's output arrays in its memory
    unsigned long long *const *arrivalAt;  // per receiver: this rank's arrival counter there
    int nrecv;
    const int32_t *senders;                // ranks this one receives from
    int nsend;
    const unsigned long long *arrival;     // local arrival counters, indexed by sender rank
    unsigned long long *expect;            // boundary launches completed so far, per sender rank
    unsigned int *done;
    int *error;
    long long timeout_cycles;
};

template <class R>
struct StageArgs {
    int nE, nC;              // local entity counts (= strides of the slot-major arrays)
    int nCown;               // cells computed by this rank (the rest are halo copies)
    const int32_t *blockList;  // block ids to run (interior / boundary part) or nullptr = all
    // static (see mesh.cuh)
    const int2 *ce;
    const int32_t *eoe;      // (S2, nE) absent -> self
    const int32_t *eoc;      // (S, nC)  (edge << 1) | (sign > 0)
    const uint8_t *nEoE, *nEoC;
    const int32_t *blkEdgeStart;
    const uint8_t *posE;       // per edge: position in the edgesOnCell row of cell 1 | (position in the row of cell 2) << 4
    const uint8_t *blkDerived; // per block: 1 = rebuild edgesOnEdge from edgesOnCell (mesh.cuh), nullptr = never
    const R *gdc, *wf, *dv, *invArea, *H;
    // dynamic
    const R *uOld, *hOld;    // provisional state the tendencies are evaluated at
    const R *uCur, *hCur;    // state at the start of the step
    R *uAcc, *hAcc;          // accumulator == the other time level
    R *uOut, *hOut;          // next provisional state (stages 1-3)
    R a, b;
    R f0;                    // uniform fEdge (FOLD = false): weights stay unfolded, (w*u)*f0 formed as the reference does;
                             // FOLD = true: wf already holds weightsOnEdge*fEdge[eoe] (variable f; differs by round-off)
    const PushStage<R> *push;  // PUSH launches only (device memory); nullptr otherwise
    int pf, pfDist;            // L2 prefetch of the streaming operands (moka_b200.cu: Options::stage_prefetch), distance in launched blocks
    int wStride;               // TMA = 1 launches only: elements between the staged weight rows in shared memory
    const R *wfI;              // TMA = 3: the weights again, slot-INTERLEAVED -- 16 bytes per edge and slot group (k_build_wf_interleaved)
    const R *wfB;              // TMA = 2: the weights again, BLOCK-major -- block b's S2 rows back to back, each padded to 16 bytes,
    const long long *wfBOff;   //          starting at element wfBOff[b] (16-byte aligned): one contiguous run, one bulk copy per block
#ifdef MOKAB_TRACE
    int traceKind;             // TRACE builds: stage | part << 4 (common.cuh)
#endif
};

// STAGE: 1 = first, 2 = middle (2 and 3), 4 = last.  S2/S: compile-time maxEdges2/maxEdges (0 = runtime).
#ifndef MOKAB_MINBLOCKS
#define MOKAB_MINBLOCKS 5   // <= 48 registers, 5 blocks (40 warps) per SM: measured best of 4/5/6 (profiles/README.md)
#endif
// resident blocks per SM of the edgesOnEdge-rebuilding variant: Float64 needs 64 registers to hold the ten weights
// across the index reconstruction without spilling (4 blocks), Float32 fits in 48 (5 blocks) -- measured r01h:
// F64 2.81 / 2.49 / 2.09 G cell-steps/s at 4 / 5 / 6 blocks, F32 3.62 / 4.05 / 3.83
#ifndef MOKAB_DER_MINBLOCKS_F64
#define MOKAB_DER_MINBLOCKS_F64 4
#endif
#ifndef MOKAB_DER_MINBLOCKS_F32
#define MOKAB_DER_MINBLOCKS_F32 5
#endif
// TMA = 3 (opt-in, "stage_tma" = 3): the Coriolis weights of an edge -- 80 of the ~100 bytes it streams in Float64 -- land in
// SHARED memory through per-thread asynchronous copies (cp.async.cg, 16 bytes = the weights of one edge for two (Float64) /
// four (Float32) consecutive slots of the slot-interleaved copy wfI) instead of in registers.  They are issued first, fly
// while the thread does its index loads and gathers, and are read back (conflict-free: every thread reads what it copied,
// no barrier) only where the weighted sum starts.  The ten weights no longer occupy twenty registers across the gather
// latency, so a Float64 thread fits in 51 registers (5 resident blocks instead of 4) and a Float32 thread in 42 (6 instead of
// 5): more warps per SM for a kernel that ncu shows latency-bound at 47 % occupancy (profiles/README.md r02b).
#ifndef MOKAB_CPA_MINBLOCKS_F64
#define MOKAB_CPA_MINBLOCKS_F64 5
#endif
#ifndef MOKAB_CPA_MINBLOCKS_F32
#define MOKAB_CPA_MINBLOCKS_F32 6
#endif
template <class R> __host__ __device__ constexpr int cpa_vec() { return 16 / (int)sizeof(R); }                      // weights per 16-byte copy
template <class R, int S2T> __host__ __device__ constexpr int cpa_groups() { return (S2T + cpa_vec<R>() - 1) / cpa_vec<R>(); }
__device__ __forceinline__ void cp_async_16(void *smem_dst, const void *gmem_src)
{
#ifdef MOKAB_SIM
    memcpy(smem_dst, gmem_src, 16);
#else
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
#endif
}
__device__ __forceinline__ void cp_async_wait_all()
{
#ifndef MOKAB_SIM
    asm volatile("cp.async.wait_all;" ::: "memory");
#endif
}

template <class R> constexpr int der_minblocks() { return sizeof(R) == 8 ? MOKAB_DER_MINBLOCKS_F64 : MOKAB_DER_MINBLOCKS_F32; }
// resident blocks of kThreads threads the stage kernel is compiled for (MOKAB_DER_RAW_BLOCKS_*: tuning builds that want a
// count the 256-thread scale cannot express, e.g. nine blocks of 128 threads)
template <class R, bool DER, int TMA> constexpr int stage_blocks()
{
#ifdef MOKAB_DER_RAW_BLOCKS_F64
    if (DER && TMA == 0 && sizeof(R) == 8) return MOKAB_DER_RAW_BLOCKS_F64;
#endif
#ifdef MOKAB_DER_RAW_BLOCKS_F32
    if (DER && TMA == 0 && sizeof(R) == 4) return MOKAB_DER_RAW_BLOCKS_F32;
#endif
    if (TMA == 3) return MOKAB_BLOCKS_SCALED(sizeof(R) == 8 ? MOKAB_CPA_MINBLOCKS_F64 : MOKAB_CPA_MINBLOCKS_F32);
    return MOKAB_BLOCKS_SCALED(TMA != 0 ? 3 : DER ? der_minblocks<R>() : MOKAB_MINBLOCKS);
}
// TMA = 1 (opt-in, MOKAB_STAGE_TMA=1; compile-time row widths only): the Coriolis weights of the block's edges -- ten
// contiguous runs, 80 of the ~160 streamed bytes per edge in Float64 -- are fetched by ONE thread with bulk asynchronous copies
// (cp.async.bulk, the 1-D TMA path) into shared memory behind an mbarrier, at the very top of the kernel; the threads meanwhile
// issue their index loads and gathers and wait on the barrier only where the first weighted sum starts.  The DRAM latency of
// the bulk of the bytes is thereby decoupled from the register file, which is what limits the resident warps of the plain
// kernel.  A bulk copy needs 16-byte aligned addresses and sizes: each row is fetched from the aligned address below its
// first element (`shift` elements early) and rounded up, the arrays carry a few elements of padding at the end.
// TMA = 2 (MOKAB_STAGE_TMA=2): the same over a block-major copy of the weights (wfB) -- the block's rows are one contiguous,
// aligned run, ONE bulk copy, no shifts.
template <class R> __host__ __device__ constexpr int tma_align() { return 16 / (int)sizeof(R); }
#ifdef MOKAB_SIM
#define MOKAB_DYN_SMEM(name) unsigned char *name = ::mokab_sim::dynamic_smem()
#else
#define MOKAB_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif

// Programmatic dependent launch (opt-in, "stage_pdl"): consecutive stage launches of one stream overlap the TAIL of stage s with
// the static half of stage s + 1.  Every block lets the next launch start as soon as it is resident itself
// (griddepcontrol.launch_dependents at entry), and waits for the previous launch to have completed and flushed
// (griddepcontrol.wait) only where it first touches the state -- its connectivity, metrics and weight copies (the L2
// prefetches, the cp.async weights, cellsOnEdge, g/dc: ~60 % of the bytes of its first edge iteration) are issued before
// that.  Both instructions are no-ops in a launch without the programmatic attribute.
// Loads of the STATE (the output of the previous launch) are plain coherent loads, not the read-only path (__ldg): a
// programmatically dependent launch ("stage_pdl") starts before the previous one has finished, so lines of a buffer that an
// older, still running launch reads can enter L1 AFTER this launch's start-of-grid invalidation and be stale once the buffer
// has been rewritten and this launch reads it.  Seen on hardware (r02h: wrong results on a 3-block mesh with __ldg);
// griddepcontrol.wait orders plain loads.  Measured cost of the change: none (r02i: 2.864 vs 2.851 G at 2048 x 2048, 4.281 vs
// 4.270 G in Float32).  MOKAB_STATE_LOADS_LDG builds the read-only variant for A/B; it ignores "stage_pdl".
template <class T>
__device__ __forceinline__ T ld_state(const T *p)
{
#ifdef MOKAB_STATE_LOADS_LDG
    return __ldg(p);
#else
    return *p;
#endif
}
__device__ __forceinline__ void pdl_launch_dependents()
{
#ifndef MOKAB_SIM
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
__device__ __forceinline__ void pdl_wait()
{
#ifndef MOKAB_SIM
    asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}

template <class R, int STAGE, int S2T, int ST, bool FOLD, bool DER, bool PUSH = false, int TMA = 0>
__global__ void __launch_bounds__(kThreads, (stage_blocks<R, DER, TMA>()))
k_rk_stage(const StageArgs<R> A, int S2rt, int Srt)
{
    pdl_launch_dependents();
    MOKAB_TRACE_BEGIN();
    const int S2 = S2T ? S2T : S2rt;
    const int S = ST ? ST : Srt;
    const int nE = A.nE, nC = A.nC;
    const int b = A.blockList ? A.blockList[blockIdx.x] : blockIdx.x;
    const int cBase = b * kTC;
    static_assert(TMA == 0 || (S2T != 0 && ST != 0), "the TMA variants stage compile-time many weight rows");
    [[maybe_unused]] unsigned char *cpa_slot = nullptr;             // TMA = 3: this thread's 16-byte slots, one per slot group, kThreads * 16 bytes apart
    if constexpr (TMA == 3) {
        MOKAB_DYN_SMEM(dyn3);
        cpa_slot = dyn3 + (size_t)threadIdx.x * 16;
    }
    [[maybe_unused]] const R *sw = nullptr;
    [[maybe_unused]] bool weights_landed = false;
#ifndef MOKAB_SIM
    [[maybe_unused]] typename cuda::barrier<cuda::thread_scope_block>::arrival_token tma_token;
    [[maybe_unused]] cuda::barrier<cuda::thread_scope_block> *tma_bar = nullptr;
#endif
    if constexpr (TMA == 1 || TMA == 2) {
        MOKAB_DYN_SMEM(dyn);
        R *dst = reinterpret_cast<R *>(dyn);
        sw = dst;
        constexpr int AL = tma_align<R>();
        const int eb0 = A.blkEdgeStart[b], nb = A.blkEdgeStart[b + 1] - eb0;
        [[maybe_unused]] const int nbp = (nb + AL - 1) / AL * AL;              // TMA = 2: padded row length of this block
#ifdef MOKAB_SIM
        if (threadIdx.x == 0) {
            if constexpr (TMA == 2) {
                for (size_t k = 0; k < (size_t)S2T * nbp; ++k) dst[k] = A.wfB[A.wfBOff[b] + k];
            } else {
                for (int i = 0; i < S2T; ++i) {
                    const size_t g0 = (size_t)i * nE + eb0, a0 = g0 & ~(size_t)(AL - 1);
                    const size_t cnt = (g0 - a0 + nb + AL - 1) / AL * AL;
                    for (size_t k = 0; k < cnt; ++k) dst[(size_t)i * A.wStride + k] = A.wf[a0 + k];
                }
            }
        }
        __syncthreads();
        weights_landed = true;
#else
#pragma nv_diag_suppress static_var_with_dynamic_init
        __shared__ cuda::barrier<cuda::thread_scope_block> bar;
        tma_bar = &bar;
        if (threadIdx.x == 0) {
            init(&bar, blockDim.x);
            cuda::device::experimental::fence_proxy_async_shared_cta();
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned int bytes = 0;
            if constexpr (TMA == 2) {
                bytes = (unsigned int)((size_t)S2T * nbp * sizeof(R));
                if (bytes) cuda::device::experimental::cp_async_bulk_global_to_shared(dst, A.wfB + A.wfBOff[b], bytes, bar);
            } else
            for (int i = 0; i < S2T; ++i) {
                const size_t g0 = (size_t)i * nE + eb0, a0 = g0 & ~(size_t)(AL - 1);
                const unsigned int cnt = (unsigned int)((g0 - a0 + nb + AL - 1) / AL * AL);
                if (cnt == 0) continue;                                      // a block that owns no edge
                cuda::device::experimental::cp_async_bulk_global_to_shared(dst + (size_t)i * A.wStride, A.wf + a0, cnt * (unsigned int)sizeof(R), bar);
                bytes += cnt * (unsigned int)sizeof(R);
            }
            tma_token = cuda::device::barrier_arrive_tx(bar, 1, bytes);
        } else {
            tma_token = bar.arrive();
        }
#endif
    }
    if constexpr (PUSH) {
        pdl_wait();
#ifndef MOKAB_SIM   // (the simulated runtime cannot spin inside a kernel: the host enqueues the same predicate before the launch)
        const PushStage<R> &P = *A.push;
        if ((int)threadIdx.x < P.nsend) {
            const int q = P.senders[threadIdx.x];
            const long long t0 = clock64();
            while (p2p::load_acquire_system(P.arrival + q) < P.expect[q]) {
                if (*(volatile int *)P.error || clock64() - t0 > P.timeout_cycles) {
                    atomicExch(P.error, 1);
                    break;
                }
                __nanosleep(64);
            }
        }
#endif
        __syncthreads();
        MOKAB_TRACE_MARK();
    }

    constexpr bool kDer = DER && ST != 0 && S2T != 0;
    const bool derived = kDer && A.blkDerived && A.blkDerived[b];

    // ---- L2 prefetch of the streaming operands (opt-in, A.pf) ----------------------------------------------------------------
    // The kernel is bound by dependent DRAM latencies (stream -> index -> gather) at the occupancy its registers allow, not by
    // bytes.  `prefetch.global.L2` costs no register and no shared memory: (bit 0) at entry a block pulls the streams of its
    // own second and third edge iteration and of its cell phase into L2 while the first iteration waits on DRAM, so three of
    // its four streaming hops become L2 hits; (bit 1) it does the same for the whole working set of the block that is
    // launched A.pfDist blocks later -- about one wave of resident blocks, i.e. the block that takes over this block's slot.
    if constexpr (S2T != 0 && !PUSH) {
        if (A.pf) {
            auto pf_edges = [&](int ea, int eb, bool far) {
                for (int e = ea + (int)threadIdx.x; e < eb; e += kThreads) {
                    prefetch_l2(A.ce + e);
                    prefetch_l2(A.gdc + e);
                    if (kDer && A.blkDerived) prefetch_l2(A.posE + e);
                    else {
#pragma unroll
                        for (int i = 0; i < S2T; ++i) prefetch_l2(A.eoe + (size_t)i * nE + e);
                    }
                    if constexpr (TMA == 0) {
#pragma unroll
                        for (int i = 0; i < S2T; ++i) prefetch_l2(A.wf + (size_t)i * nE + e);
                    }
                    if (STAGE != 4) prefetch_l2(A.uCur + e);
                    if (STAGE != 1) prefetch_l2(A.uAcc + e);
                    if (far && STAGE != 1) prefetch_l2(A.uOld + e);      // most of the gathers of a block land in its own index range
                }
            };
            auto pf_cells = [&](int c, bool far) {
                if (c >= A.nCown) return;
#pragma unroll
                for (int i = 0; i < ST; ++i) prefetch_l2(A.eoc + (size_t)i * nC + c);
                prefetch_l2(A.invArea + c);
                prefetch_l2(A.H + c);
                if (STAGE != 4) prefetch_l2(A.hCur + c);
                if (STAGE != 1) prefetch_l2(A.hAcc + c);
                if (far && STAGE != 1) prefetch_l2(A.hOld + c);
            };
            if (A.pf & 1) {
                pf_edges(A.blkEdgeStart[b] + kThreads, A.blkEdgeStart[b + 1], false);
                pf_cells(cBase + (int)threadIdx.x, false);
            }
            if ((A.pf & 2) && blockIdx.x + (unsigned)A.pfDist < gridDim.x) {
                const int bn = A.blockList ? A.blockList[blockIdx.x + A.pfDist] : (int)blockIdx.x + A.pfDist;
                const int ea = A.blkEdgeStart[bn], eb = A.blkEdgeStart[bn + 1];
                pf_edges(ea, ((A.pf & 1) && eb > ea + kThreads) ? ea + kThreads : eb, true);
                pf_cells(bn * kTC + (int)threadIdx.x, true);
            }
        }
    }

    // ---- edges owned by this block's cells ----------------------------------------------------------
    const int e0 = A.blkEdgeStart[b], e1 = A.blkEdgeStart[b + 1];
    for (int e = e0 + threadIdx.x; e < e1; e += kThreads) {
        // (fetching cellsOnEdge / the position byte of the thread's next edge one iteration ahead was measured slower:
        //  2.68 vs 2.81 G cell-steps/s, the extra live registers spill -- profiles/README.md r01j)
        const int2 c = ld_stream(A.ce + e);
        const unsigned ppCur = (kDer && derived) ? ld_stream(A.posE + e) : 0u;
        R k;
        if constexpr (S2T != 0) {
            // Every streaming load of this edge is issued before anything waits on one of them: the padded slots of
            // the slot-major arrays are defined (index = the edge itself, weight = 0; k_build_fused_edges), so neither
            // the index nor the weight loads depend on nEdgesOnEdge, and one DRAM latency covers them all.
            int idx[S2T ? S2T : 1];
            R w[S2T ? S2T : 1];
            if constexpr (TMA == 3) {
                constexpr int NG = cpa_groups<R, S2T>();
#pragma unroll
                for (int gq = 0; gq < NG; ++gq)
                    cp_async_16(cpa_slot + (size_t)gq * kThreads * 16, A.wfI + ((size_t)gq * nE + e) * cpa_vec<R>());
            }
            const unsigned pp = ppCur;
            if (!(kDer && derived)) {
#pragma unroll
                for (int i = 0; i < S2T; ++i) idx[i] = ld_stream(A.eoe + (size_t)i * nE + e);
            }
            if constexpr (TMA == 0) {
#ifdef MOKAB_ENABLE_WF_BLOCK_MAJOR   // tuning build (libmoka_b200_wfb.so; the run-time branch costs the default kernel a spill)
                if (A.wfB) {    // "stage_wf_block_major": the block's S2T weight rows back to back (one contiguous ~60 KB run per block)
                    constexpr int AL = tma_align<R>();
                    const int nbp = (e1 - e0 + AL - 1) / AL * AL;
                    const R *wb = A.wfB + A.wfBOff[b] + (e - e0);
#pragma unroll
                    for (int i = 0; i < S2T; ++i) w[i] = ld_stream(wb + (size_t)i * nbp);
                } else
#endif
                {
#pragma unroll
                    for (int i = 0; i < S2T; ++i) w[i] = ld_stream(A.wf + (size_t)i * nE + e);
                }
            }
            const R g = ld_stream(A.gdc + e);
            pdl_wait();                                                 // everything above is static; the state comes next
            // the RK operands are independent of the tendency: issue their loads now (they may alias the stores
            // below, so the compiler cannot hoist them itself)
            const R cur = (STAGE == 4) ? R(0) : A.uCur[e];
            const R accIn = (STAGE == 1) ? R(0) : A.uAcc[e];
            const R h1 = ld_state(A.hOld + c.x), h2 = ld_state(A.hOld + c.y);
            const R H1 = kPert<R> ? R(0) : __ldg(A.H + c.x), H2 = kPert<R> ? R(0) : __ldg(A.H + c.y);
            if (kDer && derived) {
                // posE: bits 0-2 position of e in the row of cell 1, bits 3-5 in the row of cell 2, bit 7 = both rows
                // have ST entries and the edge is not masked (the branch-free common case)
                const int p1 = pp & 7, p2 = (pp >> 3) & 7;
                if (pp & 128u) {
                    constexpr int H = ST - 1;
#pragma unroll
                    for (int i = 0; i < S2T; ++i) {
                        int r = (i < H ? p1 : p2) + 1 + (i < H ? i : i - H);
                        r -= r >= ST ? ST : 0;
                        idx[i] = i < 2 * H ? (__ldg(A.eoc + (size_t)r * nC + (i < H ? c.x : c.y)) >> 1) : e;
                    }
                } else {
                    const int n1 = __ldg(A.nEoC + c.x);
                    const int n2 = c.x == c.y ? 1 : (int)__ldg(A.nEoC + c.y);
#pragma unroll
                    for (int i = 0; i < S2T; ++i) {
                        int id = e;
                        if (i < n1 - 1) {
                            int r = p1 + 1 + i;
                            r -= r >= n1 ? n1 : 0;
                            id = __ldg(A.eoc + (size_t)r * nC + c.x) >> 1;
                        } else if (i < n1 + n2 - 2) {
                            int r = p2 + 1 + i - (n1 - 1);
                            r -= r >= n2 ? n2 : 0;
                            id = __ldg(A.eoc + (size_t)r * nC + c.y) >> 1;
                        }
                        idx[i] = id;
                    }
                }
            }
            R uu[S2T ? S2T : 1];
#pragma unroll
            for (int i = 0; i < S2T; ++i) uu[i] = ld_state(A.uOld + idx[i]);
            if constexpr (TMA == 3) {   // this thread's own copies have landed: read them back (16-byte shared-memory loads, no bank conflicts)
                cp_async_wait_all();
                constexpr int NG = cpa_groups<R, S2T>(), V = cpa_vec<R>();
#pragma unroll
                for (int gq = 0; gq < NG; ++gq) {
                    const R *q = reinterpret_cast<const R *>(cpa_slot + (size_t)gq * kThreads * 16);
#pragma unroll
                    for (int j = 0; j < V; ++j)
                        if (gq * V + j < S2T) w[gq * V + j] = q[j];
                }
            }
            if constexpr (TMA == 1 || TMA == 2) {   // the staged rows: wait for the bulk copies once, then plain shared-memory reads (conflict-free: edge-major)
#ifndef MOKAB_SIM
                if (!weights_landed) {
                    tma_bar->wait(std::move(tma_token));
                    weights_landed = true;
                }
#endif
                constexpr int AL = tma_align<R>();
                if constexpr (TMA == 2) {
                    const int nbp = (e1 - e0 + AL - 1) / AL * AL;
#pragma unroll
                    for (int i = 0; i < S2T; ++i) w[i] = sw[i * nbp + (e - e0)];
                } else {
#pragma unroll
                    for (int i = 0; i < S2T; ++i) w[i] = sw[(size_t)i * A.wStride + (((size_t)i * nE + e0) & (size_t)(AL - 1)) + (e - e0)];
                }
            }
            // tend = 0 - (g/dc)*(ssh2 - ssh1), then += (w*u)*f slot by slot (pressure_gradient.jl:63, coriolis :70-72)
            k = kPert<R> ? -mul_rn(g, add_rn(h2, -h1)) : -mul_rn(g, add_rn(add_rn(h2, -H2), -add_rn(h1, -H1)));
#pragma unroll
            for (int i = 0; i < S2T; ++i) k = add_rn(k, FOLD ? mul_rn(w[i], uu[i]) : mul_rn(mul_rn(w[i], uu[i]), A.f0));
            if (STAGE != 4) A.uOut[e] = add_rn(cur, mul_rn(A.a, k));       // Provis = Curr + a*tend (time_integration.jl:124)
            if (STAGE == 1) A.uAcc[e] = add_rn(cur, mul_rn(A.b, k));       // New = Curr + b1*tend    (:108-110, :134)
            else            A.uAcc[e] = add_rn(accIn, mul_rn(A.b, k));     // New += b*tend           (:134)
            if constexpr (PUSH) {   // this stage's output value, straight into the arrays of whoever holds this edge as a halo copy
                const PushStage<R> &P = *A.push;
                const R v = (STAGE != 4) ? add_rn(cur, mul_rn(A.a, k)) : add_rn(accIn, mul_rn(A.b, k));
                for (int j = P.startE[e]; j < P.startE[e + 1]; ++j) P.peerU[P.slotE[j]][P.dstE[j]] = v;
            }
            continue;
        }
        const int n = ld_stream(A.nEoE + e);
        pdl_wait();
        const R cur = (STAGE == 4) ? R(0) : A.uCur[e];
        const R accIn = (STAGE == 1) ? R(0) : A.uAcc[e];
        const R h1 = ld_state(A.hOld + c.x), h2 = ld_state(A.hOld + c.y);
        const R H1 = kPert<R> ? R(0) : __ldg(A.H + c.x), H2 = kPert<R> ? R(0) : __ldg(A.H + c.y);
        k = kPert<R> ? -mul_rn(ld_stream(A.gdc + e), add_rn(h2, -h1))
                     : -mul_rn(ld_stream(A.gdc + e), add_rn(add_rn(h2, -H2), -add_rn(h1, -H1)));
        for (int i = 0; i < n; ++i) {
            const R wu = mul_rn(ld_stream(A.wf + (size_t)i * nE + e), ld_state(A.uOld + ld_stream(A.eoe + (size_t)i * nE + e)));
            k = add_rn(k, FOLD ? wu : mul_rn(wu, A.f0));
        }
        if (STAGE != 4) A.uOut[e] = add_rn(cur, mul_rn(A.a, k));       // Provis = Curr + a*tend (time_integration.jl:124)
        if (STAGE == 1) A.uAcc[e] = add_rn(cur, mul_rn(A.b, k));       // New = Curr + b1*tend    (:108-110, :134)
        else            A.uAcc[e] = add_rn(accIn, mul_rn(A.b, k));     // New += b*tend           (:134)
        if constexpr (PUSH) {
            const PushStage<R> &P = *A.push;
            const R v = (STAGE != 4) ? add_rn(cur, mul_rn(A.a, k)) : add_rn(accIn, mul_rn(A.b, k));
            for (int j = P.startE[e]; j < P.startE[e + 1]; ++j) P.peerU[P.slotE[j]][P.dstE[j]] = v;
        }
    }

    // ---- cells of this block ----------------------------------------------------------------------------
    const int cc = cBase + threadIdx.x;
    pdl_wait();                                                         // (a block that owns no edge has not waited yet)
    if (cc < A.nCown) {
        const int n = ld_stream(A.nEoC + cc);
        // Float32: the arrays hold the perturbation h - H (see kPert); the flux needs the whole thickness
        const R hc = kPert<R> ? add_rn(ld_state(A.hOld + cc), __ldg(A.H + cc)) : ld_state(A.hOld + cc);
        const R cur = (STAGE == 4) ? R(0) : A.hCur[cc];
        const R accIn = (STAGE == 1) ? R(0) : A.hAcc[cc];
        R acc = R(0);
        if constexpr (ST != 0) {
            int ee[ST ? ST : 1];
#pragma unroll
            for (int i = 0; i < ST; ++i) ee[i] = i < n ? (kDer ? __ldg(A.eoc + (size_t)i * nC + cc) : ld_stream(A.eoc + (size_t)i * nC + cc)) : -1;
            int2 cs[ST ? ST : 1];
            R uu[ST ? ST : 1], dd[ST ? ST : 1];
#pragma unroll
            for (int i = 0; i < ST; ++i) {
                const int e = ee[i] >= 0 ? (ee[i] >> 1) : 0;
                cs[i] = __ldg(A.ce + e);
                uu[i] = ld_state(A.uOld + e);
                dd[i] = __ldg(A.dv + e);
            }
            const R invA = ld_stream(A.invArea + cc);
#pragma unroll
            for (int i = 0; i < ST; ++i) {
                const int other = cs[i].x == cc ? cs[i].y : cs[i].x;
                const R ho = kPert<R> ? add_rn(ld_state(A.hOld + other), __ldg(A.H + other)) : ld_state(A.hOld + other);
                // flux = u*hEdge (DiagnosticVars.jl:158-173), hEdge = 0.5*(h1+h2) (Operators.jl:201-222),
                // tend += flux*dv*sign*invArea (horizontal_advection.jl:64-65)
                const R f = mul_rn(mul_rn(mul_rn(uu[i], mul_rn(R(0.5), add_rn(hc, ho))), dd[i]), invA);
                if (ee[i] >= 0) acc = add_rn(acc, (ee[i] & 1) ? f : -f);
            }
        } else {
            const R invA = ld_stream(A.invArea + cc);
            for (int i = 0; i < n; ++i) {
                const int ex = ld_stream(A.eoc + (size_t)i * nC + cc);
                const int e = ex >> 1;
                const int2 cs = __ldg(A.ce + e);
                const int other = cs.x == cc ? cs.y : cs.x;
                const R ho = kPert<R> ? add_rn(ld_state(A.hOld + other), __ldg(A.H + other)) : ld_state(A.hOld + other);
                const R f = mul_rn(mul_rn(mul_rn(ld_state(A.uOld + e), mul_rn(R(0.5), add_rn(hc, ho))), __ldg(A.dv + e)), invA);
                acc = add_rn(acc, (ex & 1) ? f : -f);
            }
        }
        const R k = acc;
        if (STAGE != 4) A.hOut[cc] = add_rn(cur, mul_rn(A.a, k));
        if (STAGE == 1) A.hAcc[cc] = add_rn(cur, mul_rn(A.b, k));
        else            A.hAcc[cc] = add_rn(accIn, mul_rn(A.b, k));
        if constexpr (PUSH) {
            const PushStage<R> &P = *A.push;
            const R v = (STAGE != 4) ? add_rn(cur, mul_rn(A.a, k)) : add_rn(accIn, mul_rn(A.b, k));
            for (int j = P.startC[cc]; j < P.startC[cc + 1]; ++j) P.peerH[P.slotC[j]][P.dstC[j]] = v;
        }
    }
    if constexpr (PUSH) {   // one fence per block after the barrier (kernels_p2p.cuh), the last block to finish opens the next launch's gate and ticks the receivers
        const PushStage<R> &P = *A.push;
        __syncthreads();
        if (threadIdx.x == 0) {
            p2p::fence_system();
            const unsigned int ticket = p2p::add_device(P.done, 1u);
            if (ticket == gridDim.x - 1) {
                *P.done = 0u;
                for (int i = 0; i < P.nsend; ++i) P.expect[P.senders[i]] += 1ull;
                p2p::fence_system();
                for (int i = 0; i < P.nrecv; ++i) p2p::add_system(P.arrivalAt[i], 1ull);
            }
        }
    }
#ifdef MOKAB_TRACE
    MOKAB_TRACE_END((unsigned)A.traceKind);
#endif
}

// ---- fused RungeKutta4 stage for MULTI-LEVEL states (nVertLevels = K > 1) ------------------------------------------------
// The reference's kernels carry `for k in 1:maxLevelEdgeTop[iEdge]` (pressure_gradient.jl:61-64, horizontal_advection_and_
// coriolis.jl:69-73, horizontal_advection.jl:60-66) around arithmetic that is independent per level except for the free
// surface: one pressure gradient, -g/dc (ssh2 - ssh1), for the whole column.  State arrays are level-major (every level a
// contiguous array over the entities), so a thread loads its edge's / cell's connectivity, weights and metrics ONCE and walks
// the column with them in registers: the static bytes -- 355 of the 483 B a single-level stage moves per cell -- are
// amortised over K levels (K = 10: 1 635 B per cell and stage for ten levels, 164 B per level).  ssh of a provisional state is
// produced by the cell phase of the stage that writes it (sshOut = (h[0] + h[1] + ...) - H, level order; project-defined for
// K > 1) and gathered by the next stage's edge phase; with K = 1 every value equals the single-level
// kernel's bit for bit (tested).  Float64, explicit edgesOnEdge, whole mesh (no halo parts).
struct StageArgsML {
    int nE, nC, K;
    int nCown;               // cells computed by this rank (the rest are halo copies: decomposed meshes)
    const int2 *ce;
    const int32_t *eoe;      // (S2, nE) absent -> self
    const int32_t *eoc;      // (S, nC)  (edge << 1) | (sign > 0)
    const uint8_t *nEoE, *nEoC;
    const int32_t *blkEdgeStart;
    const double *gdc, *wf, *dv, *invArea, *H;
    const double *uOld, *hOld, *sshOld;   // provisional state the tendencies are evaluated at (K levels; ssh one)
    const double *uCur, *hCur;            // state at the start of the step
    double *uAcc, *hAcc;                  // accumulator == the other time level
    double *uOut, *hOut, *sshOut;         // next provisional state (stages 1-3); stage 4: sshOut = ssh of the new state
    double a, b, f0;
};

template <int STAGE, int S2T, int ST, bool FOLD>
#ifndef MOKAB_ML_MINBLOCKS
#define MOKAB_ML_MINBLOCKS 2   // <= 128 registers: the column loop keeps a row of indices, weights and metrics live (3 blocks spill)
#endif
__global__ void __launch_bounds__(kThreads, MOKAB_BLOCKS_SCALED(MOKAB_ML_MINBLOCKS))
k_rk_stage_ml(const StageArgsML A, int S2rt, int Srt)
{
    const int S2 = S2T ? S2T : S2rt, S = ST ? ST : Srt;
    const int nE = A.nE, nC = A.nC, K = A.K;
    const int b = blockIdx.x;
    constexpr int MAXS2 = S2T ? S2T : 32, MAXS = ST ? ST : 16;       // mesh_create admits maxEdges2 <= 32, maxEdges <= 16
    const int e0 = A.blkEdgeStart[b], e1 = A.blkEdgeStart[b + 1];
    for (int e = e0 + threadIdx.x; e < e1; e += kThreads) {
        const int2 c = ld_stream(A.ce + e);
        int idx[MAXS2];
        double w[MAXS2];
        const int n = S2T ? S2T : (int)ld_stream(A.nEoE + e);
#pragma unroll
        for (int i = 0; i < MAXS2; ++i)
            if (i < S2) { idx[i] = ld_stream(A.eoe + (size_t)i * nE + e); w[i] = ld_stream(A.wf + (size_t)i * nE + e); }
        const double g = ld_stream(A.gdc + e);
        // one pressure gradient for the column: 0 - (g/dc) (ssh2 - ssh1)   (pressure_gradient.jl:63)
        const double p = -mul_rn(g, add_rn(__ldg(A.sshOld + c.y), -__ldg(A.sshOld + c.x)));
        for (int k = 0; k < K; ++k) {
            const size_t o = (size_t)k * nE;
            double t = p;
#pragma unroll
            for (int i = 0; i < MAXS2; ++i)
                if (i < n) {
                    const double wu = mul_rn(w[i], __ldg(A.uOld + o + idx[i]));
                    t = add_rn(t, FOLD ? wu : mul_rn(wu, A.f0));
                }
            const double cur = (STAGE == 4) ? 0.0 : A.uCur[o + e];
            const double accIn = (STAGE == 1) ? 0.0 : A.uAcc[o + e];
            if (STAGE != 4) A.uOut[o + e] = add_rn(cur, mul_rn(A.a, t));
            if (STAGE == 1) A.uAcc[o + e] = add_rn(cur, mul_rn(A.b, t));
            else            A.uAcc[o + e] = add_rn(accIn, mul_rn(A.b, t));
        }
    }
    const int cc = b * kTC + threadIdx.x;
    if (cc < A.nCown) {
        const int n = ld_stream(A.nEoC + cc);
        int ed[MAXS], other[MAXS];
        double dd[MAXS];
        bool pos[MAXS];
#pragma unroll
        for (int i = 0; i < MAXS; ++i)
            if (i < S) {
                const int ex = i < n ? ld_stream(A.eoc + (size_t)i * nC + cc) : -1;
                ed[i] = ex >= 0 ? (ex >> 1) : 0;
                pos[i] = ex >= 0 && (ex & 1);
                const int2 cs = __ldg(A.ce + ed[i]);
                other[i] = cs.x == cc ? cs.y : cs.x;
                dd[i] = __ldg(A.dv + ed[i]);
            }
        const double invA = ld_stream(A.invArea + cc);
        double col = 0.0;
        for (int k = 0; k < K; ++k) {
            const size_t oe = (size_t)k * nE, oc = (size_t)k * nC;
            const double hc = __ldg(A.hOld + oc + cc);
            double acc = 0.0;
#pragma unroll
            for (int i = 0; i < MAXS; ++i)
                if (i < n) {
                    const double f = mul_rn(mul_rn(mul_rn(__ldg(A.uOld + oe + ed[i]), mul_rn(0.5, add_rn(hc, __ldg(A.hOld + oc + other[i])))), dd[i]), invA);
                    acc = add_rn(acc, pos[i] ? f : -f);
                }
            const double cur = (STAGE == 4) ? 0.0 : A.hCur[oc + cc];
            const double accIn = (STAGE == 1) ? 0.0 : A.hAcc[oc + cc];
            const double out = add_rn(cur, mul_rn(A.a, acc));
            const double accOut = (STAGE == 1) ? add_rn(cur, mul_rn(A.b, acc)) : add_rn(accIn, mul_rn(A.b, acc));
            if (STAGE != 4) A.hOut[oc + cc] = out;
            A.hAcc[oc + cc] = accOut;
            const double v = (STAGE != 4) ? out : accOut;            // the thickness whose free surface the next stage / the caller reads
            col = k == 0 ? v : add_rn(col, v);
        }
        A.sshOut[cc] = add_rn(col, -ld_stream(A.H + cc));
    }
}

// ---- fused ForwardEuler step (the reference's live stepper) ----------------------------------------------------
// One kernel per step instead of the reference's 16 launches + 3 full-state copies (time_integration.jl:150-193,
// SURVEY.md section 2a), bit-identical to the reference-order kernel sequence INCLUDING its ordering artefact:
//   flux_n = u_n * hEdge_{n-1}   (diagnostic_compute! forms the flux before it refreshes layerThicknessEdge,
//                                 DiagnosticVars.jl:112-116; hEdge starts as zeros)
//   tendU  = 0 - (g/dc)(ssh2 - ssh1) + sum_i (w_i * u[eoe_i]) * f[eoe_i]      (normalVelocity.jl:21-53)
//   tendH  = sum_i ((flux[e_i] * dv[e_i]) * sign_i) * invArea                 (layerThickness.jl:14-28)
//   u' = u + dt tendU ; h' = h + dt tendH ; ssh' = h' - H ; hEdge_n = (h[c1] + h[c2]) / 2 of the OLD h
// The new state goes to the other time level (so "previous" holds the old state exactly as advanceTimeLevels! leaves
// it, without copying) and hEdge ping-pongs with it.  thicknessFlux, velocityDivCell and the tendency arrays are not
// written per step: the old state and the old hEdge stay resident, and the host side re-creates those arrays with the
// reference-order kernels when somebody asks for them (moka_b200.cu: fe_materialize).  relativeVorticity accumulates
// over the steps in the reference (Operators.jl:135), so on meshes with vertex arrays the curl kernel still runs
// every step.
struct FeArgs {
    int nE, nC, nCown;
    const int2 *ce;
    const int32_t *eoe;        // (S2, nE) absent -> self
    const int32_t *eoc;        // (S, nC)  (edge << 1) | (sign > 0)
    const uint8_t *nEoC;
    const int32_t *blkEdgeStart;
    const double *gdc, *woe, *fE, *dv, *invArea, *H;
    const double *u, *h, *ssh, *hEold;
    double *uNew, *hNew, *sshNew, *hEnew;
    double dt, f0;
    const int32_t *blockList;  // LIST launches (domain-decomposed runs: interior / boundary blocks): the blocks to process
};

template <int S2T, int ST, bool UNIF, bool LIST = false>
__global__ void __launch_bounds__(kThreads, MOKAB_BLOCKS_SCALED(4))
k_fe_step(const FeArgs A)
{
    const int nE = A.nE, nC = A.nC;
    int b = blockIdx.x;
    if constexpr (LIST) b = A.blockList[blockIdx.x];
    const int e0 = A.blkEdgeStart[b], e1 = A.blkEdgeStart[b + 1];
    for (int e = e0 + threadIdx.x; e < e1; e += kThreads) {
        const int2 c = ld_stream(A.ce + e);
        int idx[S2T];
        double w[S2T];
#pragma unroll
        for (int i = 0; i < S2T; ++i) idx[i] = ld_stream(A.eoe + (size_t)i * nE + e);
#pragma unroll
        for (int i = 0; i < S2T; ++i) w[i] = ld_stream(A.woe + (size_t)i * nE + e);
        const double g = ld_stream(A.gdc + e);
        const double uo = A.u[e];
        const double s1 = __ldg(A.ssh + c.x), s2 = __ldg(A.ssh + c.y);
        const double h1 = __ldg(A.h + c.x), h2 = __ldg(A.h + c.y);
        double uu[S2T], ff[S2T];
#pragma unroll
        for (int i = 0; i < S2T; ++i) uu[i] = __ldg(A.u + idx[i]);
        if (!UNIF) {
#pragma unroll
            for (int i = 0; i < S2T; ++i) ff[i] = __ldg(A.fE + idx[i]);
        }
        double t = add_rn(0.0, -mul_rn(g, add_rn(s2, -s1)));
#pragma unroll
        for (int i = 0; i < S2T; ++i) t = add_rn(t, mul_rn(mul_rn(w[i], uu[i]), UNIF ? A.f0 : ff[i]));
        A.uNew[e] = add_rn(uo, mul_rn(A.dt, t));                     // UpdateStateVariable!, time_integration.jl:196-202
        A.hEnew[e] = mul_rn(0.5, add_rn(h1, h2));                    // interpolateCell2Edge!, Operators.jl:201-222
    }

    const int cc = b * kTC + threadIdx.x;
    if (cc < A.nCown) {
        const int n = ld_stream(A.nEoC + cc);
        int ee[ST];
#pragma unroll
        for (int i = 0; i < ST; ++i) ee[i] = i < n ? ld_stream(A.eoc + (size_t)i * nC + cc) : -1;
        double uu[ST], he[ST], dd[ST];
#pragma unroll
        for (int i = 0; i < ST; ++i) {
            const int e = ee[i] >= 0 ? (ee[i] >> 1) : 0;
            uu[i] = __ldg(A.u + e);
            he[i] = __ldg(A.hEold + e);
            dd[i] = __ldg(A.dv + e);
        }
        const double invA = ld_stream(A.invArea + cc);
        const double ho = A.h[cc];
        double t = 0.0;
#pragma unroll
        for (int i = 0; i < ST; ++i) {
            const double fd = mul_rn(mul_rn(uu[i], he[i]), dd[i]);   // (flux * dv), flux = u * hEdge_old
            if (ee[i] >= 0) t = add_rn(t, mul_rn((ee[i] & 1) ? fd : -fd, invA));
        }
        const double hn = add_rn(ho, mul_rn(A.dt, t));
        A.hNew[cc] = hn;
        A.sshNew[cc] = add_rn(hn, -ld_stream(A.H + cc));             // Update_ssh!, time_integration.jl:205-212
    }
}

// ---- construction of the fused-form arrays from the reference-form device arrays ---------------------
template <class R>
__global__ void __launch_bounds__(256)
k_build_fused_edges(int nE, int S2, const double *__restrict__ dc, const double *__restrict__ dv,
                    const double *__restrict__ fE, const int32_t *__restrict__ eoe, const double *__restrict__ woe,
                    const uint8_t *__restrict__ nEoE, R *__restrict__ gdc, R *__restrict__ dvR, R *__restrict__ wf,
                    int32_t *__restrict__ eoeF, int foldF)
{
    const int e = blockIdx.x * 256 + threadIdx.x;
    if (e >= nE) return;
    gdc[e] = (R)__dmul_rn(9.80616, 1.0 / dc[e]);
    dvR[e] = (R)dv[e];
    const int n = nEoE[e];
    for (int i = 0; i < S2; ++i) {
        const size_t k = (size_t)i * nE + e;
        const int x = i < n ? eoe[k] : -1;
        wf[k] = x >= 0 ? (foldF ? (R)__dmul_rn(woe[k], fE[x]) : (R)woe[k]) : R(0);
        if (eoeF) eoeF[k] = x >= 0 ? x : e;
    }
}

// the weights once more, slot-interleaved (TMA = 3): element ((g * nE + e) * V + j) = weight of slot g * V + j of edge e, V = 16 bytes / sizeof(R)
template <class R>
__global__ void __launch_bounds__(256)
k_build_wf_interleaved(int nE, int S2, const R *__restrict__ wf, R *__restrict__ wfI)
{
    constexpr int V = 16 / (int)sizeof(R);
    const int e = blockIdx.x * 256 + threadIdx.x;
    if (e >= nE) return;
    const int NG = (S2 + V - 1) / V;
    for (int g = 0; g < NG; ++g)
        for (int j = 0; j < V; ++j) {
            const int i = g * V + j;
            wfI[((size_t)g * nE + e) * V + j] = i < S2 ? wf[(size_t)i * nE + e] : R(0);
        }
}

// the weights once more, block-major (TMA = 2): block b's S2 rows back to back at wfBOff[b], each padded to a multiple of 16 bytes
template <class R>
__global__ void __launch_bounds__(256)
k_build_wf_block_major(int nE, int S2, const int32_t *__restrict__ blkEdgeStart, const long long *__restrict__ off,
                       const R *__restrict__ wf, R *__restrict__ wfB)
{
    const int b = blockIdx.x;
    const int e0 = blkEdgeStart[b], nb = blkEdgeStart[b + 1] - e0;
    constexpr int AL = 16 / (int)sizeof(R);
    const int nbp = (nb + AL - 1) / AL * AL;
    for (int k = threadIdx.x; k < S2 * nbp; k += 256) {
        const int i = k / nbp, j = k - i * nbp;
        wfB[off[b] + k] = j < nb ? wf[(size_t)i * nE + e0 + j] : R(0);
    }
}

template <class R>
__global__ void __launch_bounds__(256)
k_build_fused_cells(int nC, int S, const double *__restrict__ area, const double *__restrict__ H,
                    const int32_t *__restrict__ eoc, const int32_t *__restrict__ sgn, const uint8_t *__restrict__ nEoC,
                    R *__restrict__ invArea, R *__restrict__ HR, int32_t *__restrict__ eocF)
{
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c >= nC) return;
    invArea[c] = (R)(1.0 / area[c]);
    HR[c] = (R)H[c];
    if (!eocF) return;
    const int n = nEoC[c];
    for (int i = 0; i < S; ++i) {
        const size_t k = (size_t)i * nC + c;
        eocF[k] = i < n ? ((eoc[k] << 1) | (sgn[k] > 0 ? 1 : 0)) : -1;
    }
}

}  // namespace fused

// ---- halo messages: combined index space [cells | edges] (device numbering) ---------------------------------
template <class R>
__global__ void __launch_bounds__(256)
k_halo_pack(int n, int nC, const int32_t *__restrict__ idx, const R *__restrict__ h, const R *__restrict__ u, R *__restrict__ buf)
{
    MOKAB_TRACE_BEGIN();
    const int k = blockIdx.x * 256 + threadIdx.x;
    if (k < n) {
        const int i = idx[k];
        buf[k] = i < nC ? h[i] : u[i - nC];
    }
    MOKAB_TRACE_END(110u);
}
template <class R>
__global__ void __launch_bounds__(256)
k_halo_unpack(int n, int nC, const int32_t *__restrict__ idx, const R *__restrict__ buf, R *__restrict__ h, R *__restrict__ u)
{
    MOKAB_TRACE_BEGIN();
    const int k = blockIdx.x * 256 + threadIdx.x;
    if (k < n) {
        const int i = idx[k];
        if (i < nC) h[i] = buf[k];
        else u[i - nC] = buf[k];
    }
    MOKAB_TRACE_END(111u);
}

// Multi-level states: one message per peer carries K + 1 planes -- the K levels of (h on halo cells, u on halo edges) and the
// free surface of the halo cells (its edge part unused) -- so a stage needs ONE exchange whatever K.  Segment q of the message
// (the entities rank q gets / sends, segOff[q] ... segOff[q + 1] of the single-level list) holds its planes back to back.
template <class R>
__device__ __forceinline__ size_t halo_ml_slot(int k, int plane, int K, const int32_t *__restrict__ segOff, int nseg)
{
    int lo = 0, hi = nseg;                 // the segment of item k: the last q with segOff[q] <= k
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (segOff[mid] <= k) lo = mid; else hi = mid;
    }
    const int cnt = segOff[lo + 1] - segOff[lo];
    return (size_t)segOff[lo] * (size_t)(K + 1) + (size_t)plane * cnt + (size_t)(k - segOff[lo]);
}
template <class R>
__global__ void __launch_bounds__(256)
k_halo_pack_ml(int n, int nC, int nE, int K, const int32_t *__restrict__ idx, const int32_t *__restrict__ segOff, int nseg,
               const R *__restrict__ h, const R *__restrict__ u, const R *__restrict__ ssh, R *__restrict__ buf)
{
    const int nb = (n + 255) / 256;                                   // grid = nb blocks per plane x (K + 1) planes
    const int plane = blockIdx.x / nb, k = (blockIdx.x - plane * nb) * 256 + threadIdx.x;
    if (k >= n) return;
    const int i = idx[k];
    R v;
    if (plane < K) v = i < nC ? h[(size_t)plane * nC + i] : u[(size_t)plane * nE + (i - nC)];
    else v = i < nC ? ssh[i] : R(0);
    buf[halo_ml_slot<R>(k, plane, K, segOff, nseg)] = v;
}
template <class R>
__global__ void __launch_bounds__(256)
k_halo_unpack_ml(int n, int nC, int nE, int K, const int32_t *__restrict__ idx, const int32_t *__restrict__ segOff, int nseg,
                 const R *__restrict__ buf, R *__restrict__ h, R *__restrict__ u, R *__restrict__ ssh)
{
    const int nb = (n + 255) / 256;
    const int plane = blockIdx.x / nb, k = (blockIdx.x - plane * nb) * 256 + threadIdx.x;
    if (k >= n) return;
    const int i = idx[k];
    const R v = buf[halo_ml_slot<R>(k, plane, K, segOff, nseg)];
    if (plane < K) {
        if (i < nC) h[(size_t)plane * nC + i] = v;
        else u[(size_t)plane * nE + (i - nC)] = v;
    } else if (i < nC) {
        ssh[i] = v;
    }
}

// ---- deterministic reductions (replace sumArray, reference run_loop.jl:47-51) --------------------------
namespace reduce {
constexpr int kBlocks = 1024, kThreads = 256;

__device__ __forceinline__ double block_sum(double v)
{
    __shared__ double sm[kThreads / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    v = threadIdx.x < kThreads / 32 ? sm[threadIdx.x] : 0.0;
    if (threadIdx.x < 32) {
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    }
    return v;  // valid in thread 0
}

// which: 0 sum ssh^2 ; 1 sum area*h ; 2 potential energy sum area*g/2*ssh^2
template <class R>
__global__ void __launch_bounds__(kThreads)
k_cells(int which, int64_t nC, int K, int64_t stride, const R *__restrict__ h, const R *__restrict__ H, const double *__restrict__ area,
        double *__restrict__ partial)
{
    double s = 0.0;
    for (int64_t c = (int64_t)blockIdx.x * kThreads + threadIdx.x; c < nC; c += (int64_t)kBlocks * kThreads) {
        R col = h[c];
        for (int k = 1; k < K; ++k) col += h[(int64_t)k * stride + c];      // the whole column (K levels, level-major)
        const double ssh = fused::kPert<R> ? (double)col : (double)(col - H[c]);
        const double htot = fused::kPert<R> ? (double)H[c] + (double)col : (double)col;
        if (which == 0) s += ssh * ssh;
        else if (which == 1) s += area[c] * htot;
        else s += area[c] * (0.5 * 9.80616) * ssh * ssh;
    }
    s = block_sum(s);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

// kinetic energy sum_e (dc*dv/2) * hEdge * u^2
template <class R>
__global__ void __launch_bounds__(kThreads)
k_edges_ke(int64_t nE, int K, int64_t strideE, int64_t strideC, const int2 *__restrict__ ce, const double *__restrict__ dc,
           const double *__restrict__ dv, const R *__restrict__ u, const R *__restrict__ h, const R *__restrict__ H,
           double *__restrict__ partial)
{
    double s = 0.0;
    for (int64_t e = (int64_t)blockIdx.x * kThreads + threadIdx.x; e < nE; e += (int64_t)kBlocks * kThreads) {
        const int2 c = ce[e];
        for (int k = 0; k < K; ++k) {
            double he = 0.5 * ((double)h[(int64_t)k * strideC + c.x] + (double)h[(int64_t)k * strideC + c.y]);
            if (fused::kPert<R>) he += 0.5 * ((double)H[c.x] + (double)H[c.y]);
            const double ue = (double)u[(int64_t)k * strideE + e];
            s += 0.5 * dc[e] * dv[e] * he * ue * ue;
        }
    }
    s = block_sum(s);
    if (threadIdx.x == 0) partial[blockIdx.x] += s;
}

__global__ void __launch_bounds__(kThreads) k_final(const double *__restrict__ partial, double *__restrict__ out)
{
    double s = 0.0;
    for (int i = threadIdx.x; i < kBlocks; i += kThreads) s += partial[i];
    s = block_sum(s);
    if (threadIdx.x == 0) out[0] = s;
}
}  // namespace reduce

}  // namespace mokab
