// kernels_adjoint.cuh -- reverse mode (discrete adjoint) of the fused RungeKutta4 stage.
//
// Stands behind what the reference obtains from Enzyme: `autodiff(Reverse, ocn_run_loop, Duplicated(Prog,
// d_Prog), ...)` (test/enzyme/test_Enzyme_end2end.jl:30-110, ext/MPASEnzymeExt.jl), which on CUDA returns
// NaN in the reference (`@test_broken`, test_Enzyme_end2end.jl:182-186).  Here the adjoint is written by
// hand in GATHER form, so it has the forward kernel's structure: one thread per edge / cell, no atomics,
// deterministic.
//
// One RK4 step is x' = x + sum_s b_s k_s, k_s = F(y_s), y_1 = x, y_{s+1} = x + a_s k_s.  With lam' the
// adjoint of x':  kbar_4 = b_4 lam';  ybar_s = J_s^T kbar_s;  kbar_{s-1} = b_{s-1} lam' + a_{s-1} ybar_s;
// lam = lam' + sum_s ybar_s.  Per stage this kernel forms ybar_s = (dF/dy)^T kbar_s at y_s:
//   ubar[e] = hEdge[e] * G[e] + sum_j wfT[j,e] * kbar_u[eoeT[j,e]]
//             G[e] = dv[e] * (q[c2] - q[c1]),  q[c] = invArea[c] * kbar_h[c]           (transpose of
//             horizontal_advection.jl:64-65 with flux = u*hEdge; of horizontal_advection_and_coriolis.jl:70-72)
//   hbar[c] = sum_{e of c} ( -sign(c,e) * gdc[e] * kbar_u[e]  +  u[e]/2 * G[e] )      (transpose of
//             pressure_gradient.jl:63 with ssh = h - H; of Operators.jl:201-222)
// and fuses the RK bookkeeping (accumulate lam, emit kbar_{s-1}) exactly like the forward stage does.
// kbar_h is carried pre-multiplied by invArea (q) so the cell-side gather needs one array, not two.
// (eoeT, wfT) is the transpose of the Coriolis stencil, built once per mesh (moka_b200.cu: ensure_adjoint).
// On masked (solid-wall) edges cellsOnEdge[2] == cellsOnEdge[1]: no pressure term, hEdge = h[c1], one cell side.
#pragma once
#include "common.cuh"

namespace mokab {
namespace adjoint {

constexpr int kThreads = 256;

template <class R>
struct AdjArgs {
    int nE, nC, nCown;
    int S2T, S;
    const int2 *ce;
    const int32_t *eoeT;      // (S2T, nE) edges whose Coriolis sum reads this edge
    const int32_t *eoc;       // (S, nC)   (edge << 1) | (sign > 0)
    const uint8_t *nEoET, *nEoC;
    const int32_t *blkEdgeStart;
    const R *gdc, *wfT, *dv, *invArea;
    const R *uY, *hY;         // the state y_s the Jacobian is taken at
    const R *kuIn, *kqIn;     // kbar_u[e], invArea[c]*kbar_h[c] of this stage (FIRST: lam' itself, scaled on the fly)
    const R *lamU, *lamH;     // lam'
    R *accU, *accH;           // lam, accumulated over the four stages
    R *kuOut, *kqOut;         // kbar of the previous stage (not written by LAST)
    R aPrev, bPrev, bThis;    // a_{s-1}, b_{s-1};  b_4 (FIRST only)
};

// MODE: 0 = FIRST (RK stage 4: kbar = b_4 lam', acc = lam' + ybar), 1 = MIDDLE (stages 3, 2), 2 = LAST (stage 1)
template <class R, int MODE>
__global__ void __launch_bounds__(kThreads, 4)
k_rk_stage_adj(const AdjArgs<R> A)
{
    const int nE = A.nE, nC = A.nC;
    const int b = blockIdx.x;
    auto ku = [&](int e) -> R { return MODE == 0 ? A.bThis * __ldg(A.lamU + e) : __ldg(A.kuIn + e); };
    auto kq = [&](int c) -> R { return MODE == 0 ? A.bThis * __ldg(A.invArea + c) * __ldg(A.lamH + c) : __ldg(A.kqIn + c); };

    const int e0 = A.blkEdgeStart[b], e1 = A.blkEdgeStart[b + 1];
    for (int e = e0 + threadIdx.x; e < e1; e += kThreads) {
        const int2 c = ld_stream(A.ce + e);
        const int n = ld_stream(A.nEoET + e);
        const R lam = MODE == 2 ? R(0) : A.lamU[e];
        const R accIn = MODE == 0 ? lam : A.accU[e];
        const R q1 = kq(c.x), h1 = __ldg(A.hY + c.x);
        const bool masked = c.x == c.y;
        const R q2 = masked ? R(0) : kq(c.y), h2 = masked ? h1 : __ldg(A.hY + c.y);
        const R G = ld_stream(A.dv + e) * (q2 - q1);
        R yb = R(0.5) * (h1 + h2) * G;
        for (int j = 0; j < n; ++j) {
            const int x = ld_stream(A.eoeT + (size_t)j * nE + e);
            yb += ld_stream(A.wfT + (size_t)j * nE + e) * ku(x);
        }
        A.accU[e] = accIn + yb;
        if (MODE != 2) A.kuOut[e] = A.bPrev * lam + A.aPrev * yb;
    }

    const int cc = b * kThreads + threadIdx.x;
    if (cc < A.nCown) {
        const int n = ld_stream(A.nEoC + cc);
        const R lam = MODE == 2 ? R(0) : A.lamH[cc];
        const R accIn = MODE == 0 ? lam : A.accH[cc];
        const R qc = kq(cc);
        R yb = R(0);
        for (int i = 0; i < n; ++i) {
            const int ex = ld_stream(A.eoc + (size_t)i * nC + cc);
            const int e = ex >> 1;
            const R sgn = (ex & 1) ? R(1) : R(-1);
            const int2 cs = __ldg(A.ce + e);
            const bool masked = cs.x == cs.y;
            const int other = cs.x == cc ? cs.y : cs.x;
            const R kue = ku(e);
            const R G = __ldg(A.dv + e) * sgn * (masked ? qc : qc - kq(other));
            yb += (masked ? R(1) : R(0.5)) * __ldg(A.uY + e) * G;
            if (!masked) yb -= sgn * __ldg(A.gdc + e) * kue;
        }
        A.accH[cc] = accIn + yb;
        if (MODE != 2) A.kqOut[cc] = ld_stream(A.invArea + cc) * (A.bPrev * lam + A.aPrev * yb);
    }
}

// d_h += d_ssh (ssh = h - restingThicknessSum, Update_ssh! time_integration.jl:205-212), d_ssh <- 0
template <class R>
__global__ void __launch_bounds__(256) k_fold_dssh(int64_t n, R *__restrict__ dssh, R *__restrict__ dh)
{
    const int64_t c = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (c < n) {
        dh[c] += dssh[c];
        dssh[c] = R(0);
    }
}

// seed of J = sum ssh^2 (sumArray, run_loop.jl:47-51): d_ssh = 2 * ssh, ssh = h - H
template <class R>
__global__ void __launch_bounds__(256)
k_seed_ssh2(int64_t n, const R *__restrict__ h, const R *__restrict__ H, R *__restrict__ dssh)
{
    const int64_t c = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (c < n) dssh[c] = R(2) * (h[c] - H[c]);
}

}  // namespace adjoint
}  // namespace mokab
