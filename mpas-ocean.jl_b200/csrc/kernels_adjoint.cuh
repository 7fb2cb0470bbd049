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

constexpr int kThreads = MOKAB_BLOCK_CELLS;

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
    const R *H;               // Float32 only: hY holds h - H (kernels_fused.cuh: kPert), the flux Jacobian needs the whole thickness
    const R *uY, *hY;         // the state y_s the Jacobian is taken at
    const R *kuIn, *kqIn;     // kbar_u[e], invArea[c]*kbar_h[c] of this stage (FIRST: lam' itself, scaled on the fly)
    const R *lamU, *lamH;     // lam'
    R *accU, *accH;           // lam, accumulated over the four stages
    R *kuOut, *kqOut;         // kbar of the previous stage (not written by LAST)
    R aPrev, bPrev, bThis;    // a_{s-1}, b_{s-1};  b_4 (FIRST only)
    const R *kuP;             // ML launches only: kbar_u summed over the levels of the column (k_sum_levels), for the pressure term
};

// MODE: 0 = FIRST (RK stage 4: kbar = b_4 lam', acc = lam' + ybar), 1 = MIDDLE (stages 3, 2), 2 = LAST (stage 1)
// S2TT / ST: compile-time row widths of the transposed Coriolis stencil / edgesOnCell (0 = runtime loops).  With
// compile-time widths all index and weight loads of a row are issued first, then all gathers, so ~20 (edge) /
// ~30 (cell) independent loads are in flight per thread -- the kernel is latency-bound otherwise (ncu: 0.24
// eligible warps per cycle with rolled loops).
#ifndef MOKAB_ADJ_MINBLOCKS
#define MOKAB_ADJ_MINBLOCKS 4
#endif
// ML = true: one LEVEL of a multi-level state per launch (the arrays of the arguments point at that level).  Coriolis and
// thickness flux act level by level, so everything is the single-level kernel -- except the pressure gradient, which is ONE
// term -g/dc (ssh2 - ssh1), ssh = sum_k h_k - H, shared by the levels of a column: its adjoint is formed from kuP, the level
// sum of kbar_u, and goes to every level of hbar alike.
template <class R, int MODE, int S2TT, int ST, bool ML = false>
__global__ void __launch_bounds__(kThreads, MOKAB_BLOCKS_SCALED(MOKAB_ADJ_MINBLOCKS))
k_rk_stage_adj(const AdjArgs<R> A)
{
    const int nE = A.nE, nC = A.nC;
    const int b = blockIdx.x;
    auto ku = [&](int e) -> R { return MODE == 0 ? A.bThis * __ldg(A.lamU + e) : __ldg(A.kuIn + e); };
    auto hTot = [&](int c) -> R { return sizeof(R) == 4 ? __ldg(A.hY + c) + __ldg(A.H + c) : __ldg(A.hY + c); };
    auto kq = [&](int c) -> R { return MODE == 0 ? A.bThis * __ldg(A.invArea + c) * __ldg(A.lamH + c) : __ldg(A.kqIn + c); };

    const int e0 = A.blkEdgeStart[b], e1 = A.blkEdgeStart[b + 1];
    for (int e = e0 + threadIdx.x; e < e1; e += kThreads) {
        const int2 c = ld_stream(A.ce + e);
        const int n = ld_stream(A.nEoET + e);
        const R lam = MODE == 2 ? R(0) : A.lamU[e];
        const R accIn = MODE == 0 ? lam : A.accU[e];
        const bool masked = c.x == c.y;
        R yb;
        if constexpr (S2TT != 0) {
            int idx[S2TT ? S2TT : 1];
            R w[S2TT ? S2TT : 1], kk[S2TT ? S2TT : 1];
#pragma unroll
            for (int j = 0; j < S2TT; ++j) idx[j] = j < n ? ld_stream(A.eoeT + (size_t)j * nE + e) : e;
#pragma unroll
            for (int j = 0; j < S2TT; ++j) w[j] = j < n ? ld_stream(A.wfT + (size_t)j * nE + e) : R(0);
#pragma unroll
            for (int j = 0; j < S2TT; ++j) kk[j] = ku(idx[j]);
            const R q1 = kq(c.x), h1 = hTot(c.x);
            const R q2 = masked ? R(0) : kq(c.y), h2 = masked ? h1 : hTot(c.y);
            yb = R(0.5) * (h1 + h2) * (ld_stream(A.dv + e) * (q2 - q1));
#pragma unroll
            for (int j = 0; j < S2TT; ++j) yb += w[j] * kk[j];
        } else {
            const R q1 = kq(c.x), h1 = hTot(c.x);
            const R q2 = masked ? R(0) : kq(c.y), h2 = masked ? h1 : hTot(c.y);
            yb = R(0.5) * (h1 + h2) * (ld_stream(A.dv + e) * (q2 - q1));
            for (int j = 0; j < n; ++j) {
                const int x = ld_stream(A.eoeT + (size_t)j * nE + e);
                yb += ld_stream(A.wfT + (size_t)j * nE + e) * ku(x);
            }
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
        if constexpr (ST != 0) {
            int ee[ST ? ST : 1];
#pragma unroll
            for (int i = 0; i < ST; ++i) ee[i] = i < n ? ld_stream(A.eoc + (size_t)i * nC + cc) : -1;
            // two half-rows at a time: 5 gathers per slot would not fit the register budget for a whole row
            constexpr int CH = (ST % 3 == 0) ? 3 : (ST % 2 == 0 ? 2 : 1);
#pragma unroll
            for (int i0 = 0; i0 < ST; i0 += CH) {
                int2 cs[CH];
                R kue[CH], dd[CH], uu[CH], gg[CH], qo[CH];
#pragma unroll
                for (int i = 0; i < CH; ++i) {
                    const int e = ee[i0 + i] >= 0 ? (ee[i0 + i] >> 1) : 0;
                    cs[i] = __ldg(A.ce + e);
                    kue[i] = ML ? __ldg(A.kuP + e) : ku(e);
                    dd[i] = __ldg(A.dv + e);
                    uu[i] = __ldg(A.uY + e);
                    gg[i] = __ldg(A.gdc + e);
                }
#pragma unroll
                for (int i = 0; i < CH; ++i) qo[i] = kq(cs[i].x == cc ? cs[i].y : cs[i].x);
#pragma unroll
                for (int i = 0; i < CH; ++i) {
                    if (ee[i0 + i] < 0) continue;
                    const R sgn = (ee[i0 + i] & 1) ? R(1) : R(-1);
                    const bool masked = cs[i].x == cs[i].y;
                    const R G = dd[i] * sgn * (masked ? qc : qc - qo[i]);
                    yb += (masked ? R(1) : R(0.5)) * uu[i] * G;
                    if (!masked) yb -= sgn * gg[i] * kue[i];
                }
            }
        } else {
            for (int i = 0; i < n; ++i) {
                const int ex = ld_stream(A.eoc + (size_t)i * nC + cc);
                const int e = ex >> 1;
                const R sgn = (ex & 1) ? R(1) : R(-1);
                const int2 cs = __ldg(A.ce + e);
                const bool masked = cs.x == cs.y;
                const int other = cs.x == cc ? cs.y : cs.x;
                const R kue = ML ? __ldg(A.kuP + e) : ku(e);
                const R G = __ldg(A.dv + e) * sgn * (masked ? qc : qc - kq(other));
                yb += (masked ? R(1) : R(0.5)) * __ldg(A.uY + e) * G;
                if (!masked) yb -= sgn * __ldg(A.gdc + e) * kue;
            }
        }
        A.accH[cc] = accIn + yb;
        if (MODE != 2) A.kqOut[cc] = ld_stream(A.invArea + cc) * (A.bPrev * lam + A.aPrev * yb);
    }
}

// ---- adjoint of the ForwardEuler step (the stepper the reference differentiates, test_Enzyme_end2end.jl:78-96) ------
// Forward (fused::k_fe_step; u, h, ssh, hE = the step's inputs, hE the LAGGED layerThicknessEdge):
//   u'[e] = u[e] + dt * (-(g/dc[e]) (ssh[c2] - ssh[c1]) + sum_i w[i,e] u[x_i] f[x_i])
//   h'[c] = h[c] + dt * invArea[c] * sum_i sign[i,c] dv[e_i] u[e_i] hE[e_i]
//   ssh'  = h' - H ;  hE'[e] = (h[c1] + h[c2]) / 2
// Reverse, gather form (lam* = adjoints of the step's outputs, q = invArea * (lamH + lamS) carried from the previous call):
//   G[e]    = dt * dv[e] * (q[c2] - q[c1])                                  (adjoint of the thickness flux)
//   outU[e] = lamU[e] + dt * sum_j wT[j,e] lamU[xT_j] + hE[e] * G[e]        (wT: transposed Coriolis stencil, f folded)
//   outE[e] = u[e] * G[e]
//   outS[c] = -dt * sum_i sign[i,c] (g/dc[e_i]) lamU[e_i]                   (ssh is a separate input: only step 1 reads the array)
//   outH[c] = lamH[c] + lamS[c] + sum_i lamE[e_i] / 2
//   qOut[c] = invArea[c] * (outH[c] + outS[c])
// The Jacobian depends on the trajectory only through (u, hE): those two edge arrays are what the tape holds per step.
struct FeAdjArgs {
    int nE, nC, nCown, S2T, S;
    const int2 *ce;
    const int32_t *eoeT, *eoc;
    const uint8_t *nEoET, *nEoC;
    const int32_t *blkEdgeStart;
    const double *gdc, *wT, *dv, *invArea;
    const double *uN, *hEN;                        // taped inputs of the step being reversed
    const double *lamU, *lamH, *lamS, *lamE, *qIn;
    double *outU, *outH, *outS, *outE, *qOut;
    double dt;
};

template <int S2TT, int ST>
__global__ void __launch_bounds__(kThreads, MOKAB_BLOCKS_SCALED(MOKAB_ADJ_MINBLOCKS))
k_fe_step_adj(const FeAdjArgs A)
{
    const int nE = A.nE, nC = A.nC;
    const int b = blockIdx.x;
    const int e0 = A.blkEdgeStart[b], e1 = A.blkEdgeStart[b + 1];
    for (int e = e0 + threadIdx.x; e < e1; e += kThreads) {
        const int2 c = ld_stream(A.ce + e);
        const int n = ld_stream(A.nEoET + e);
        const bool masked = c.x == c.y;
        const double lam = A.lamU[e], un = ld_stream(A.uN + e), he = ld_stream(A.hEN + e);
        double cor = 0.0;
        if constexpr (S2TT != 0) {
            int idx[S2TT ? S2TT : 1];
            double w[S2TT ? S2TT : 1], kk[S2TT ? S2TT : 1];
#pragma unroll
            for (int j = 0; j < S2TT; ++j) idx[j] = j < n ? ld_stream(A.eoeT + (size_t)j * nE + e) : e;
#pragma unroll
            for (int j = 0; j < S2TT; ++j) w[j] = j < n ? ld_stream(A.wT + (size_t)j * nE + e) : 0.0;
#pragma unroll
            for (int j = 0; j < S2TT; ++j) kk[j] = __ldg(A.lamU + idx[j]);
#pragma unroll
            for (int j = 0; j < S2TT; ++j) cor += w[j] * kk[j];
        } else {
            for (int j = 0; j < n; ++j)
                cor += ld_stream(A.wT + (size_t)j * nE + e) * __ldg(A.lamU + ld_stream(A.eoeT + (size_t)j * nE + e));
        }
        const double q1 = __ldg(A.qIn + c.x), q2 = masked ? 0.0 : __ldg(A.qIn + c.y);
        const double G = A.dt * (ld_stream(A.dv + e) * (q2 - q1));
        A.outU[e] = lam + A.dt * cor + he * G;
        A.outE[e] = un * G;
    }

    const int cc = b * kThreads + threadIdx.x;
    if (cc < A.nCown) {
        const int n = ld_stream(A.nEoC + cc);
        const double mu = A.lamH[cc] + A.lamS[cc];
        double sbar = 0.0, avg = 0.0;
        if constexpr (ST != 0) {
            int ee[ST ? ST : 1];
#pragma unroll
            for (int i = 0; i < ST; ++i) ee[i] = i < n ? ld_stream(A.eoc + (size_t)i * nC + cc) : -1;
            int2 cs[ST ? ST : 1];
            double lu[ST ? ST : 1], le[ST ? ST : 1], gg[ST ? ST : 1];
#pragma unroll
            for (int i = 0; i < ST; ++i) {
                const int e = ee[i] >= 0 ? (ee[i] >> 1) : 0;
                cs[i] = __ldg(A.ce + e);
                lu[i] = __ldg(A.lamU + e);
                le[i] = __ldg(A.lamE + e);
                gg[i] = __ldg(A.gdc + e);
            }
#pragma unroll
            for (int i = 0; i < ST; ++i) {
                if (ee[i] < 0) continue;
                const bool masked = cs[i].x == cs[i].y;
                const double sgn = (ee[i] & 1) ? 1.0 : -1.0;
                if (!masked) sbar -= sgn * gg[i] * lu[i];
                avg += (masked ? 1.0 : 0.5) * le[i];
            }
        } else {
            for (int i = 0; i < n; ++i) {
                const int ex = ld_stream(A.eoc + (size_t)i * nC + cc);
                const int e = ex >> 1;
                const int2 cs = __ldg(A.ce + e);
                const bool masked = cs.x == cs.y;
                const double sgn = (ex & 1) ? 1.0 : -1.0;
                if (!masked) sbar -= sgn * __ldg(A.gdc + e) * __ldg(A.lamU + e);
                avg += (masked ? 1.0 : 0.5) * __ldg(A.lamE + e);
            }
        }
        const double oS = A.dt * sbar, oH = mu + avg;
        A.outS[cc] = oS;
        A.outH[cc] = oH;
        A.qOut[cc] = ld_stream(A.invArea + cc) * (oH + oS);
    }
}

// start of the ForwardEuler reverse sweep: lamS <- the seed on ssh, lamE <- 0 (d_Diag.layerThicknessEdge), q <- invArea (lamH + lamS)
__global__ void __launch_bounds__(256)
k_fe_adj_begin(int64_t nC, int64_t nE, const double *__restrict__ invArea, const double *__restrict__ dssh,
               const double *__restrict__ lamH, double *__restrict__ lamS, double *__restrict__ lamE, double *__restrict__ q)
{
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i < nC) {
        lamS[i] = dssh[i];
        q[i] = invArea[i] * (lamH[i] + dssh[i]);
    }
    if (i < nE) lamE[i] = 0.0;
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

// multi-level states: out[e] = factor * sum_k in[k * n + e] (level order), and the seed on ssh folded into every level of d_h
template <class R>
__global__ void __launch_bounds__(256) k_sum_levels(int64_t n, int K, R factor, const R *__restrict__ in, R *__restrict__ out)
{
    const int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (e >= n) return;
    R s = in[e];
    for (int k = 1; k < K; ++k) s += in[(size_t)k * n + e];
    out[e] = factor * s;
}
template <class R>
__global__ void __launch_bounds__(256) k_fold_dssh_levels(int64_t n, int K, R *__restrict__ dssh, R *__restrict__ dh)
{
    const int64_t c = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (c < n) {
        for (int k = 0; k < K; ++k) dh[(size_t)k * n + c] += dssh[c];
        dssh[c] = R(0);
    }
}

// seed of J = sum ssh^2 (sumArray, run_loop.jl:47-51): d_ssh = 2 * ssh, ssh = h - H
template <class R>
__global__ void __launch_bounds__(256)
k_seed_ssh2(int64_t n, const R *__restrict__ h, const R *__restrict__ H, R *__restrict__ dssh)
{
    const int64_t c = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (c < n) dssh[c] = sizeof(R) == 4 ? R(2) * h[c] : R(2) * (h[c] - H[c]);   // Float32 `h` arrays hold h - H
}

// the same seed from the ssh ARRAY (ForwardEuler: ssh is a prognostic array of its own, equal to h - H only after a step)
template <class R>
__global__ void __launch_bounds__(256) k_seed_ssh2_array(int64_t n, const R *__restrict__ ssh, R *__restrict__ dssh)
{
    const int64_t c = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (c < n) dssh[c] = R(2) * ssh[c];
}

// ---- operator-level transposes (test/enzyme/test_Enzyme_Operators.jl differentiates exactly these two) -------
// GradientOnEdge (Operators.jl:84-100): grad[e] = (s[c2] - s[c1]) / dc[e]
//   => sbar[c] = sum_{e of c} sign(c,e) * gbar[e] / dc[e]      (sign = edgeSignOnCell: -1 on the c1 side)
__global__ void __launch_bounds__(256)
k_gradient_on_edge_vjp(int nC, const int32_t *__restrict__ eoc, const int32_t *__restrict__ sgn, const uint8_t *__restrict__ nEoC,
                       const int2 *__restrict__ ce, const double *__restrict__ dc, const double *__restrict__ gbar,
                       double *__restrict__ sbar)
{
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c >= nC) return;
    double acc = 0.0;
    const int n = nEoC[c];
    for (int i = 0; i < n; ++i) {
        const int e = eoc[(size_t)i * nC + c];
        const int2 cs = ce[e];
        if (cs.x == cs.y) continue;                       // masked edge: zero gradient
        acc += (double)sgn[(size_t)i * nC + c] * gbar[e] / dc[e];
    }
    sbar[c] = acc;
}

// DivergenceOnCell (Operators.jl:12-44): div[c] = -(1/area[c]) * sum_i sign[i,c] * dv[e_i] * F[e_i]
//   => Fbar[e] = dv[e] * (dbar[c1]/area[c1] - dbar[c2]/area[c2])
__global__ void __launch_bounds__(256)
k_divergence_on_cell_vjp(int nE, const int2 *__restrict__ ce, const double *__restrict__ dv, const double *__restrict__ area,
                         const double *__restrict__ dbar, double *__restrict__ fbar)
{
    const int e = blockIdx.x * 256 + threadIdx.x;
    if (e >= nE) return;
    const int2 c = ce[e];
    const double t1 = dbar[c.x] / area[c.x];
    const double t2 = c.x == c.y ? 0.0 : dbar[c.y] / area[c.y];
    fbar[e] = dv[e] * (t1 - t2);
}

}  // namespace adjoint
}  // namespace mokab
