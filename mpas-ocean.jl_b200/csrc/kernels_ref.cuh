// kernels_ref.cuh -- operator-level kernels in the REFERENCE's operation order (Float64).
//
// One kernel per reference entry point; inside, every output element is produced by exactly the
// sequence of IEEE operations the reference's KernelAbstractions kernels perform (Julia evaluates
// products left to right and forms no FMA on the CPU path).  The explicit
// __dmul_rn/__dadd_rn intrinsics stop nvcc from contracting to FMA, so these kernels agree with the
// CPU oracle bit for bit.  They back the src/ocn entry points, the ForwardEuler stepper and the
// unfused RungeKutta4 used as an on-device cross-check of the fused path.
#pragma once
#include "common.cuh"

namespace mokab {
namespace ref {

constexpr int kThreads = 256;
static inline int blocks_for(int64_t n) { return (int)((n + kThreads - 1) / kThreads); }

// computeNormalVelocityTendency!: ZeroOutVector! + SSHGradOnEdge! + coriolis_force_tendency_kernel!
// (reference Operators.jl:225-231, pressure_gradient.jl:45-65, horizontal_advection_and_coriolis.jl:50-75)
__global__ void __launch_bounds__(kThreads)
k_tend_normal_velocity(int nE, int S2, const int2 *__restrict__ ce, const double *__restrict__ dc,
                       const int32_t *__restrict__ eoe, const double *__restrict__ woe, const uint8_t *__restrict__ nEoE,
                       const double *__restrict__ fE, const double *__restrict__ ssh, const double *__restrict__ u,
                       double *__restrict__ tend)
{
    const int e = blockIdx.x * kThreads + threadIdx.x;
    if (e >= nE) return;
    const int2 c = ce[e];
    const double inv = 1.0 / dc[e];
    double t = 0.0;
    t = __dadd_rn(t, -__dmul_rn(__dmul_rn(9.80616, inv), __dadd_rn(ssh[c.y], -ssh[c.x])));
    const int n = nEoE[e];
    for (int i = 0; i < n; ++i) {
        const int x = eoe[(size_t)i * nE + e];
        if (x < 0) continue;
        t = __dadd_rn(t, __dmul_rn(__dmul_rn(woe[(size_t)i * nE + e], u[x]), fE[x]));
    }
    tend[e] = t;
}

// computeLayerThicknessTendency!: ZeroOutVector! + thicknessFluxDivOnCell!
// (reference layerThickness.jl:14-28, horizontal_advection.jl:42-69)
__global__ void __launch_bounds__(kThreads)
k_tend_layer_thickness(int nC, const int32_t *__restrict__ eoc, const int32_t *__restrict__ sgn,
                       const uint8_t *__restrict__ nEoC, const double *__restrict__ area, const double *__restrict__ dv,
                       const double *__restrict__ flux, double *__restrict__ tend)
{
    const int c = blockIdx.x * kThreads + threadIdx.x;
    if (c >= nC) return;
    const double invArea = 1.0 / area[c];
    double t = 0.0;
    const int n = nEoC[c];
    for (int i = 0; i < n; ++i) {
        const int e = eoc[(size_t)i * nC + c];
        t = __dadd_rn(t, __dmul_rn(__dmul_rn(__dmul_rn(flux[e], dv[e]), (double)sgn[(size_t)i * nC + c]), invArea));
    }
    tend[c] = t;
}

// diagnostic_compute! edge part: compute_thicknessFlux! with the stale hEdge, then interpolateCell2Edge
// (reference DiagnosticVars.jl:141-173 and :126-139, Operators.jl:201-222).  The DivergenceOnCell_P1
// scratch write into layerThicknessEdge (DiagnosticVars.jl:187-190) is dead: it is overwritten here.
__global__ void __launch_bounds__(kThreads)
k_diag_edges(int nE, const int2 *__restrict__ ce, const double *__restrict__ u, const double *__restrict__ h,
             double *__restrict__ hEdge, double *__restrict__ flux)
{
    const int e = blockIdx.x * kThreads + threadIdx.x;
    if (e >= nE) return;
    flux[e] = __dmul_rn(u[e], hEdge[e]);
    const int2 c = ce[e];
    hEdge[e] = __dmul_rn(0.5, __dadd_rn(h[c.x], h[c.y]));
}

// DivergenceOnCell_P1 + _P2 (reference Operators.jl:12-44)
__global__ void __launch_bounds__(kThreads)
k_divergence_on_cell(int nC, const int32_t *__restrict__ eoc, const int32_t *__restrict__ sgn,
                     const uint8_t *__restrict__ nEoC, const double *__restrict__ area, const double *__restrict__ dv,
                     const double *__restrict__ vec, double *__restrict__ div)
{
    const int c = blockIdx.x * kThreads + threadIdx.x;
    if (c >= nC) return;
    double d = 0.0;
    const int n = nEoC[c];
    for (int i = 0; i < n; ++i) {
        const int e = eoc[(size_t)i * nC + c];
        d = __dadd_rn(d, -__dmul_rn(__dmul_rn(vec[e], dv[e]), (double)sgn[(size_t)i * nC + c]));
    }
    div[c] = d / area[c];
}

// CurlOnVertex (reference Operators.jl:122-149): accumulates into curl, which the reference never zeroes.
__global__ void __launch_bounds__(kThreads)
k_curl_on_vertex(int nV, int D, const int32_t *__restrict__ eov, const int32_t *__restrict__ sgn,
                 const double *__restrict__ areaTri, const double *__restrict__ dc, const double *__restrict__ vec,
                 double *__restrict__ curl)
{
    const int v = blockIdx.x * kThreads + threadIdx.x;
    if (v >= nV) return;
    const double inv = 1.0 / areaTri[v];
    double acc = curl[v];
    for (int j = 0; j < D; ++j) {
        const int e = eov[(size_t)j * nV + v];
        acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(__dmul_rn(dc[e], inv), vec[e]), (double)sgn[(size_t)j * nV + v]));
    }
    curl[v] = acc;
}

// GradientOnEdge (reference Operators.jl:84-100)
__global__ void __launch_bounds__(kThreads)
k_gradient_on_edge(int nE, const int2 *__restrict__ ce, const double *__restrict__ dc, const double *__restrict__ s,
                   double *__restrict__ grad)
{
    const int e = blockIdx.x * kThreads + threadIdx.x;
    if (e >= nE) return;
    const int2 c = ce[e];
    grad[e] = __dadd_rn(s[c.y], -s[c.x]) / dc[e];
}

// interpolateCell2Edge (reference Operators.jl:201-222)
__global__ void __launch_bounds__(kThreads)
k_interpolate_cell2edge(int nE, const int2 *__restrict__ ce, const double *__restrict__ cellv, double *__restrict__ edgev)
{
    const int e = blockIdx.x * kThreads + threadIdx.x;
    if (e >= nE) return;
    const int2 c = ce[e];
    edgev[e] = __dmul_rn(0.5, __dadd_rn(cellv[c.x], cellv[c.y]));
}

// compute_thicknessFlux! (reference DiagnosticVars.jl:158-173)
__global__ void __launch_bounds__(kThreads)
k_mul(int64_t n, const double *__restrict__ a, const double *__restrict__ b, double *__restrict__ out)
{
    const int64_t j = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (j < n) out[j] = __dmul_rn(a[j], b[j]);
}

// UpdateStateVariable! (reference time_integration.jl:196-202) and the RK4 broadcasts
// Provis = Curr + a*tend / New = New + b*tend (:124-125, :134-135): out = x + a*t
__global__ void __launch_bounds__(kThreads)
k_axpy(int64_t n, const double *x, double a, const double *__restrict__ t, double *out)
{
    const int64_t j = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (j < n) out[j] = __dadd_rn(x[j], __dmul_rn(a, t[j]));
}

}  // namespace ref

// Update_ssh! (reference time_integration.jl:205-212); also used for F32 states
template <class R>
__global__ void __launch_bounds__(256)
k_update_ssh(int64_t n, const R *__restrict__ h, const R *__restrict__ H, R *__restrict__ ssh)
{
    const int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (j < n) ssh[j] = sizeof(R) == 4 ? h[j] : h[j] - H[j];   // Float32 `h` arrays hold h - H already (kernels_fused.cuh: kPert)
}

// caller order -> device order (dst[new] = src[perm[new]]) and back
template <class R>
__global__ void __launch_bounds__(256)
k_permute_in(int64_t n, const int32_t *__restrict__ perm, const R *__restrict__ src, R *__restrict__ dst)
{
    const int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (j < n) dst[j] = src[perm[j]];
}
template <class R>
__global__ void __launch_bounds__(256)
k_permute_out(int64_t n, const int32_t *__restrict__ perm, const R *__restrict__ src, R *__restrict__ dst)
{
    const int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (j < n) dst[perm[j]] = src[j];
}

// Multi-level fields: the caller's arrays are the reference's (nVertLevels, n) column-major arrays (level fastest,
// PrognosticVars.jl:10-16), the device keeps every level as one contiguous array over the (renumbered) entities
template <class R>
__global__ void __launch_bounds__(256)
k_permute_in_lv(int64_t n, int K, const int32_t *__restrict__ perm, const R *__restrict__ src, R *__restrict__ dst)
{
    const int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (j >= n) return;
    const R *col = src + (int64_t)K * perm[j];
    for (int k = 0; k < K; ++k) dst[(int64_t)k * n + j] = col[k];
}
template <class R>
__global__ void __launch_bounds__(256)
k_permute_out_lv(int64_t n, int K, const int32_t *__restrict__ perm, const R *__restrict__ src, R *__restrict__ dst)
{
    const int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (j >= n) return;
    R *col = dst + (int64_t)K * perm[j];
    for (int k = 0; k < K; ++k) col[k] = src[(int64_t)k * n + j];
}
// Update_ssh! for a column of K levels (project-defined for K > 1): ssh = (h[0] + h[1] + ...) - restingThicknessSum
__global__ void __launch_bounds__(256)
k_update_ssh_lv(int64_t n, int K, const double *__restrict__ h, const double *__restrict__ H, double *__restrict__ ssh)
{
    const int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (j >= n) return;
    double col = h[j];
    for (int k = 1; k < K; ++k) col = __dadd_rn(col, h[(int64_t)k * n + j]);
    ssh[j] = __dadd_rn(col, -H[j]);
}

// Float32 states: layerThickness crosses the boundary as the whole thickness, the device arrays hold h - H (kernels_fused.cuh: kPert)
__global__ void __launch_bounds__(256)
k_permute_in_pert(int64_t n, const int32_t *__restrict__ perm, const float *__restrict__ src, const double *__restrict__ H, float *__restrict__ dst)
{
    const int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (j < n) dst[j] = (float)((double)src[perm[j]] - H[j]);
}
__global__ void __launch_bounds__(256)
k_permute_out_pert(int64_t n, const int32_t *__restrict__ perm, const float *__restrict__ src, const double *__restrict__ H, float *__restrict__ dst)
{
    const int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (j < n) dst[perm[j]] = (float)(H[j] + (double)src[j]);
}

template <class T, class U>
__global__ void __launch_bounds__(256) k_convert(int64_t n, const T *__restrict__ src, U *__restrict__ dst)
{
    const int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (j < n) dst[j] = (U)src[j];
}

}  // namespace mokab
