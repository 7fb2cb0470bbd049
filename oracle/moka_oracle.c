/* ORACLE (test infrastructure, not product code): C/OpenMP restatement of MOKA's forward path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  It is the "reference-equivalent CPU restatement" of SURVEY.md section 8d:
 * one C function per reference KernelAbstractions kernel, same unfused passes, same
 * (slot, entity) column-major layouts, Float64, Int32 1-based connectivity (0 = absent), same
 * left-to-right operation order (build with -ffp-contract=off so no FMA is formed, matching
 * Julia's CPU code generation).  OpenMP `schedule(static)` over contiguous index blocks stands
 * in for the KernelAbstractions CPU backend's workgroup partition.
 *
 * PARITY PINNING: the reference cannot run here (no Julia).  Its tests pin only the operator
 * kernels (test/ocn/test_Operators.jl:52-53,72-73,90-91; reproduced in tests/).  Tendency,
 * ForwardEuler and run-loop field values are unpinned by the reference; RungeKutta4 is dead
 * code there (src/forward/time_integration.jl:61-148) and is project-defined here.
 *
 * Paths cited below are relative to /root/reference.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    int64_t nCells, nEdges, nVertices, maxEdges, maxEdges2, vertexDegree;
    /* Edges (src/infra/MPASMesh/HorzMesh.jl:64-95) */
    const int32_t *cellsOnEdge;      /* (2, nEdges) */
    const int32_t *edgesOnEdge;      /* (maxEdges2, nEdges) */
    const int32_t *nEdgesOnEdge;     /* (nEdges) */
    const double  *weightsOnEdge;    /* (maxEdges2, nEdges) */
    const double  *dcEdge, *dvEdge, *fEdge;
    /* PrimaryCells (HorzMesh.jl:102-132) */
    const int32_t *edgesOnCell;      /* (maxEdges, nCells) */
    const int32_t *edgeSignOnCell;   /* (maxEdges, nCells) */
    const int32_t *nEdgesOnCell;
    const double  *areaCell;
    /* DualCells (HorzMesh.jl:135-162) */
    const int32_t *edgesOnVertex;    /* (vertexDegree, nVertices) */
    const int32_t *edgeSignOnVertex; /* (maxEdges, nVertices) */
    const double  *areaTriangle;
    /* VerticalMesh (src/infra/MPASMesh/VertMesh.jl:3-17) */
    const int32_t *maxLevelEdgeTop;  /* (nEdges), all ones (VertMesh.jl:31-36) */
    const double  *restingThicknessSum; /* (nCells) */
} ora_mesh;

typedef struct {
    /* PrognosticVars, two time levels (src/ocn/PrognosticVars.jl:6-57) */
    double *ssh[2], *normalVelocity[2], *layerThickness[2];
    /* DiagnosticVars (src/ocn/DiagnosticVars.jl:6-73) */
    double *layerThicknessEdge, *thicknessFlux, *velocityDivCell, *relativeVorticity;
    /* TendencyVars (src/ocn/Tendencies/TendencyVars.jl:7-49) */
    double *tendNormalVelocity, *tendLayerThickness;
    /* RK4 work arrays (provisional and accumulator) */
    double *uProvis, *hProvis, *sshProvis, *uNew, *hNew;
} ora_state;

/* The reference's CPU path fans its workgroups over all Julia threads; under torchrun the environment pins OMP_NUM_THREADS=1,
 * so the benchmark's reference arm sets the team size explicitly (bench.py: run_reference). */
void ora_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int ora_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ZeroOutVector!, src/ocn/Operators.jl:225-231 */
void ora_zero_out_vector(double *x, int64_t n)
{
#pragma omp parallel for schedule(static)
    for (int64_t j = 0; j < n; ++j) x[j] = 0.0;
}

/* SSHGradOnEdge!, src/ocn/Tendencies/normalVelocity/pressure_gradient.jl:45-65 */
void ora_ssh_grad_on_edge(const ora_mesh *m, double *tend, const double *ssh)
{
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < m->nEdges; ++e) {
        int32_t c1 = m->cellsOnEdge[2 * e], c2 = m->cellsOnEdge[2 * e + 1];
        double inv = 1.0 / m->dcEdge[e];
        for (int32_t k = 0; k < m->maxLevelEdgeTop[e]; ++k)
            tend[e] -= 9.80616 * inv * (ssh[c2 - 1] - ssh[c1 - 1]);
    }
}

/* coriolis_force_tendency_kernel!, .../horizontal_advection_and_coriolis.jl:50-75 */
void ora_coriolis_force_tendency(const ora_mesh *m, double *tend, const double *u)
{
    const int64_t S = m->maxEdges2;
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < m->nEdges; ++e) {
        for (int32_t i = 0; i < m->nEdgesOnEdge[e]; ++i) {
            int32_t eoe = m->edgesOnEdge[S * e + i];
            if (eoe == 0) continue;
            for (int32_t k = 0; k < m->maxLevelEdgeTop[e]; ++k)
                tend[e] += m->weightsOnEdge[S * e + i] * u[eoe - 1] * m->fEdge[eoe - 1];
        }
    }
}

/* thicknessFluxDivOnCell!, src/ocn/Tendencies/layerThickness/horizontal_advection.jl:42-69 */
void ora_thickness_flux_div_on_cell(const ora_mesh *m, double *tend, const double *flux)
{
    const int64_t S = m->maxEdges;
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < m->nCells; ++c) {
        double invArea = 1.0 / m->areaCell[c];
        for (int32_t i = 0; i < m->nEdgesOnCell[c]; ++i) {
            int32_t e = m->edgesOnCell[S * c + i];
            for (int32_t k = 0; k < m->maxLevelEdgeTop[e - 1]; ++k)
                tend[c] += flux[e - 1] * m->dvEdge[e - 1] * (double)m->edgeSignOnCell[S * c + i] * invArea;
        }
    }
}

/* compute_thicknessFlux!, src/ocn/DiagnosticVars.jl:158-173 */
void ora_compute_thickness_flux(double *flux, const double *u, const double *hEdge, int64_t n)
{
#pragma omp parallel for schedule(static)
    for (int64_t j = 0; j < n; ++j) flux[j] = u[j] * hEdge[j];
}

/* DivergenceOnCell_P1 / _P2, src/ocn/Operators.jl:12-44 */
void ora_divergence_on_cell(const ora_mesh *m, double *div, const double *vec, double *temp)
{
    const int64_t S = m->maxEdges;
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < m->nEdges; ++e) temp[e] = vec[e] * m->dvEdge[e];
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < m->nCells; ++c) {
        double d = 0.0;
        for (int32_t i = 0; i < m->nEdgesOnCell[c]; ++i) {
            int32_t e = m->edgesOnCell[S * c + i];
            d -= temp[e - 1] * (double)m->edgeSignOnCell[S * c + i];
        }
        div[c] = d / m->areaCell[c];
    }
}

/* CurlOnVertex, src/ocn/Operators.jl:122-149 (accumulates; never zeroed, :135) */
void ora_curl_on_vertex(const ora_mesh *m, double *curl, const double *vec)
{
    const int64_t S = m->maxEdges, D = m->vertexDegree;
#pragma omp parallel for schedule(static)
    for (int64_t v = 0; v < m->nVertices; ++v) {
        double inv = 1.0 / m->areaTriangle[v];
        for (int64_t j = 0; j < D; ++j) {
            int32_t e = m->edgesOnVertex[D * v + j];
            curl[v] += m->dcEdge[e - 1] * inv * vec[e - 1] * (double)m->edgeSignOnVertex[S * v + j];
        }
    }
}

/* GradientOnEdge, src/ocn/Operators.jl:84-100 */
void ora_gradient_on_edge(const ora_mesh *m, double *grad, const double *s)
{
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < m->nEdges; ++e) {
        int32_t c1 = m->cellsOnEdge[2 * e], c2 = m->cellsOnEdge[2 * e + 1];
        grad[e] = (s[c2 - 1] - s[c1 - 1]) / m->dcEdge[e];
    }
}

/* interpolateCell2Edge, src/ocn/Operators.jl:201-222 */
void ora_interpolate_cell2edge(const ora_mesh *m, double *edgeValue, const double *cellValue)
{
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < m->nEdges; ++e) {
        int32_t c1 = m->cellsOnEdge[2 * e], c2 = m->cellsOnEdge[2 * e + 1];
        edgeValue[e] = 0.5 * (cellValue[c1 - 1] + cellValue[c2 - 1]);
    }
}

/* advance_2d_array / advance_3d_array, src/forward/time_integration.jl:42-59 */
static void ora_advance_array(double *prev, const double *next, int64_t n)
{
#pragma omp parallel for schedule(static)
    for (int64_t j = 0; j < n; ++j) prev[j] = next[j];
}

/* UpdateStateVariable!, src/forward/time_integration.jl:196-202 */
void ora_update_state_variable(double *var, const double *tend, double dt, int64_t n)
{
#pragma omp parallel for schedule(static)
    for (int64_t j = 0; j < n; ++j) var[j] = var[j] + dt * tend[j];
}

/* Update_ssh!, src/forward/time_integration.jl:205-212 */
void ora_update_ssh(double *ssh, const double *h, const double *H, int64_t n)
{
#pragma omp parallel for schedule(static)
    for (int64_t j = 0; j < n; ++j) ssh[j] = h[j] - H[j];
}

/* sumArray, src/forward/run_loop.jl:47-51 (serial, index order) */
double ora_sum_array(const double *a, int64_t n)
{
    double s = 0.0;
    for (int64_t j = 0; j < n; ++j) s = s + a[j] * a[j];
    return s;
}

/* advanceTimeLevels!, src/forward/time_integration.jl:10-40 */
void ora_advance_time_levels(const ora_mesh *m, ora_state *s)
{
    ora_advance_array(s->ssh[0], s->ssh[1], m->nCells);
    ora_advance_array(s->normalVelocity[0], s->normalVelocity[1], m->nEdges);
    ora_advance_array(s->layerThickness[0], s->layerThickness[1], m->nCells);
}

/* diagnostic_compute!, src/ocn/DiagnosticVars.jl:108-117, in the reference's order */
void ora_diagnostic_compute(const ora_mesh *m, ora_state *s, const double *u, const double *h)
{
    ora_compute_thickness_flux(s->thicknessFlux, u, s->layerThicknessEdge, m->nEdges);       /* stale hEdge (Q1) */
    ora_divergence_on_cell(m, s->velocityDivCell, u, s->layerThicknessEdge);                /* scratch alias :187-190 */
    if (m->nVertices) ora_curl_on_vertex(m, s->relativeVorticity, u);                       /* accumulates (Q2) */
    ora_interpolate_cell2edge(m, s->layerThicknessEdge, h);
}

/* computeNormalVelocityTendency!, src/ocn/Tendencies/normalVelocity/normalVelocity.jl:21-53 */
void ora_compute_normal_velocity_tendency(const ora_mesh *m, double *tend, const double *ssh, const double *u)
{
    ora_zero_out_vector(tend, m->nEdges);
    ora_ssh_grad_on_edge(m, tend, ssh);
    ora_coriolis_force_tendency(m, tend, u);
}

/* computeLayerThicknessTendency!, src/ocn/Tendencies/layerThickness/layerThickness.jl:14-28 */
void ora_compute_layer_thickness_tendency(const ora_mesh *m, double *tend, const double *flux)
{
    ora_zero_out_vector(tend, m->nCells);
    ora_thickness_flux_div_on_cell(m, tend, flux);
}

/* ocn_timestep(::ForwardEuler), src/forward/time_integration.jl:150-193 */
void ora_timestep_forward_euler(const ora_mesh *m, ora_state *s, double dt)
{
    ora_advance_time_levels(m, s);
    ora_diagnostic_compute(m, s, s->normalVelocity[1], s->layerThickness[1]);
    ora_compute_normal_velocity_tendency(m, s->tendNormalVelocity, s->ssh[1], s->normalVelocity[1]);
    ora_compute_layer_thickness_tendency(m, s->tendLayerThickness, s->thicknessFlux);
    ora_update_state_variable(s->normalVelocity[1], s->tendNormalVelocity, dt, m->nEdges);
    ora_update_state_variable(s->layerThickness[1], s->tendLayerThickness, dt, m->nCells);
    ora_update_ssh(s->ssh[1], s->layerThickness[1], m->restingThicknessSum, m->nCells);
}

static void ora_axpy_out(double *out, const double *x, double a, const double *t, int64_t n)
{
#pragma omp parallel for schedule(static)
    for (int64_t j = 0; j < n; ++j) out[j] = x[j] + a * t[j];
}

/* ocn_timestep(::RungeKutta4) as intended, src/forward/time_integration.jl:61-148.
 * Project-defined semantics: diagnostics consistent with the provisional
 * state (hEdge, then flux, then tendencies), the unfused per-stage structure of :112-137. */
void ora_timestep_rk4(const ora_mesh *m, ora_state *s, double dt)
{
    const double a[3] = { dt / 2.0, dt / 2.0, dt };                                  /* :77 */
    const double b[4] = { dt / 6.0, dt / 3.0, dt / 3.0, dt / 6.0 };                  /* :78 */
    const int64_t nE = m->nEdges, nC = m->nCells;
    ora_advance_time_levels(m, s);
    const double *uCur = s->normalVelocity[0], *hCur = s->layerThickness[0];
    ora_advance_array(s->uProvis, uCur, nE);  ora_advance_array(s->hProvis, hCur, nC);
    ora_advance_array(s->uNew, uCur, nE);     ora_advance_array(s->hNew, hCur, nC);   /* :108-110 */
    for (int st = 0; st < 4; ++st) {
        ora_update_ssh(s->sshProvis, s->hProvis, m->restingThicknessSum, nC);         /* :127 */
        ora_interpolate_cell2edge(m, s->layerThicknessEdge, s->hProvis);              /* :130 */
        ora_compute_thickness_flux(s->thicknessFlux, s->uProvis, s->layerThicknessEdge, nE);
        ora_compute_normal_velocity_tendency(m, s->tendNormalVelocity, s->sshProvis, s->uProvis);  /* :114 */
        ora_compute_layer_thickness_tendency(m, s->tendLayerThickness, s->thicknessFlux);          /* :115 */
        if (st < 3) {                                                                 /* :124-125 */
            ora_axpy_out(s->uProvis, uCur, a[st], s->tendNormalVelocity, nE);
            ora_axpy_out(s->hProvis, hCur, a[st], s->tendLayerThickness, nC);
        }
        ora_axpy_out(s->uNew, s->uNew, b[st], s->tendNormalVelocity, nE);             /* :134-135 */
        ora_axpy_out(s->hNew, s->hNew, b[st], s->tendLayerThickness, nC);
    }
    ora_advance_array(s->normalVelocity[1], s->uNew, nE);                             /* :140-144 */
    ora_advance_array(s->layerThickness[1], s->hNew, nC);
    ora_update_ssh(s->ssh[1], s->hNew, m->restingThicknessSum, nC);                   /* :136 */
}

/* ocn_run_loop, src/forward/run_loop.jl:8-45 with the clock reduced to a step count.
 * stepper: 0 = ForwardEuler (live reference path), 1 = RungeKutta4 (intended). */
void ora_run_loop(const ora_mesh *m, ora_state *s, double dt, int64_t nsteps, int stepper)
{
    for (int64_t i = 0; i < nsteps; ++i) {
        if (stepper == 0) ora_timestep_forward_euler(m, s, dt);
        else              ora_timestep_rk4(m, s, dt);
    }
}
