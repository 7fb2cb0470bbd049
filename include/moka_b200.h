/* moka_b200.h -- C ABI of libmoka_b200.so: the B200 (sm_100a) implementation of MPAS-Ocean.jl's
 * ("MOKA") forward-model hot path: tendency evaluation + ForwardEuler / RungeKutta4 stepping on the
 * MPAS Voronoi C-grid.
 *
 * The reference has no FFI: its seam is Julia multiple dispatch on a `backend` value threaded through
 * the constructors and entry points listed below.  Each function here names the
 * reference interface it stands behind (paths relative to the reference repository).  A Julia `B200`
 * architecture type forwards those entry points here with `ccall`; in this repository
 * the same ABI is driven from Python ctypes (mpas-ocean.jl_b200/moka_b200/).
 *
 * Conventions
 *  - plain C: opaque handles, plain pointers and sizes, no C++/torch types.
 *  - every function returns 0 on success, non-zero on error; mokab_last_error() gives the message
 *    (the reference raises `error(msg)`: src/Architectures.jl:23,31,39).  No exceptions cross the ABI.
 *  - host arrays are in the REFERENCE layout and numbering: column-major (slot, entity), Int32 1-based
 *    connectivity with 0 = absent, exactly what ReadHorzMesh returns (src/infra/MPASMesh/HorzMesh.jl:
 *    166-290).  The library renumbers cells/edges/vertices for locality internally; mokab_state_set/get
 *    always speak the caller's numbering.
 *  - the caller owns host memory (copied during the call, never retained); the library owns device
 *    memory behind handles.  A context is bound to one device; calls on one context must be serialised
 *    by the caller (the reference drives from a single task: src/forward/run_loop.jl:8-22).
 *  - there is no CPU fallback: every entry point needs a CUDA device and fails loudly without one.
 */
#ifndef MOKA_B200_H
#define MOKA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MOKAB_VERSION 100

typedef struct mokab_ctx   mokab_ctx;
typedef struct mokab_mesh  mokab_mesh;
typedef struct mokab_state mokab_state;

/* element type of a state (reference: Float64 hard-coded, src/ocn/PrognosticVars.jl:91-93) */
enum { MOKAB_F64 = 0, MOKAB_F32 = 1 };

/* fields addressable with mokab_state_set / mokab_state_get */
enum {
    MOKAB_SSH = 0,                 /* Prog.ssh[end]            (nCells)    PrognosticVars.jl:6-18 */
    MOKAB_NORMAL_VELOCITY = 1,     /* Prog.normalVelocity[end] (nEdges)                            */
    MOKAB_LAYER_THICKNESS = 2,     /* Prog.layerThickness[end] (nCells)                            */
    MOKAB_SSH_PREV = 3,            /* Prog.ssh[1]              time level "previous"               */
    MOKAB_NORMAL_VELOCITY_PREV = 4,
    MOKAB_LAYER_THICKNESS_PREV = 5,
    MOKAB_LAYER_THICKNESS_EDGE = 6, /* Diag.layerThicknessEdge (nEdges)    DiagnosticVars.jl:6-73  */
    MOKAB_THICKNESS_FLUX = 7,       /* Diag.thicknessFlux      (nEdges)                            */
    MOKAB_VELOCITY_DIV_CELL = 8,    /* Diag.velocityDivCell    (nCells)                            */
    MOKAB_RELATIVE_VORTICITY = 9,   /* Diag.relativeVorticity  (nVertices)                         */
    MOKAB_TEND_NORMAL_VELOCITY = 10,/* Tend.tendNormalVelocity (nEdges)    TendencyVars.jl:7-49    */
    MOKAB_TEND_LAYER_THICKNESS = 11,/* Tend.tendLayerThickness (nCells)                            */
    /* shadow state d_Prog of the reverse mode (ocn_init_shadows, src/forward/init.jl:32-40; the
     * `Duplicated(Prog, d_Prog)` argument at test/enzyme/test_Enzyme_end2end.jl:78-96)             */
    MOKAB_D_SSH = 12,               /* d_Prog.ssh[end]            (nCells)                           */
    MOKAB_D_NORMAL_VELOCITY = 13,   /* d_Prog.normalVelocity[end] (nEdges)                           */
    MOKAB_D_LAYER_THICKNESS = 14    /* d_Prog.layerThickness[end] (nCells)                           */
};

/* reductions (replaces the serial sumArray kernel, src/forward/run_loop.jl:47-51) */
enum {
    MOKAB_SUM_SSH2 = 0,   /* sum_c ssh[c]^2                          (what sumArray computes)        */
    MOKAB_SUM_MASS = 1,   /* sum_c areaCell[c] * layerThickness[c]                                   */
    MOKAB_SUM_ENERGY = 2  /* sum_c area*g*ssh^2/2 + sum_e (dc*dv/2)*hEdge*u^2 (potential + kinetic; |u|^2 ~ 2 u_n^2) */
};

/* entity kinds for mokab_mesh_get_perm */
enum { MOKAB_CELLS = 0, MOKAB_EDGES = 1, MOKAB_VERTICES = 2 };

/* RungeKutta4 implementations */
enum {
    MOKAB_RK4_FUSED = 0,    /* one fused tendency+update kernel per stage, CUDA-graph resident loop      */
    MOKAB_RK4_UNFUSED = 1   /* the reference's per-stage kernel sequence, reference operation order      */
};

/* mesh_create flags */
enum {
    MOKAB_MESH_RENUMBER = 1u, /* locality renumbering (space-filling curve); 0 keeps the caller's order */
    MOKAB_MESH_EXPLICIT_EOE = 2u,/* always read edgesOnEdge from memory.  Default: the fused kernel rebuilds it from
                                    edgesOnCell wherever the mesh follows the MPAS ordering (verified per edge at
                                    mesh_create, per-block fallback); bit-identical results, 16 % fewer DRAM bytes */
    MOKAB_MESH_KEEP_WIDTHS = 4u, /* keep device rows as wide as the caller's maxEdges / maxEdges2.  Default: as wide as
                                    the longest live row (nEdgesOnCell / nEdgesOnEdge), so padded files of hexagons
                                    still take the compile-time-width kernels */
    MOKAB_MESH_EDGES_BY_CELL = 8u /* number the edges of a block cell by cell (a cell's edges adjacent).  Default: slot-major
                                    inside every block (every cell's first owned edge, then every cell's second, ...), so
                                    that a warp works on one kind of edge of 32 consecutive cells and its gathers coalesce */
};

/* Host view of the reference mesh structs.  Pointers marked (opt) may be NULL.
 *   Edges          src/infra/MPASMesh/HorzMesh.jl:64-95
 *   PrimaryCells   src/infra/MPASMesh/HorzMesh.jl:102-132
 *   DualCells      src/infra/MPASMesh/HorzMesh.jl:135-162
 *   VerticalMesh   src/infra/MPASMesh/VertMesh.jl:3-17 (single stacked layer: nVertLevels == 1)      */
typedef struct mokab_mesh_desc {
    int64_t nCells, nEdges, nVertices;          /* nVertices may be 0: no curl diagnostic           */
    int64_t maxEdges, maxEdges2, vertexDegree;
    /* Edges */
    const int32_t *cellsOnEdge;      /* (2, nEdges)                                                  */
    const int32_t *verticesOnEdge;   /* (2, nEdges)            (opt; needed to derive edgeSignOnVertex) */
    const int32_t *edgesOnEdge;      /* (maxEdges2, nEdges)                                          */
    const int32_t *nEdgesOnEdge;     /* (nEdges)                                                     */
    const double  *weightsOnEdge;    /* (maxEdges2, nEdges)                                          */
    const double  *dcEdge, *dvEdge;  /* (nEdges)                                                     */
    const double  *fEdge;            /* (nEdges)               (opt: zeros, HorzMesh.jl:257-262)     */
    const double  *xEdge, *yEdge, *zEdge; /* (opt; unused by the kernels)                            */
    /* PrimaryCells */
    const int32_t *edgesOnCell;      /* (maxEdges, nCells)                                           */
    const int32_t *nEdgesOnCell;     /* (nCells)                                                     */
    const int32_t *edgeSignOnCell;   /* (maxEdges, nCells)     (opt: derived as HorzMesh.jl:292-311) */
    const double  *areaCell;         /* (nCells)                                                     */
    const double  *xCell, *yCell, *zCell; /* (opt; drive the locality renumbering)                   */
    /* DualCells */
    const int32_t *edgesOnVertex;    /* (vertexDegree, nVertices) (opt if nVertices == 0)            */
    const int32_t *edgeSignOnVertex; /* (maxEdges, nVertices)  (opt: derived as HorzMesh.jl:313-332) */
    const double  *areaTriangle;     /* (nVertices)                                                  */
    /* VerticalMesh */
    const double  *restingThicknessSum; /* (nCells)            VertMesh.jl:73                        */
    /* masks (legacy glossary src/infra/Mesh.jl:110-114; project-defined for non-periodic meshes)    */
    const int32_t *boundaryEdge;     /* (nEdges) (opt) 1 = solid-wall edge: u = tendU = 0            */
    /* domain decomposition (no reference counterpart): the first nCellsOwned cells
     * and nEdgesOwned edges are computed by this rank, the rest are halo copies filled by
     * mokab_halo_unpack.  0 = everything is owned.  Connectivity rows of halo entities are not read.  */
    int64_t nCellsOwned, nEdgesOwned;
} mokab_mesh_desc;

/* ---- context: Architectures.jl backend object ------------------------------------------------ */
/* `B200()` architecture instance on CUDA device `device` (src/Architectures.jl:12; the reference
 * picks its backend at src/driver/mpas_ocean.jl:28). */
int  mokab_init(int device, mokab_ctx **out);
int  mokab_finalize(mokab_ctx *ctx);
/* KA.synchronize(backend) (e.g. src/ocn/Tendencies/normalVelocity/pressure_gradient.jl:39) */
int  mokab_synchronize(mokab_ctx *ctx);
/* Use a caller-owned cudaStream_t for all work of this context (NULL restores the library's own). */
int  mokab_set_stream(mokab_ctx *ctx, void *cuda_stream);
/* CUDA-event stopwatch on the context's stream (milliseconds between start and stop). */
int  mokab_timer_start(mokab_ctx *ctx);
int  mokab_timer_stop(mokab_ctx *ctx, double *elapsed_ms);
/* number of kernels this context has launched (graph replays count their kernel nodes) */
int  mokab_launch_count(mokab_ctx *ctx, int64_t *out);
/* page-locked host memory for state_set/get staging (Adapt.adapt source/target buffers) */
int  mokab_host_alloc(void **out, int64_t bytes);
int  mokab_host_free(void *p);
const char *mokab_last_error(void);
int  mokab_version(void);

/* ---- mesh: ReadHorzMesh + VerticalMesh + Adapt.adapt_structure(backend, mesh) ---------------- */
/* Upload a mesh (HorzMesh.jl:334-355 does the sign fields + H2D copy; VertMesh.jl:46-82). */
int  mokab_mesh_create(mokab_ctx *ctx, const mokab_mesh_desc *desc, uint32_t flags, mokab_mesh **out);
int  mokab_mesh_destroy(mokab_mesh *mesh);
/* perm[new] = old (0-based) for MOKAB_CELLS / MOKAB_EDGES / MOKAB_VERTICES */
int  mokab_mesh_get_perm(const mokab_mesh *mesh, int kind, int32_t *perm_out);
int  mokab_mesh_device_bytes(const mokab_mesh *mesh, int64_t *out);

/* ---- state: PrognosticVars + DiagnosticVars + TendencyVars on the backend ------------------- */
/* Zero-initialised (DiagnosticVars.jl:90-93, TendencyVars.jl:61-62); two time levels. */
int  mokab_state_create(mokab_ctx *ctx, const mokab_mesh *mesh, int dtype, mokab_state **out);
/* The same with nVertLevels >= 1 levels (VerticalMesh.nVertLevels, VertMesh.jl:3-17): layerThickness / normalVelocity and the
 * Diag / Tend arrays are the reference's (nVertLevels, n) column-major arrays (level fastest, PrognosticVars.jl:10-16), ssh
 * stays (nCells).  The reference's kernels carry the level loops (pressure_gradient.jl:61-64, horizontal_advection_and_
 * coriolis.jl:69-73, horizontal_advection.jl:60-66: one pressure gradient for the column, Coriolis and thickness flux per
 * level) but its drivers fill level 1 only (DiagnosticVars.jl:158-173, time_integration.jl:205-212); for nVertLevels > 1 the
 * semantics are project-defined: every level is stepped, ssh = sum over the levels of layerThickness -
 * restingThicknessSum.  Float64; supported: state set/get, the src/ocn entry points, both steppers (RungeKutta4 fused --
 * static data read once per column -- and unfused; ForwardEuler as the reference's kernel sequence), mokab_reduce, and on
 * decomposed meshes mokab_timestep_rk4_decomposed (whole-part stage launches, ONE halo message of nVertLevels + 1 planes per
 * stage -- the levels of (layerThickness, normalVelocity) and the free surface --, packed exchange whatever the halo mode,
 * captured graphs) and mokab_reduce_decomposed; the reverse mode of RungeKutta4 on undecomposed meshes (mokab_tape_begin ...
 * mokab_adjoint_rk4: the forward recompute is the column kernel, every adjoint stage the single-level gather kernel per level
 * with the pressure term -- one gradient per column -- taken from the level sum of kbar_u; MOKAB_D_* fields are (nVertLevels, n)
 * like the state's), on decomposed meshes too (one K + 1 plane message per halo copy of the sweep; checked with emulated ranks
 * only).  Not supported: the ForwardEuler reverse mode, the staged entry points, ForwardEuler on decomposed meshes. */
int  mokab_state_create_levels(mokab_ctx *ctx, const mokab_mesh *mesh, int dtype, int nVertLevels, mokab_state **out);
int  mokab_state_levels(const mokab_state *state, int *nVertLevels);
int  mokab_state_destroy(mokab_state *state);
/* Host arrays hold the state's dtype, caller numbering.  Setting MOKAB_LAYER_THICKNESS does NOT touch
 * ssh (the reference reads both from file, PrognosticVars.jl:85-99).  `set` of a time-level-`end`
 * prognostic field also fills the "previous" level, like the deepcopy at PrognosticVars.jl:49-53. */
int  mokab_state_set(mokab_state *state, int field, const void *host);
int  mokab_state_get(mokab_state *state, int field, void *host);
/* Pipelined variants for callers that stream states through the device (ensembles, per-step output,
 * the reference's write_netcdf adapt-to-CPU at src/infra/OutPut.jl:122-124 without stalling the loop):
 * the PCIe copy runs on a copy stream of the state, only the (un)permute kernel is ordered with the
 * context's stream, so the upload of the next inputs and the download of the last result overlap the
 * step kernels.  `host_pinned` must be page-locked (mokab_host_alloc) and stay valid -- and, for get,
 * unread -- until mokab_state_synchronize.  set_async does not touch the "previous" time level. */
int  mokab_state_set_async(mokab_state *state, int field, const void *host_pinned);
int  mokab_state_get_async(mokab_state *state, int field, void *host_pinned);
int  mokab_state_synchronize(mokab_state *state);

/* ---- src/ocn entry points (operator level; reference operation order, Float64 bit-faithful) --- */
/* diagnostic_compute!(Mesh, Diag, Prog)                      src/ocn/DiagnosticVars.jl:108-117     */
int  mokab_diagnostic_compute(mokab_state *state);
/* The same four diagnostics of the current state WITHOUT the reference's ordering artefacts: layerThicknessEdge and
 * thicknessFlux of this state (the reference's flux lags one call, DiagnosticVars.jl:112-116), relativeVorticity zeroed
 * before CurlOnVertex accumulates (Operators.jl:135 is commented out in the reference) -- for output / analysis. */
int  mokab_diagnostic_compute_consistent(mokab_state *state);
/* computeNormalVelocityTendency!(Tend, Prog, Diag, Mesh, Config)   .../normalVelocity.jl:21-53     */
int  mokab_compute_normal_velocity_tendency(mokab_state *state);
/* computeLayerThicknessTendency!(Tend, Prog, Diag, Mesh, Config)   .../layerThickness.jl:14-28     */
int  mokab_compute_layer_thickness_tendency(mokab_state *state);
/* Stand-alone operators on host Float64 arrays (src/ocn/Operators.jl:46-74, 102-120, 151-177, 179-199) */
int  mokab_gradient_on_edge(mokab_ctx *ctx, const mokab_mesh *mesh, const double *scalar_cell, double *grad_edge);
int  mokab_divergence_on_cell(mokab_ctx *ctx, const mokab_mesh *mesh, const double *vec_edge, double *div_cell);
int  mokab_curl_on_vertex(mokab_ctx *ctx, const mokab_mesh *mesh, const double *vec_edge, double *curl_vertex_inout);
int  mokab_interpolate_cell2edge(mokab_ctx *ctx, const mokab_mesh *mesh, const double *cell_value, double *edge_value);
/* Reverse mode of the two operators the reference differentiates with Enzyme (test/enzyme/
 * test_Enzyme_Operators.jl:40-125 GradientOnEdge!, :127-227 DivergenceOnCell!): given the adjoint of the output,
 * return the adjoint of the input (the `Duplicated(Scalar, d_Scalar)` / `Duplicated(VecEdge, d_VecEdge)` shadows). */
int  mokab_gradient_on_edge_vjp(mokab_ctx *ctx, const mokab_mesh *mesh, const double *d_grad_edge, double *d_scalar_cell);
int  mokab_divergence_on_cell_vjp(mokab_ctx *ctx, const mokab_mesh *mesh, const double *d_div_cell, double *d_vec_edge);

/* ---- src/forward entry points ------------------------------------------------------------------ */
/* ocn_run_loop + ocn_timestep(::ForwardEuler): `nsteps` steps of src/forward/time_integration.jl:
 * 150-193 with the reference's semantics (lagged hEdge, accumulating vorticity), Float64 bit-faithful.  On hexagonal
 * single-domain meshes one fused kernel per step (+ the curl when the mesh has vertex arrays): the new state goes to the
 * other time level, thicknessFlux / velocityDivCell / tend* are re-created on demand when read; elsewhere, and in the
 * _unfused variant, the reference's kernel sequence (one kernel per reference kernel). */
int  mokab_timestep_forward_euler(mokab_state *state, double dt, int64_t nsteps);
int  mokab_timestep_forward_euler_unfused(mokab_state *state, double dt, int64_t nsteps);
/* ocn_run_loop + ocn_timestep(::RungeKutta4) as intended by time_integration.jl:61-148; `impl` is
 * MOKAB_RK4_FUSED or MOKAB_RK4_UNFUSED.  On return Prog.*[end] is the new state, Prog.*[1] the state
 * one step earlier, ssh = layerThickness - restingThicknessSum. */
int  mokab_timestep_rk4(mokab_state *state, double dt, int64_t nsteps, int impl);
/* sumArray replacement (deterministic two-level reduction) over the OWNED entities; result always Float64. */
int  mokab_reduce(mokab_state *state, int which, double *out);

/* ---- reverse mode: what the reference gets from Enzyme ----------------------------------------------
 * `autodiff(Reverse, ocn_run_loop, Duplicated(Prog, d_Prog), ...)` (test/enzyme/test_Enzyme_end2end.jl:30-110,
 * ext/MPASEnzymeExt.jl) -- NaN on CUDA in the reference (test_Enzyme_end2end.jl:182-186).  Here: a hand-written
 * discrete adjoint of the fused RungeKutta4 path (csrc/kernels_adjoint.cuh).  Usage:
 *   mokab_tape_begin(state, nsteps);  mokab_timestep_rk4(state, dt, nsteps, MOKAB_RK4_FUSED);
 *   mokab_adjoint_seed(state, MOKAB_SUM_SSH2)      (or mokab_state_set of the MOKAB_D_* fields);
 *   mokab_adjoint_rk4(state);  mokab_state_get(state, MOKAB_D_NORMAL_VELOCITY / MOKAB_D_LAYER_THICKNESS, ...)
 * leaves dJ/d(initial normalVelocity, layerThickness) in the shadow fields. */
/* Start recording: every following RK4_FUSED step stores its input state (max_steps * (nEdges + nCells) elements).
 * max_steps = 0 is allowed: the reverse sweep of an empty tape returns the seed (the gradient of the objective at the
 * initial state). */
int  mokab_tape_begin(mokab_state *state, int64_t max_steps);
int  mokab_tape_length(mokab_state *state, int64_t *out);
/* d_ssh = dJ/dssh of the current state for J = `which` (MOKAB_SUM_SSH2: sumArray, run_loop.jl:47-51); d_u = d_h = 0 */
int  mokab_adjoint_seed(mokab_state *state, int which);
/* Reverse sweep over the recorded steps (newest first); stops recording and empties the tape.  d_ssh is folded
 * into d_layerThickness first (ssh = layerThickness - restingThicknessSum, time_integration.jl:205-212). */
int  mokab_adjoint_rk4(mokab_state *state);
/* The same for ForwardEuler -- the stepper test_Enzyme_end2end.jl:78-96 actually differentiates -- including its ordering
 * artefact (the thickness flux of step n uses the layerThicknessEdge left by step n-1, DiagnosticVars.jl:112-116).  While
 * recording, every mokab_timestep_forward_euler step stores its normalVelocity and the lagged layerThicknessEdge
 * (2 * max_steps * nEdges elements).  On return MOKAB_D_NORMAL_VELOCITY / MOKAB_D_LAYER_THICKNESS / MOKAB_D_SSH hold
 * dJ/d(initial normalVelocity, layerThickness, ssh): ssh is an input of its own here (only the first step's pressure
 * gradient reads the array), exactly as in Enzyme's d_Prog.ssh[end].  A tape holds steps of one stepper only. */
int  mokab_adjoint_forward_euler(mokab_state *state);
/* Domain-decomposed states (mokab_decomp_setup): the same calls.  mokab_timestep_rk4_decomposed /
 * mokab_timestep_forward_euler_decomposed record every rank's part of the trajectory (halo copies included) while a tape is
 * open; mokab_adjoint_seed seeds the owned cells; mokab_adjoint_rk4 / mokab_adjoint_forward_euler run the same gather-form
 * kernels over the owned entities and, after every reversed stage / step, refresh the halo copies of what the next one
 * gathers (recomputed stage states, kbar, lambda; ForwardEuler: lambda_u, lambda_hEdge, q) through the packed exchange of
 * the communicator, over the SAME send / receive lists as the forward exchange -- in gather form no transposed,
 * accumulating exchange is needed.  All ranks must make these calls together (they are collective).  The gradient comes
 * back per rank on its owned entities; J is mokab_reduce_decomposed(MOKAB_SUM_SSH2) before the seed. */

/* ---- staged RungeKutta4 for domain-decomposed runs (one process per GPU) --------------------------
 * The same fused stage kernel, launched per stage and per part so the host can overlap the halo
 * exchange of stage s with the interior of stage s.  `cuda_stream` NULL = the
 * context's stream.  Entities are addressed in the combined local index space [cells | edges]
 * (edge k -> nCells + k), caller numbering, 0-based. */
enum { MOKAB_PART_ALL = 0, MOKAB_PART_INTERIOR = 1, MOKAB_PART_BOUNDARY = 2,
       /* the boundary blocks with the direct-store halo exchange folded in (after mokab_p2p_setup): wait for the neighbours'
        * previous stage, compute, store every value a neighbour needs into its memory, tick its arrival counter */
       MOKAB_PART_BOUNDARY_PUSH = 3,
       /* every block, with the direct-store exchange folded in: ONE launch per stage -- for parts so small that launch
        * latency, not bytes, sets the stage time (after mokab_p2p_setup) */
       MOKAB_PART_ALL_PUSH = 4 };
/* send_idx: owned entities whose values neighbours need; recv_idx: halo entities, in message order.
 * Blocks that hold a send entity join the BOUNDARY part, so a message can be packed as soon as the
 * boundary launch of a stage has finished.  Call before creating states on the mesh. */
int  mokab_halo_setup(mokab_mesh *mesh, int64_t n_send, const int32_t *send_idx, int64_t n_recv, const int32_t *recv_idx);
/* stage 0 = the current state Prog.*[end]; stage s = 1..4 = the output of RK stage s of the step in flight (4 is also the
 * (layerThickness, normalVelocity) a staged ForwardEuler step wrote); stage 5 = the (ssh, layerThicknessEdge) it wrote.
 * pack: device message buffer (state dtype, n_send elements) <- values; unpack: halo slots <- message. */
int  mokab_halo_pack(mokab_state *state, int stage, void *send_buf_device, void *cuda_stream);
int  mokab_halo_unpack(mokab_state *state, int stage, const void *recv_buf_device, void *cuda_stream);
/* one fused RK stage (1..4) over MOKAB_PART_ALL / _INTERIOR / _BOUNDARY blocks */
int  mokab_rk4_stage(mokab_state *state, double dt, int stage, int part, void *cuda_stream);
/* after stage 4 (and its exchange): the other time level becomes Prog.*[end] */
int  mokab_rk4_finish_step(mokab_state *state);
/* ---- staged ForwardEuler for domain-decomposed runs: ocn_timestep(::ForwardEuler) `time_integration.jl:150-193` ----------
 * One step of the fused ForwardEuler kernel over the selected blocks (MOKAB_PART_ALL / _INTERIOR / _BOUNDARY), reading
 * Prog.*[end] and writing the other time level for the OWNED cells and edges: normalVelocity, layerThickness, ssh and the
 * layerThicknessEdge of the state it read (the lagged value the next step's thicknessFlux uses, DiagnosticVars.jl:108-117).
 * Then: mokab_halo_pack/unpack with stage 4 and stage 5 (exchange in between), then mokab_forward_euler_finish_step.
 * Float64; connectivity rows of at most (10, 6) or (12, 7) entries; same bits as mokab_timestep_forward_euler on the whole mesh. */
int  mokab_forward_euler_stage(mokab_state *state, double dt, int part, void *cuda_stream);
int  mokab_forward_euler_finish_step(mokab_state *state);
/* ssh = layerThickness - restingThicknessSum on both time levels (Update_ssh!, time_integration.jl:205-212) */
int  mokab_refresh_ssh(mokab_state *state, void *cuda_stream);
/* ---- halo exchange by direct stores into the peers' memory (csrc/kernels_p2p.cuh) ---------------------------------------
 * The alternative to pack -> all-to-all -> unpack named by BASELINE.json's north_star ("or direct P2P stores"): per RK stage
 * the sender writes its halo values straight into the receivers' state arrays over NVLink and bumps a per-sender arrival
 * counter there (mokab_halo_push); the receiver's next boundary launch is preceded by a one-warp kernel that waits for the
 * counters of the ranks it receives from (mokab_halo_wait).  No message buffers, no collective, graph-capturable.  Set-up,
 * once per state, after mokab_halo_setup:
 *   1. every rank:  mokab_halo_recv_device_indices(mesh, idx)  -- where its halo entities live in ITS arrays, message order
 *      (c >= 0: cell c; d < 0: edge -d - 1), and hands each sender the segment that sender fills;
 *   2. every rank:  mokab_p2p_export(state, rank, blob)  (mokab_p2p_blob_size bytes: addresses + CUDA IPC handles of the
 *      state arrays and the arrival counters), all blobs gathered on every rank;
 *   3. every rank:  mokab_p2p_setup(state, rank, nranks, blobs, receivers..., push_counts, dst_idx, senders...) with, per
 *      receiver in message order, the indices obtained in step 1.
 * Ranks of other processes are mapped with cudaIpcOpenMemHandle, ranks emulated inside one process use the addresses. */
int  mokab_halo_recv_device_indices(const mokab_mesh *mesh, int32_t *out);
int  mokab_p2p_blob_size(int64_t *out);
int  mokab_p2p_export(mokab_state *state, int rank, void *blob);
int  mokab_p2p_setup(mokab_state *state, int rank, int nranks, const void *blobs, int n_receivers, const int32_t *receiver_ranks,
                     const int64_t *push_counts, const int32_t *dst_idx, int n_senders, const int32_t *sender_ranks);
/* stage as in mokab_halo_pack */
int  mokab_halo_push(mokab_state *state, int stage, void *cuda_stream);
int  mokab_halo_wait(mokab_state *state, void *cuda_stream);
/* With MOKAB_PART_BOUNDARY_PUSH launches there is nothing else to call per stage; after the LAST stage enqueued (before the
 * host, an upload or anything else touches the halo slots) this waits until the neighbours' last stores have arrived. */
int  mokab_halo_wait_arrivals(mokab_state *state, void *cuda_stream);
/* Unmap the peers' memory (every rank, after the last step and before any rank destroys its state: a mapping must not
 * outlive the allocation it maps).  The state is usable again after another mokab_p2p_setup. */
int  mokab_p2p_close(mokab_state *state);
/* 1 if a wait ever timed out (~2 s: a peer died or the ranks' schedules diverged); the GPU is never left spinning */
int  mokab_p2p_error(mokab_state *state, int *out);
/* ---- domain-decomposed stepping inside the library (csrc/comm.cuh, csrc/decomposed.cuh) --------------------------------------
 * One process per GPU.  The reference has no multi-device path (its driver builds one backend, src/driver/mpas_ocean.jl:28);
 * this stands behind BASELINE.json's north_star: "halo exchange runs as NCCL send/recv (or direct P2P stores) over NVLink,
 * overlapped with interior-cell compute".  The library owns the exchange, the two streams, the events and the captured step
 * graphs; the host program only brings the ranks together, the way NCCL itself asks:
 *   rank 0:      mokab_comm_get_unique_id(id)           (MOKAB_COMM_ID_BYTES bytes; hand them to every rank: MPI.jl, a file, ...)
 *   every rank:  mokab_comm_init(ctx, id, rank, nranks, &comm)
 *                mokab_mesh_create(local mesh, nCellsOwned / nEdgesOwned) ; mokab_halo_setup(mesh, send list, recv list)
 *                mokab_state_create ; mokab_state_set ...
 *                mokab_decomp_setup(state, comm, send_counts, recv_counts, MOKAB_HALO_NCCL, 0)
 *                mokab_timestep_rk4_decomposed(state, dt, nsteps)   |   mokab_timestep_forward_euler_decomposed(...)
 *                mokab_reduce_decomposed(state, MOKAB_SUM_SSH2, &s) ;  mokab_state_get ...
 *                mokab_decomp_close(state) ; mokab_state_destroy ; mokab_comm_destroy(comm)
 * send_counts[q] / recv_counts[q]: how many consecutive entries of the send / recv list of mokab_halo_setup go to / come from
 * rank q (the lists are ordered by rank).  Collective calls: comm_init, decomp_setup, the timestep calls, reduce_decomposed,
 * comm_barrier / allreduce / allgather, decomp_close -- every rank makes them in the same order. */
#define MOKAB_COMM_ID_BYTES 128
typedef struct mokab_comm mokab_comm;
enum { MOKAB_HALO_NCCL = 0,       /* pack -> ncclSend/ncclRecv per neighbour (one group) -> unpack, on the halo stream             */
       MOKAB_HALO_P2P = 1,        /* direct stores into the neighbours' state arrays (CUDA IPC) + arrival counters: push / wait kernels */
       MOKAB_HALO_P2P_FUSED = 2,  /* the same stores issued by the boundary blocks themselves (MOKAB_PART_BOUNDARY_PUSH)            */
       MOKAB_HALO_P2P_LL = 3      /* flag-in-data: every value travels as 8-byte {32 data bits, 32-bit exchange number} packets into a
                                     receive area of the neighbour (CUDA IPC), whose wait kernel polls the packets and unpacks them --
                                     no fence, no counter on the sending side (kernels_p2p.cuh); also carries ForwardEuler's two
                                     messages per step and the halo copies of the reverse sweep                                     */ };
enum { MOKAB_DECOMP_NO_OVERLAP = 1u, /* exchange after each whole stage on one stream (diagnostic)                                   */
       MOKAB_DECOMP_NO_GRAPH = 2u    /* launch every step from the host instead of replaying captured 1- / 2-step graphs (diagnostic) */ };
int  mokab_comm_get_unique_id(void *id_out);
int  mokab_comm_init(mokab_ctx *ctx, const void *id, int rank, int nranks, mokab_comm **out);
int  mokab_comm_destroy(mokab_comm *comm);
int  mokab_comm_rank(const mokab_comm *comm, int *rank, int *nranks);
/* host-level collectives for the driver around the steps (synchronous): op 0 = sum (in rank order: the same bits on every
 * rank), 1 = max, 2 = min; allgather: `bytes` bytes from every rank, rank-major */
int  mokab_comm_barrier(mokab_comm *comm);
int  mokab_comm_allreduce_f64(mokab_comm *comm, double *inout, int64_t n, int op);
int  mokab_comm_allgather_bytes(mokab_comm *comm, const void *mine, int64_t bytes, void *all);
int  mokab_decomp_setup(mokab_state *state, mokab_comm *comm, const int64_t *send_counts, const int64_t *recv_counts, int halo_mode,
                        uint32_t flags);
int  mokab_decomp_set_flags(mokab_state *state, uint32_t flags);
/* ocn_run_loop + ocn_timestep over the ranks of `comm`: the same results, bit for bit, as the single-domain entry points on
 * the undecomposed mesh (time_integration.jl:61-148 / :150-193).  Asynchronous like their single-domain counterparts. */
int  mokab_timestep_rk4_decomposed(mokab_state *state, double dt, int64_t nsteps);
int  mokab_timestep_forward_euler_decomposed(mokab_state *state, double dt, int64_t nsteps);
/* mokab_reduce over the owned entities of every rank, summed in rank order (sumArray, run_loop.jl:47-51) */
int  mokab_reduce_decomposed(mokab_state *state, int which, double *out);
/* drain both streams of the state; fails if a direct-store halo wait timed out (MOKAB_P2P_TIMEOUT_S, default 20 s) */
int  mokab_decomp_synchronize(mokab_state *state);
int  mokab_decomp_close(mokab_state *state);

/* Tuning switches of the fused stage kernel (process-wide; no reference counterpart -- the reference's only knob is the
 * workgroup size hard-coded at each launch, e.g. src/forward/time_integration.jl:181).  Results are bit-identical under every
 * setting.  "stage_prefetch": bit 0 = a block pulls the streams of its later iterations into L2 at entry, bit 1 = those of
 * the block launched "stage_prefetch_distance" blocks later (0 = one wave of resident blocks); "stage_tma": 1 / 2 = the
 * Coriolis weights through bulk asynchronous copies (slot-major rows / a block-major copy), 3 = through per-thread cp.async into
 * shared memory (one more resident block per SM); "stage_auto": with stage_tma = 3, every Float64 launch takes whichever of
 * the cp.async and the plain kernel a wave-quantisation model prefers for its number of blocks (small grids -- 512 x 512, the
 * parts of an 8-GPU run -- are up to 25 % faster on the plain kernel, which keeps one block per SM fewer resident);
 * "stage_pdl": stage launches carry the programmatic-stream-serialization attribute, so the static half of stage s + 1
 * overlaps the tail of stage s.  Defaults: stage_tma = 3, stage_prefetch = 1, stage_auto = 1, stage_pdl = 0
 * (profiles/README.md r02d, r02i); the environment overrides them (MOKAB_STAGE_PREFETCH, MOKAB_STAGE_PREFETCH_DISTANCE,
 * MOKAB_STAGE_TMA, MOKAB_STAGE_AUTO, MOKAB_STAGE_PDL). */
int  mokab_set_option(const char *name, int64_t value);
int  mokab_get_option(const char *name, int64_t *value);
/* Stage timeline (diagnostic; TRACE builds only -- `make -C mpas-ocean.jl_b200 libmoka_b200_trace.so`; the default build returns
 * an error).  Between mokab_trace_begin(ctx, capacity) and mokab_trace_read every block of the stage and halo kernels appends a
 * 48-byte record {u32 kind, block, grid, pad; u64 globaltimer at entry, at exit, after the wait for the peers} to a ring of
 * `capacity` records; kind = RK stage | part << 4 for stage launches, 100 / 101 = halo push / wait kernels, 110 / 111 = pack /
 * unpack.  `count` returns how many records were written in all.  tools/trace_stages.py prints the timeline of a step. */
int  mokab_trace_begin(mokab_ctx *ctx, int64_t capacity);
int  mokab_trace_read(mokab_ctx *ctx, void *records, int64_t max_records, int64_t *count);
/* number of interior / boundary blocks of the fused kernel (diagnostic) */
int  mokab_mesh_block_counts(const mokab_mesh *mesh, int64_t *interior, int64_t *boundary);
/* number of fused-kernel blocks, and how many of them rebuild edgesOnEdge from edgesOnCell (diagnostic) */
int  mokab_mesh_derived_blocks(const mokab_mesh *mesh, int64_t *blocks, int64_t *derived);

#ifdef __cplusplus
}
#endif
#endif /* MOKA_B200_H */
