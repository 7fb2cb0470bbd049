"""Host-side mirror of the reference's Julia API for the forward hot path, over the C ABI.

Julia is not installed in this image, so the host side above
libmoka_b200.so is written in Python; the Julia shim that does the same with `ccall` ships as
source in ../julia/MokaB200.jl.  Names, argument order and error behaviour follow the reference
(`!` dropped from mutating function names):

  B200, on_architecture, check_*_args          src/Architectures.jl:12-46
  Mesh (HorzMesh + VerticalMesh)               src/infra/MPASMesh/{HorzMesh,VertMesh,MPASMesh}.jl
  PrognosticVars / DiagnosticVars / TendencyVars   src/ocn/{PrognosticVars,DiagnosticVars}.jl, Tendencies/TendencyVars.jl
  diagnostic_compute, computeNormalVelocityTendency, computeLayerThicknessTendency   src/ocn/**
  GradientOnEdge, DivergenceOnCell, CurlOnVertex, interpolateCell2Edge   src/ocn/Operators.jl
  ocn_timestep, ocn_run_loop, ForwardEuler, RungeKutta4        src/forward/{time_integration,run_loop}.jl
  ocn_init_alarms (dt rule)                    src/forward/init.jl:111-127

All device work happens in libmoka_b200.so; numpy is only the host array container.
"""
from __future__ import annotations

import ctypes as C
import weakref

import numpy as np

from . import _lib as L
from ._lib import MokaError

GRAVITY = 9.80616


# ---- src/Architectures.jl ---------------------------------------------------------------------------
class B200:
    """The `B200` architecture: a KA.GPU-like backend bound to one CUDA device (a context)."""

    def __init__(self, device: int = 0):
        h = C.c_void_p()
        L.check(L.lib().mokab_init(device, C.byref(h)))
        self.handle = h
        self.device = device
        self._fin = weakref.finalize(self, L.lib().mokab_finalize, h)

    def synchronize(self) -> None:                      # KA.synchronize(backend)
        L.check(L.lib().mokab_synchronize(self.handle))

    def set_stream(self, cuda_stream: int | None) -> None:
        L.check(L.lib().mokab_set_stream(self.handle, C.c_void_p(cuda_stream or 0)))

    def timer_start(self) -> None:
        L.check(L.lib().mokab_timer_start(self.handle))

    def timer_stop(self) -> float:
        ms = C.c_double()
        L.check(L.lib().mokab_timer_stop(self.handle, C.byref(ms)))
        return ms.value

    def launch_count(self) -> int:
        n = C.c_int64()
        L.check(L.lib().mokab_launch_count(self.handle, C.byref(n)))
        return n.value

    def pinned(self, shape, dtype=np.float64) -> np.ndarray:
        """Page-locked host array (staging buffer for state set/get); freed with the backend."""
        count = int(np.prod(shape))
        n = count * np.dtype(dtype).itemsize
        p = C.c_void_p()
        L.check(L.lib().mokab_host_alloc(C.byref(p), n))
        weakref.finalize(self, L.lib().mokab_host_free, p)
        buf = (C.c_char * max(n, 1)).from_address(p.value)
        return np.frombuffer(buf, dtype=dtype, count=count).reshape(shape)


def check_typeof_args(args) -> None:                    # Architectures.jl:19-25
    if len({type(a).__name__ for a in args}) > 1:
        raise MokaError("Input arguments must be of all the same type")


def check_eltype_args(args):                            # Architectures.jl:35-46
    if len({np.asarray(a).dtype for a in args}) > 1:
        raise MokaError("All input arguments must have the same eltype")
    return np.asarray(args[0]).dtype


# ---- src/infra/MPASMesh -------------------------------------------------------------------------------
class Mesh:
    """Mesh{HorzMesh, VerticalMesh} on the B200 backend (MPASMesh.jl:19-29).

    `fields` is a dict with the MPAS variable names the reference's readers use
    (HorzMesh.jl:166-290, VertMesh.jl:46-82), reference layouts; it stands behind
    `ReadHorzMesh(path; backend)` + `VerticalMesh(path, hmesh; backend)` + the sign fields of
    HorzMesh.jl:292-332 (computed by the library when `edgeSignOnCell` is not supplied).
    """

    def __init__(self, fields: dict, backend: B200, renumber: bool = True, explicit_eoe: bool = False, keep_widths: bool = False,
                 edges_by_cell: bool = False):
        if not isinstance(backend, B200):
            raise MokaError("Mesh: backend must be a B200 architecture")
        self.backend = backend
        self.nCells, self.nEdges = int(fields["nCells"]), int(fields["nEdges"])
        self.nVertices = int(fields.get("nVertices", 0)) if fields.get("edgesOnVertex") is not None else 0
        self.maxEdges, self.maxEdges2 = int(fields["maxEdges"]), int(fields["maxEdges2"])
        self.vertexDegree = int(fields.get("vertexDegree", 3))
        self.nVertLevels = int(fields.get("nVertLevels", 1))          # VertMesh.nVertLevels: the level count of the states on this mesh
        d = L.MeshDesc()
        d.nCells, d.nEdges, d.nVertices = self.nCells, self.nEdges, self.nVertices
        d.maxEdges, d.maxEdges2, d.vertexDegree = self.maxEdges, self.maxEdges2, self.vertexDegree
        self.nCellsOwned = int(fields.get("nCellsOwned", 0)) or self.nCells
        self.nEdgesOwned = int(fields.get("nEdgesOwned", 0)) or self.nEdges
        d.nCellsOwned, d.nEdgesOwned = self.nCellsOwned, self.nEdgesOwned
        keep = []
        rts = fields.get("restingThicknessSum")
        if rts is None:                                               # VertMesh.jl:73
            rts = np.asarray(fields["restingThickness"], dtype=np.float64).reshape(self.nCells, -1).sum(axis=1)
        src = dict(fields)
        src["restingThicknessSum"] = rts
        if self.nVertices == 0:
            for k in ("edgesOnVertex", "edgeSignOnVertex", "areaTriangle", "verticesOnEdge"):
                src[k] = None
        for name, ptype in L._DESC_PTRS:
            a = src.get(name)
            if a is None:
                setattr(d, name, None)
                continue
            a = np.ascontiguousarray(a, dtype=np.int32 if ptype is L._I32P else np.float64)
            keep.append(a)
            setattr(d, name, a.ctypes.data_as(ptype))
        h = C.c_void_p()
        L.check(L.lib().mokab_mesh_create(backend.handle, C.byref(d), (L.MESH_RENUMBER if renumber else 0) | (L.MESH_EXPLICIT_EOE if explicit_eoe else 0) | (L.MESH_KEEP_WIDTHS if keep_widths else 0) | (L.MESH_EDGES_BY_CELL if edges_by_cell else 0), C.byref(h)))
        self.handle = h
        self._fin = weakref.finalize(self, L.lib().mokab_mesh_destroy, h)
        self.dcEdge_mean = float(np.mean(fields["dcEdge"]))

    def perm(self, kind: str) -> np.ndarray:
        """perm[new] = old (0-based) of the library's locality renumbering."""
        k, n = {"cells": (L.CELLS, self.nCells), "edges": (L.EDGES, self.nEdges), "vertices": (L.VERTICES, self.nVertices)}[kind]
        out = np.empty(n, np.int32)
        L.check(L.lib().mokab_mesh_get_perm(self.handle, k, out.ctypes.data_as(L._I32P)))
        return out

    def halo_setup(self, send_idx, recv_idx) -> None:
        """Register the halo message layout (combined [cells | edges] local indices, 0-based)."""
        s = np.ascontiguousarray(send_idx, dtype=np.int32)
        r = np.ascontiguousarray(recv_idx, dtype=np.int32)
        L.check(L.lib().mokab_halo_setup(self.handle, s.size, s.ctypes.data_as(L._I32P), r.size, r.ctypes.data_as(L._I32P)))

    def block_counts(self):
        a, b = C.c_int64(), C.c_int64()
        L.check(L.lib().mokab_mesh_block_counts(self.handle, C.byref(a), C.byref(b)))
        return a.value, b.value

    def derived_blocks(self):
        """(blocks of the fused kernel, blocks that rebuild edgesOnEdge from edgesOnCell instead of reading it)."""
        a, b = C.c_int64(), C.c_int64()
        L.check(L.lib().mokab_mesh_derived_blocks(self.handle, C.byref(a), C.byref(b)))
        return a.value, b.value

    def device_bytes(self) -> int:
        n = C.c_int64()
        L.check(L.lib().mokab_mesh_device_bytes(self.handle, C.byref(n)))
        return n.value


# ---- state structs ------------------------------------------------------------------------------------
class _DeviceState:
    def __init__(self, mesh: Mesh, dtype, levels: int = 1):
        self.mesh = mesh
        self.levels = int(levels)
        self.np_dtype = np.dtype(dtype)
        if self.np_dtype not in (np.dtype(np.float64), np.dtype(np.float32)):
            raise MokaError("state eltype must be Float64 or Float32")
        h = C.c_void_p()
        L.check(L.lib().mokab_state_create_levels(mesh.backend.handle, mesh.handle,
                                                  L.F64 if self.np_dtype == np.float64 else L.F32, self.levels, C.byref(h)))
        self.handle = h
        self._fin = weakref.finalize(self, L.lib().mokab_state_destroy, h)

    def set(self, field: int, host) -> None:
        a = np.ascontiguousarray(host, dtype=self.np_dtype).reshape(-1)
        n = self._len(field)
        if a.size != n:
            raise MokaError(f"field {field}: expected {n} elements, got {a.size}")
        L.check(L.lib().mokab_state_set(self.handle, field, a.ctypes.data_as(C.c_void_p)))

    def get(self, field: int, out: np.ndarray | None = None) -> np.ndarray:
        n = self._len(field)
        if out is None:
            out = np.empty(n, self.np_dtype)
        L.check(L.lib().mokab_state_get(self.handle, field, out.ctypes.data_as(C.c_void_p)))
        # multi-level fields come back as (n, nVertLevels): the memory of the reference's (nVertLevels, n) column-major array
        return out.reshape(-1, self.levels) if self.levels > 1 and n != self._entities(field) else out

    def set_async(self, field: int, host_pinned: np.ndarray) -> None:
        """Enqueue the upload of a page-locked host array (B200.pinned); returns before the copy has run."""
        if host_pinned.dtype != self.np_dtype or host_pinned.size != self._len(field) or not host_pinned.flags.c_contiguous:
            raise MokaError(f"set_async: field {field} needs a contiguous {self.np_dtype} array of {self._len(field)} elements")
        L.check(L.lib().mokab_state_set_async(self.handle, field, host_pinned.ctypes.data_as(C.c_void_p)))

    def get_async(self, field: int, out_pinned: np.ndarray) -> None:
        """Enqueue the download into a page-locked host array; valid after `synchronize()`."""
        if out_pinned.dtype != self.np_dtype or out_pinned.size != self._len(field) or not out_pinned.flags.c_contiguous:
            raise MokaError(f"get_async: field {field} needs a contiguous {self.np_dtype} array of {self._len(field)} elements")
        L.check(L.lib().mokab_state_get_async(self.handle, field, out_pinned.ctypes.data_as(C.c_void_p)))

    def synchronize(self) -> None:
        L.check(L.lib().mokab_state_synchronize(self.handle))

    def _entities(self, field: int) -> int:
        m = self.mesh
        if field in (L.SSH, L.LAYER_THICKNESS, L.SSH_PREV, L.LAYER_THICKNESS_PREV, L.VELOCITY_DIV_CELL, L.TEND_LAYER_THICKNESS,
                     L.D_SSH, L.D_LAYER_THICKNESS):
            return m.nCells
        if field == L.RELATIVE_VORTICITY:
            return m.nVertices
        return m.nEdges

    def _len(self, field: int) -> int:
        n = self._entities(field)
        return n if field in (L.SSH, L.SSH_PREV, L.D_SSH) else n * self.levels


class PrognosticVars:
    """PrognosticVars(ssh, normalVelocity, layerThickness, nTimeLevels) (PrognosticVars.jl:28-56).

    Two time levels live on the device; `ssh`, `normalVelocity`, `layerThickness` return the
    `[end]` level as host arrays, `*_prev` the `[1]` level.
    """

    def __init__(self, ssh, normalVelocity, layerThickness, nTimeLevels: int, mesh: Mesh, dtype=None):
        args = (ssh, normalVelocity, layerThickness)
        check_typeof_args(args)
        et = check_eltype_args(args)
        if nTimeLevels != 2:
            raise MokaError("nTimeLevels must be <= 2" if nTimeLevels > 2 else "nTimeLevels must be 2")  # time_integration.jl:23
        # multi-level states: normalVelocity / layerThickness of shape (n, nVertLevels) -- the memory of the reference's
        # (nVertLevels, n) arrays (PrognosticVars.jl:10-16) -- and ssh of shape (nCells)
        nv = np.asarray(normalVelocity)
        levels = int(nv.shape[1]) if nv.ndim == 2 else 1
        if levels != mesh.nVertLevels and not (levels == 1 and mesh.nVertLevels == 1):
            raise MokaError(f"PrognosticVars: normalVelocity has {levels} levels, the mesh {mesh.nVertLevels}")
        self.dev = _DeviceState(mesh, dtype or et, levels)
        self.mesh = mesh
        # Float32 states carry ONE perturbation variable, ssh = layerThickness - restingThicknessSum (include/moka_b200.h):
        # layerThickness first (a Float32 1000 m + 1 m has lost the wave's low bits), then ssh, which keeps them
        self.dev.set(L.LAYER_THICKNESS, layerThickness)
        self.dev.set(L.NORMAL_VELOCITY, normalVelocity)
        self.dev.set(L.SSH, ssh)

    def upload_async(self, normalVelocity=None, layerThickness=None, ssh=None) -> None:
        """Adapt.adapt(backend, host arrays) without stalling the device: pipelined H2D from page-locked arrays."""
        for field, a in ((L.NORMAL_VELOCITY, normalVelocity), (L.LAYER_THICKNESS, layerThickness), (L.SSH, ssh)):
            if a is not None:
                self.dev.set_async(field, a)

    def download_async(self, ssh=None, normalVelocity=None, layerThickness=None) -> None:
        """Adapt.adapt(CPU(), Prog) (what write_netcdf does, OutPut.jl:122-124) into page-locked arrays, pipelined."""
        for field, a in ((L.SSH, ssh), (L.NORMAL_VELOCITY, normalVelocity), (L.LAYER_THICKNESS, layerThickness)):
            if a is not None:
                self.dev.get_async(field, a)

    def synchronize(self) -> None:
        self.dev.synchronize()

    ssh = property(lambda s: s.dev.get(L.SSH))
    normalVelocity = property(lambda s: s.dev.get(L.NORMAL_VELOCITY))
    layerThickness = property(lambda s: s.dev.get(L.LAYER_THICKNESS))
    ssh_prev = property(lambda s: s.dev.get(L.SSH_PREV))
    normalVelocity_prev = property(lambda s: s.dev.get(L.NORMAL_VELOCITY_PREV))
    layerThickness_prev = property(lambda s: s.dev.get(L.LAYER_THICKNESS_PREV))


class DiagnosticVars:
    """DiagnosticVars(config, mesh) (DiagnosticVars.jl:75-99): zeros on the backend."""

    def __init__(self, prog: PrognosticVars):
        self.dev = prog.dev

    layerThicknessEdge = property(lambda s: s.dev.get(L.LAYER_THICKNESS_EDGE))
    thicknessFlux = property(lambda s: s.dev.get(L.THICKNESS_FLUX))
    velocityDivCell = property(lambda s: s.dev.get(L.VELOCITY_DIV_CELL))
    relativeVorticity = property(lambda s: s.dev.get(L.RELATIVE_VORTICITY))


class TendencyVars:
    """TendencyVars(config, mesh) (TendencyVars.jl:51-67): zeros on the backend."""

    def __init__(self, prog: PrognosticVars):
        self.dev = prog.dev

    tendNormalVelocity = property(lambda s: s.dev.get(L.TEND_NORMAL_VELOCITY))
    tendLayerThickness = property(lambda s: s.dev.get(L.TEND_LAYER_THICKNESS))


# ---- src/ocn ----------------------------------------------------------------------------------------------
def diagnostic_compute(Mesh_, Diag: DiagnosticVars, Prog: PrognosticVars, consistent: bool = False) -> None:
    """diagnostic_compute!(Mesh, Diag, Prog; backend) (DiagnosticVars.jl:108-117).  `consistent=True` computes the same
    four fields of the current state without the reference's ordering artefacts (lagged flux, never-zeroed vorticity)."""
    fn = L.lib().mokab_diagnostic_compute_consistent if consistent else L.lib().mokab_diagnostic_compute
    L.check(fn(Prog.dev.handle))


def computeNormalVelocityTendency(Tend, Prog, Diag, Mesh_, Config=None) -> None:
    """computeNormalVelocityTendency!(Tend, Prog, Diag, Mesh, Config; backend) (normalVelocity.jl:21-53)."""
    L.check(L.lib().mokab_compute_normal_velocity_tendency(Prog.dev.handle))


def computeLayerThicknessTendency(Tend, Prog, Diag, Mesh_, Config=None) -> None:
    """computeLayerThicknessTendency!(Tend, Prog, Diag, Mesh, Config; backend) (layerThickness.jl:14-28)."""
    L.check(L.lib().mokab_compute_layer_thickness_tendency(Prog.dev.handle))


def _op(fn, mesh: Mesh, src, n_out, out=None):
    a = np.ascontiguousarray(src, dtype=np.float64).reshape(-1)
    if out is None:
        out = np.zeros(n_out)
    L.check(fn(mesh.backend.handle, mesh.handle, L.fptr(a), L.fptr(out)))
    return out


def GradientOnEdge(grad, h_cell, mesh: Mesh):
    """GradientOnEdge!(grad, h, Mesh; backend) (Operators.jl:102-120)."""
    return _op(L.lib().mokab_gradient_on_edge, mesh, h_cell, mesh.nEdges, grad)


def DivergenceOnCell(div, vec_edge, temp, mesh: Mesh):
    """DivergenceOnCell!(DivCell, VecEdge, temp, Mesh; backend) (Operators.jl:46-74); `temp` is unused."""
    return _op(L.lib().mokab_divergence_on_cell, mesh, vec_edge, mesh.nCells, div)


def CurlOnVertex(curl, vec_edge, mesh: Mesh):
    """CurlOnVertex!(CurlVertex, VecEdge, Mesh; backend) (Operators.jl:151-177): accumulates into `curl`."""
    return _op(L.lib().mokab_curl_on_vertex, mesh, vec_edge, mesh.nVertices, curl)


def interpolateCell2Edge(edge_value, cell_value, mesh: Mesh):
    """interpolateCell2Edge!(edgeValue, cellValue, Mesh; backend) (Operators.jl:179-199)."""
    return _op(L.lib().mokab_interpolate_cell2edge, mesh, cell_value, mesh.nEdges, edge_value)


def GradientOnEdge_vjp(d_grad, mesh: Mesh):
    """Adjoint of GradientOnEdge!: d_Scalar for a given d_gradNum (autodiff(Reverse, gradient_test, ...),
    test/enzyme/test_Enzyme_Operators.jl:47-64)."""
    return _op(L.lib().mokab_gradient_on_edge_vjp, mesh, d_grad, mesh.nCells)


def DivergenceOnCell_vjp(d_div, mesh: Mesh):
    """Adjoint of DivergenceOnCell!: d_VecEdge for a given d_divNum (test_Enzyme_Operators.jl:130-152)."""
    return _op(L.lib().mokab_divergence_on_cell_vjp, mesh, d_div, mesh.nEdges)


# ---- src/forward ------------------------------------------------------------------------------------------
class ForwardEuler:       # time_integration.jl:4
    pass


class RungeKutta4:        # time_integration.jl:5
    pass


def ocn_timestep(timestep: float, Prog, Diag, Tend, Setup=None, stepper=RungeKutta4, nsteps: int = 1, fused: bool = True) -> None:
    """ocn_timestep(timestep, Prog, Diag, Tend, Setup, ::Type{ForwardEuler|RungeKutta4}; backend)
    (time_integration.jl:61-66,150-156).  `nsteps` > 1 keeps the loop on the device."""
    if stepper is ForwardEuler:
        fn = L.lib().mokab_timestep_forward_euler if fused else L.lib().mokab_timestep_forward_euler_unfused
        L.check(fn(Prog.dev.handle, float(timestep), int(nsteps)))
    elif stepper is RungeKutta4:
        L.check(L.lib().mokab_timestep_rk4(Prog.dev.handle, float(timestep), int(nsteps), L.RK4_FUSED if fused else L.RK4_UNFUSED))
    else:
        raise MokaError("ocn_timestep: unknown time stepper")


def ocn_run_loop(timestep: float, Prog, Diag, Tend, Setup, stepper, nsteps: int, sum_ssh2: bool = False, fused: bool = True):
    """ocn_run_loop(timestep, Prog, Diag, Tend, Setup, Stepper, clock, simulationAlarm, outputAlarm)
    (run_loop.jl:8-45) with the clock/alarms reduced to the number of steps they allow; the second
    method's squared-SSH sum (sumArray, :47-51) is returned when `sum_ssh2`."""
    ocn_timestep(timestep, Prog, Diag, Tend, Setup, stepper, nsteps=nsteps, fused=fused)
    if sum_ssh2:
        return reduce_sum(Prog, "ssh2")
    return None


def reduce_sum(Prog, which: str) -> float:
    out = C.c_double()
    L.check(L.lib().mokab_reduce(Prog.dev.handle, {"ssh2": L.SUM_SSH2, "mass": L.SUM_MASS, "energy": L.SUM_ENERGY}[which], C.byref(out)))
    return out.value


# ---- reverse mode (ext/MPASEnzymeExt.jl, test/enzyme/test_Enzyme_end2end.jl) -------------------------------------
class ShadowPrognosticVars:
    """d_Prog: the shadow of a PrognosticVars that `Duplicated(Prog, d_Prog)` carries through Enzyme's
    `autodiff(Reverse, ocn_run_loop, ...)` (test_Enzyme_end2end.jl:52-96); created zeroed like
    `ocn_init_shadows` (init.jl:32-40).  Lives on the device next to `Prog`."""

    def __init__(self, prog: PrognosticVars):
        self.dev = prog.dev
        for f in (L.D_SSH, L.D_NORMAL_VELOCITY, L.D_LAYER_THICKNESS):
            self.dev.set(f, np.zeros(self.dev._len(f), self.dev.np_dtype))     # (multi-level states: (n, nVertLevels) like the state itself)

    ssh = property(lambda s: s.dev.get(L.D_SSH), lambda s, v: s.dev.set(L.D_SSH, v))
    normalVelocity = property(lambda s: s.dev.get(L.D_NORMAL_VELOCITY), lambda s, v: s.dev.set(L.D_NORMAL_VELOCITY, v))
    layerThickness = property(lambda s: s.dev.get(L.D_LAYER_THICKNESS), lambda s, v: s.dev.set(L.D_LAYER_THICKNESS, v))


def ocn_init_shadows(Prog: PrognosticVars, Diag=None, Tend=None) -> ShadowPrognosticVars:
    """ocn_init_shadows(Prog, Diag, Tend; backend) (init.jl:32-40)."""
    return ShadowPrognosticVars(Prog)


def autodiff_reverse_run_loop(timestep: float, Prog: PrognosticVars, d_Prog: ShadowPrognosticVars, Diag, Tend, Setup,
                              stepper, nsteps: int, seed: str | None = "ssh2") -> float:
    """`autodiff(Enzyme.Reverse, ocn_run_loop, Duplicated(sumCPU, ..), .., Duplicated(Prog, d_Prog), ..)`
    (test_Enzyme_end2end.jl:78-96): runs `nsteps` steps of `stepper` recording the trajectory, then the reverse
    sweep.  With `seed="ssh2"` the objective is the run loop's sum of squared SSH (run_loop.jl:24-44) and its
    value is returned; with `seed=None` whatever the caller stored in `d_Prog` is the adjoint of the final state.
    On return d_Prog holds dJ/d(initial normalVelocity, layerThickness) -- and, for ForwardEuler (the stepper the
    reference differentiates), d_Prog.ssh = dJ/d(initial ssh), an input of its own there."""
    if stepper not in (RungeKutta4, ForwardEuler):
        raise MokaError("autodiff_reverse_run_loop: unknown stepper")
    h = Prog.dev.handle
    L.check(L.lib().mokab_tape_begin(h, int(nsteps)))
    if stepper is ForwardEuler:
        L.check(L.lib().mokab_timestep_forward_euler(h, float(timestep), int(nsteps)))
        J = float("nan")
        if seed is not None:
            if seed != "ssh2":
                raise MokaError("autodiff_reverse_run_loop: unknown seed")
            J = float(np.sum(np.asarray(Prog.ssh, np.float64) ** 2)) if nsteps == 0 else reduce_sum(Prog, "ssh2")
            L.check(L.lib().mokab_adjoint_seed(h, L.SUM_SSH2))
        L.check(L.lib().mokab_adjoint_forward_euler(h))
        return J
    L.check(L.lib().mokab_timestep_rk4(h, float(timestep), int(nsteps), L.RK4_FUSED))
    J = float("nan")
    if seed is not None:
        if seed != "ssh2":
            raise MokaError("autodiff_reverse_run_loop: unknown seed")
        J = reduce_sum(Prog, "ssh2")
        L.check(L.lib().mokab_adjoint_seed(h, L.SUM_SSH2))
    L.check(L.lib().mokab_adjoint_rk4(h))
    return J


def reference_dt(mesh: Mesh) -> float:
    """ocn_init_alarms dt rule (init.jl:118): floor(2*(mean(dc)/1e3)*mean(dc)/200e3) seconds."""
    d = mesh.dcEdge_mean
    return float(np.floor(2 * (d / 1e3) * d / 200e3))


def cfl_dt(dc: float, depth: float = 1000.0, cfl: float = 0.5) -> float:
    """dt = cfl*dc/sqrt(g*H) for meshes where the reference rule floors to 0 s."""
    return cfl * dc / float(np.sqrt(GRAVITY * depth))


# ---- src/inertialGravityWave.jl ----------------------------------------------------------------------------------
class inertialGravityWave:
    """Exact inertia-gravity wave (inertialGravityWave.jl:6-64); lx follows the mesh's x period."""

    def __init__(self, fields: dict):
        self.g, self.f0, self.npx, self.npy, self.eta0, self.bottom_depth = GRAVITY, 1e-4, 2.0, 2.0, 1.0, 1000.0
        self.lx = fields["x_period"] / 1e3
        self.ly = np.sqrt(3.0) / 2.0 * self.lx
        self.kx = self.npx * 2.0 * np.pi / (self.lx * 1e3)
        self.ky = self.npy * 2.0 * np.pi / (self.ly * 1e3)
        self.omega = np.sqrt(self.f0 ** 2 + self.g * self.bottom_depth * (self.kx ** 2 + self.ky ** 2))
        self.f = fields

    def exact_ssh(self, t: float):
        f = self.f
        return self.eta0 * np.cos(self.kx * f["xCell"] + self.ky * f["yCell"] - self.omega * t)

    def exact_norm_vel(self, t: float):
        f = self.f
        ph = self.kx * f["xEdge"] + self.ky * f["yEdge"] - self.omega * t
        c = self.g / (self.omega ** 2.0 - self.f0 ** 2.0)
        u = self.eta0 * (c * (self.omega * self.kx * np.cos(ph) - self.f0 * self.ky * np.sin(ph)))
        v = self.eta0 * (c * (self.omega * self.ky * np.cos(ph) + self.f0 * self.kx * np.sin(ph)))
        return u * np.cos(f["angleEdge"]) + v * np.sin(f["angleEdge"])

    def initial_state(self):
        ssh = self.exact_ssh(0.0)
        return ssh, self.exact_norm_vel(0.0), self.bottom_depth + ssh


class kelvinWave:
    """Coastal Kelvin wave on a `channel_hex` mesh (project-defined; the reference
    rejects non-periodic meshes, VertMesh.jl:50-52): coast at x = 0, c = sqrt(g H), R = c / f0,
    eta = eta0 * exp(-x / R) * cos(ky * (y + c t)), u = 0, v = -(g / c) * eta; normal velocity is zero
    on solid-wall edges (`boundaryEdge`)."""

    def __init__(self, fields: dict, mode: int = 2):
        self.g, self.f0, self.eta0, self.bottom_depth = GRAVITY, 1e-4, 1.0, 1000.0
        self.c = np.sqrt(self.g * self.bottom_depth)
        self.R = self.c / self.f0
        self.ky = mode * 2.0 * np.pi / fields["y_period"]
        self.f = fields

    def exact_ssh(self, t: float, x=None, y=None):
        x = self.f["xCell"] if x is None else x
        y = self.f["yCell"] if y is None else y
        return self.eta0 * np.exp(-x / self.R) * np.cos(self.ky * (y + self.c * t))

    def exact_norm_vel(self, t: float):
        f = self.f
        v = -(self.g / self.c) * self.exact_ssh(t, f["xEdge"], f["yEdge"])
        un = v * np.sin(f["angleEdge"])
        return np.where(f["boundaryEdge"] != 0, 0.0, un)

    def initial_state(self):
        ssh = self.exact_ssh(0.0)
        return ssh, self.exact_norm_vel(0.0), self.bottom_depth + ssh
