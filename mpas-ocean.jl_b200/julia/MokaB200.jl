# MokaB200.jl -- the `B200` architecture for MOKA (jlk9/MPAS-Ocean.jl) as a DROP-IN: the reference's own entry points,
# with the reference's own signatures, on the reference's own structs -- selected by ordinary dispatch -- forwarding to
# libmoka_b200.so with `ccall`.  The only edit to the reference is the backend selector:
#
#     src/driver/mpas_ocean.jl:28      backend = CUDABackend()      ->      backend = MokaB200.B200()
#
# (plus `include("B200.jl")` after the last include of src/MOKA.jl:47, see INTEGRATION.md).
#
# How.  `B200 <: KA.GPU` is a KernelAbstractions backend VALUE, so everything the reference does with `backend` keeps
# working unchanged: `Adapt.adapt(backend, array)` in ReadHorzMesh / VerticalMesh / PrognosticVars (HorzMesh.jl:334-355,
# VertMesh.jl:46-82, PrognosticVars.jl:59-106), `KA.zeros(backend, ...)` in DiagnosticVars / TendencyVars / the driver
# (DiagnosticVars.jl:75-99, TendencyVars.jl:51-67, mpas_ocean.jl:37), `KA.get_backend` in the constructors' checks
# (Architectures.jl:27-33).  Those calls produce `B200Array`s -- a host mirror plus, once bound, the id of the device field
# it stands for -- INSIDE the reference's `Mesh`, `PrognosticVars`, `DiagnosticVars`, `TendencyVars`.  Julia cannot dispatch
# on a keyword (`; backend = ...`), but it does dispatch on the array type parameter of those structs: the methods below
# take `Prog::PrognosticVars{F, <:B200Array}` and are therefore chosen over the reference's KernelAbstractions methods
# whenever the state lives on a B200.  No KernelAbstractions kernel is ever launched on this path, there is no CUDA.jl, no
# multi-backend switch and no CPU fallback: the first entry-point call creates the device mesh and state from the host
# mirrors (mokab_mesh_create, mokab_state_create), every later call is one `ccall`, and host mirrors are refreshed lazily
# when Julia code looks at them (getindex, Array(x), Adapt.adapt(KA.CPU(), x) in write_netcdf, OutPut.jl:117-124).
#
# STATUS: source only.  Julia is not installed in the build image (SURVEY.md Appendix A), so this file has never been
# executed; the same C ABI (include/moka_b200.h) is exercised from Python ctypes (moka_b200/api.py).

module MokaB200

import Adapt
import KernelAbstractions as KA
using Dates

using MOKA
using MOKA: Mesh, ModelSetup, GlobalConfig, PrognosticVars, DiagnosticVars, TendencyVars, ForwardEuler, RungeKutta4,
            isRinging, advance!, reset!, mycopyto!
import MOKA: ocn_timestep, ocn_run_loop, diagnostic_compute!
import MOKA.normalVelocity: computeNormalVelocityTendency!
import MOKA.layerThickness: computeLayerThicknessTendency!

export B200, B200Array

const libmoka = get(ENV, "LIBMOKA_B200", "libmoka_b200.so")

# ---- error convention: nonzero status -> error(msg) (src/Architectures.jl:23,31,39) ----------------------------------------
function check(rc::Cint)
    rc == 0 && return nothing
    error(unsafe_string(ccall((:mokab_last_error, libmoka), Cstring, ())))
end

# field ids / enums of include/moka_b200.h
const F64 = Cint(0)
const SSH, NORMAL_VELOCITY, LAYER_THICKNESS = Cint(0), Cint(1), Cint(2)
const SSH_PREV, NORMAL_VELOCITY_PREV, LAYER_THICKNESS_PREV = Cint(3), Cint(4), Cint(5)
const LAYER_THICKNESS_EDGE, THICKNESS_FLUX, VELOCITY_DIV_CELL, RELATIVE_VORTICITY = Cint(6), Cint(7), Cint(8), Cint(9)
const TEND_NORMAL_VELOCITY, TEND_LAYER_THICKNESS = Cint(10), Cint(11)
const D_SSH, D_NORMAL_VELOCITY, D_LAYER_THICKNESS = Cint(12), Cint(13), Cint(14)
const SUM_SSH2 = Cint(0)
# halo exchange of a domain-decomposed run (include/moka_b200.h): packed ncclSend/ncclRecv, direct peer stores (push / wait kernels),
# direct peer stores from inside the boundary launch, flag-in-data packets into the peers' receive areas
const HALO_NCCL, HALO_P2P, HALO_P2P_FUSED, HALO_P2P_LL = Cint(0), Cint(1), Cint(2), Cint(3)
const RK4_FUSED = Cint(0)
const MESH_RENUMBER = UInt32(1)

# ---- src/Architectures.jl: the architecture object ------------------------------------------------------------------------
"""
    B200(device = 0)

KernelAbstractions backend value for one B200 (CUDA device `device`); pass it wherever the reference takes `backend`.
"""
struct B200 <: KA.GPU
    device::Int
end
B200() = B200(0)

const CONTEXTS = Dict{Int, Ptr{Cvoid}}()            # one library context per device, created on first use
function context(b::B200)
    get!(CONTEXTS, b.device) do
        r = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:mokab_init, libmoka), Cint, (Cint, Ptr{Ptr{Cvoid}}), b.device, r))
        r[]
    end
end
KA.synchronize(b::B200) = check(ccall((:mokab_synchronize, libmoka), Cint, (Ptr{Cvoid},), context(b)))   # e.g. pressure_gradient.jl:39

# ---- the array type the reference's structs end up holding ---------------------------------------------------------------------
"""
A host mirror (`host`) that may be bound to one field of a device state.  `stale`: the device holds newer values (pulled
on first look); `dirty`: the mirror was written by Julia code since the last upload (pushed before the next device call).
"""
mutable struct B200Array{T, N} <: AbstractArray{T, N}
    host::Array{T, N}
    backend::B200
    binding::Any                      # ::Union{Nothing, Binding}
    field::Cint
    stale::Bool
    dirty::Bool
end
B200Array(a::Array{T, N}, b::B200) where {T, N} = B200Array{T, N}(a, b, nothing, Cint(-1), false, false)

Base.size(a::B200Array) = size(a.host)
Base.IndexStyle(::Type{<:B200Array}) = IndexLinear()
Base.getindex(a::B200Array, i::Int) = (pull!(a); @inbounds a.host[i])
Base.setindex!(a::B200Array, v, i::Int) = (pull!(a); a.dirty = true; @inbounds a.host[i] = v)
Base.fill!(a::B200Array, v) = (a.stale = false; a.dirty = true; fill!(a.host, v); a)
Base.Array(a::B200Array) = (pull!(a); copy(a.host))
Base.similar(a::B200Array, ::Type{T}, dims::Dims) where {T} = B200Array(Array{T}(undef, dims), a.backend)
Base.copyto!(dst::B200Array, src::AbstractArray) = (dst.stale = false; dst.dirty = true; copyto!(dst.host, src isa B200Array ? Array(src) : src); dst)
Base.copyto!(dst::Array, src::B200Array) = (pull!(src); copyto!(dst, src.host))
function Base.deepcopy_internal(a::B200Array, dict::IdDict)      # PrognosticVars deep-copies one array per time level (PrognosticVars.jl:49-53)
    haskey(dict, a) && return dict[a]
    pull!(a)
    dict[a] = B200Array(copy(a.host), a.backend)
end

KA.get_backend(a::B200Array) = a.backend                                        # Architectures.jl:27-33
Adapt.adapt_storage(b::B200, a::Array) = B200Array(copy(a), b)                  # on_architecture, Architectures.jl:12
Adapt.adapt_storage(::B200, a::B200Array) = a
Adapt.adapt_storage(::KA.CPU, a::B200Array) = Array(a)                          # write_netcdf, OutPut.jl:122-124
KA.allocate(b::B200, ::Type{T}, dims::Tuple) where {T} = B200Array(Array{T}(undef, dims), b)
KA.allocate(b::B200, ::Type{T}, dims::Int...) where {T} = KA.allocate(b, T, dims)
KA.zeros(b::B200, ::Type{T}, dims::Tuple) where {T} = fill!(KA.allocate(b, T, dims), zero(T))
KA.zeros(b::B200, ::Type{T}, dims::Int...) where {T} = KA.zeros(b, T, dims)
KA.ones(b::B200, ::Type{T}, dims::Tuple) where {T} = fill!(KA.allocate(b, T, dims), one(T))
KA.ones(b::B200, ::Type{T}, dims::Int...) where {T} = KA.ones(b, T, dims)

host(a::B200Array) = a.host
host(a::Array) = a

# ---- binding the reference's structs to a device mesh + state -----------------------------------------------------------------
mutable struct Binding
    backend::B200
    mesh::Ptr{Cvoid}
    state::Ptr{Cvoid}
    arrays::Vector{B200Array}         # every mirror bound to a field of `state`
    comm::Ptr{Cvoid}                  # mokab_comm of a domain-decomposed run (C_NULL: single device)
end
const BINDINGS = WeakKeyDict{Any, Binding}()       # PrognosticVars object (a mutable struct) -> its device state; weak: the
                                                   # finalizer of a Binding frees the device mesh and state once Prog is collected

# struct mokab_mesh_desc, field for field (include/moka_b200.h)
struct MeshDesc
    nCells::Int64; nEdges::Int64; nVertices::Int64; maxEdges::Int64; maxEdges2::Int64; vertexDegree::Int64
    cellsOnEdge::Ptr{Int32}; verticesOnEdge::Ptr{Int32}; edgesOnEdge::Ptr{Int32}; nEdgesOnEdge::Ptr{Int32}
    weightsOnEdge::Ptr{Float64}; dcEdge::Ptr{Float64}; dvEdge::Ptr{Float64}; fEdge::Ptr{Float64}
    xEdge::Ptr{Float64}; yEdge::Ptr{Float64}; zEdge::Ptr{Float64}
    edgesOnCell::Ptr{Int32}; nEdgesOnCell::Ptr{Int32}; edgeSignOnCell::Ptr{Int32}; areaCell::Ptr{Float64}
    xCell::Ptr{Float64}; yCell::Ptr{Float64}; zCell::Ptr{Float64}
    edgesOnVertex::Ptr{Int32}; edgeSignOnVertex::Ptr{Int32}; areaTriangle::Ptr{Float64}
    restingThicknessSum::Ptr{Float64}; boundaryEdge::Ptr{Int32}
    nCellsOwned::Int64; nEdgesOwned::Int64
end

i32(a) = convert(Array{Int32}, host(a))             # NCDatasets may hand back Int32 or Int64 index arrays
f64(a) = convert(Array{Float64}, host(a))

# mokab_mesh_create from the host mirrors inside the reference's Mesh (ReadHorzMesh + signIndexField! + VerticalMesh have
# already run on the host, HorzMesh.jl:292-355, VertMesh.jl:46-82; the arrays are exactly what the C ABI expects: column-major
# (slot, entity), 1-based, 0 = absent)
function create_mesh(b::B200, mesh::Mesh; nCellsOwned = 0, nEdgesOwned = 0)
    E, P, D = mesh.HorzMesh.Edges, mesh.HorzMesh.PrimaryCells, mesh.HorzMesh.DualCells
    mesh.VertMesh.nVertLevels == 1 || error("B200: only nVertLevels == 1 is supported (the reference computes level 1 only, VertMesh.jl:31-36)")
    coe, voe, eoe, neoe = i32(E.cellsOnEdge), i32(E.verticesOnEdge), i32(E.edgesOnEdge), i32(E.nEdgesOnEdge)
    woe, dc, dv, fe = f64(E.weightsOnEdge), f64(E.dcEdge), f64(E.dvEdge), f64(E.fᵉ)
    xe, ye, ze = f64(E.xᵉ), f64(E.yᵉ), f64(E.zᵉ)
    eoc, neoc, sgnc, area = i32(P.edgesOnCell), i32(P.nEdgesOnCell), i32(P.edgeSignOnCell), f64(P.areaCell)
    xc, yc, zc = f64(P.xᶜ), f64(P.yᶜ), f64(P.zᶜ)
    eov, sgnv, atri = i32(D.edgesOnVertex), i32(D.edgeSignOnVertex), f64(D.areaTriangle)
    H = vec(f64(mesh.VertMesh.restingThicknessSum))
    decomposed = nCellsOwned > 0
    r = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve coe voe eoe neoe woe dc dv fe xe ye ze eoc neoc sgnc area xc yc zc eov sgnv atri H begin
        d = MeshDesc(P.nCells, E.nEdges, decomposed ? 0 : D.nVertices, P.maxEdges, size(eoe, 1), D.vertexDegree,
                     pointer(coe), pointer(voe), pointer(eoe), pointer(neoe), pointer(woe), pointer(dc), pointer(dv), pointer(fe),
                     pointer(xe), pointer(ye), pointer(ze),
                     pointer(eoc), pointer(neoc), pointer(sgnc), pointer(area), pointer(xc), pointer(yc), pointer(zc),
                     decomposed ? Ptr{Int32}(C_NULL) : pointer(eov), decomposed ? Ptr{Int32}(C_NULL) : pointer(sgnv),
                     decomposed ? Ptr{Float64}(C_NULL) : pointer(atri),
                     pointer(H), Ptr{Int32}(C_NULL), nCellsOwned, nEdgesOwned)
        check(ccall((:mokab_mesh_create, libmoka), Cint, (Ptr{Cvoid}, Ref{MeshDesc}, UInt32, Ptr{Ptr{Cvoid}}),
                    context(b), Ref(d), MESH_RENUMBER, r))
    end
    r[]
end

function attach!(bnd::Binding, a::B200Array, field::Cint; upload::Bool)
    a.binding === bnd && return
    a.binding, a.field = bnd, field
    upload && check(ccall((:mokab_state_set, libmoka), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}), bnd.state, field, a.host))
    a.dirty, a.stale = false, !upload
    push!(bnd.arrays, a)
end

const B200Prog = PrognosticVars{F, FV1} where {F, FV1 <: B200Array{F, 1}}

# The device state behind `Prog` (created on first use from the host mirrors), with whatever of Diag / Tend is at hand attached.
function bind!(Prog::B200Prog, mesh::Mesh; Diag = nothing, Tend = nothing, nCellsOwned = 0, nEdgesOwned = 0)
    bnd = get(BINDINGS, Prog, nothing)
    if bnd === nothing
        length(Prog.ssh) == 2 || error("nTimeLevels must be 2")                                  # time_integration.jl:23
        b = KA.get_backend(Prog.ssh[end])
        m = create_mesh(b, mesh; nCellsOwned = nCellsOwned, nEdgesOwned = nEdgesOwned)
        r = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:mokab_state_create, libmoka), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ptr{Ptr{Cvoid}}), context(b), m, F64, r))
        bnd = Binding(b, m, r[], B200Array[], C_NULL)
        finalizer(bnd) do x
            ccall((:mokab_state_destroy, libmoka), Cint, (Ptr{Cvoid},), x.state)
            ccall((:mokab_mesh_destroy, libmoka), Cint, (Ptr{Cvoid},), x.mesh)
        end
        BINDINGS[Prog] = bnd
        # time level [end] first (mokab_state_set of an [end] field also fills [1], like the deepcopy of PrognosticVars.jl:49-53), then [1]
        attach!(bnd, Prog.ssh[end], SSH; upload = true)
        attach!(bnd, Prog.normalVelocity[end], NORMAL_VELOCITY; upload = true)
        attach!(bnd, Prog.layerThickness[end], LAYER_THICKNESS; upload = true)
        attach!(bnd, Prog.ssh[1], SSH_PREV; upload = true)
        attach!(bnd, Prog.normalVelocity[1], NORMAL_VELOCITY_PREV; upload = true)
        attach!(bnd, Prog.layerThickness[1], LAYER_THICKNESS_PREV; upload = true)
    end
    if Diag !== nothing
        attach!(bnd, Diag.layerThicknessEdge, LAYER_THICKNESS_EDGE; upload = true)
        attach!(bnd, Diag.thicknessFlux, THICKNESS_FLUX; upload = true)
        attach!(bnd, Diag.velocityDivCell, VELOCITY_DIV_CELL; upload = true)
        attach!(bnd, Diag.relativeVorticity, RELATIVE_VORTICITY; upload = true)
    end
    if Tend !== nothing
        attach!(bnd, Tend.tendNormalVelocity, TEND_NORMAL_VELOCITY; upload = true)
        attach!(bnd, Tend.tendLayerThickness, TEND_LAYER_THICKNESS; upload = true)
    end
    # what Julia code wrote into a mirror since the last call goes up before the device runs
    for a in bnd.arrays
        if a.dirty
            check(ccall((:mokab_state_set, libmoka), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}), bnd.state, a.field, a.host))
            a.dirty = false
        end
    end
    bnd
end

# after a device call every mirror is out of date; nothing is copied until somebody looks
mark_stale!(bnd::Binding) = foreach(a -> (a.stale = true), bnd.arrays)

function pull!(a::B200Array)
    (a.binding === nothing || !a.stale) && return
    check(ccall((:mokab_state_get, libmoka), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}), a.binding.state, a.field, a.host))
    a.stale = false
    nothing
end

seconds(timestep) = Float64(timestep isa AbstractArray ? timestep[1] : timestep)       # the driver passes a 1-element backend array (mpas_ocean.jl:37-38)

function step!(bnd::Binding, ::Type{ForwardEuler}, dt::Float64, n::Integer)
    if bnd.comm == C_NULL
        check(ccall((:mokab_timestep_forward_euler, libmoka), Cint, (Ptr{Cvoid}, Cdouble, Int64), bnd.state, dt, n))
    else
        check(ccall((:mokab_timestep_forward_euler_decomposed, libmoka), Cint, (Ptr{Cvoid}, Cdouble, Int64), bnd.state, dt, n))
    end
    mark_stale!(bnd)
end
function step!(bnd::Binding, ::Type{RungeKutta4}, dt::Float64, n::Integer)
    if bnd.comm == C_NULL
        check(ccall((:mokab_timestep_rk4, libmoka), Cint, (Ptr{Cvoid}, Cdouble, Int64, Cint), bnd.state, dt, n, RK4_FUSED))
    else
        check(ccall((:mokab_timestep_rk4_decomposed, libmoka), Cint, (Ptr{Cvoid}, Cdouble, Int64), bnd.state, dt, n))
    end
    mark_stale!(bnd)
end

# ---- src/forward/time_integration.jl ---------------------------------------------------------------------------------------
# ocn_timestep(timestep, Prog, Diag, Tend, S, ::Type{ForwardEuler}; backend)            time_integration.jl:150-156
function ocn_timestep(timestep, Prog::B200Prog, Diag::DiagnosticVars, Tend::TendencyVars, S::ModelSetup, ::Type{ForwardEuler};
                      backend = KA.get_backend(Prog.ssh[end]))
    step!(bind!(Prog, S.mesh; Diag = Diag, Tend = Tend), ForwardEuler, seconds(timestep), 1)
end
# ocn_timestep(Prog, Diag, Tend, S, ::Type{RungeKutta4}; backend)                       time_integration.jl:61-66
# (dead code in the reference; semantics in DESIGN.md section 3); dt is the clock's, as at :71-75
function ocn_timestep(Prog::B200Prog, Diag::DiagnosticVars, Tend::TendencyVars, S::ModelSetup, ::Type{RungeKutta4};
                      backend = KA.get_backend(Prog.ssh[end]))
    dt = convert(Float64, Dates.value(Second(S.timeManager.timeStep)))
    step!(bind!(Prog, S.mesh; Diag = Diag, Tend = Tend), RungeKutta4, dt, 1)
end
# ... and with the time step passed like ForwardEuler's, so that `ocn_run_loop(..., RungeKutta4, ...)` works
function ocn_timestep(timestep, Prog::B200Prog, Diag::DiagnosticVars, Tend::TendencyVars, S::ModelSetup, ::Type{RungeKutta4};
                      backend = KA.get_backend(Prog.ssh[end]))
    step!(bind!(Prog, S.mesh; Diag = Diag, Tend = Tend), RungeKutta4, seconds(timestep), 1)
end

# ---- src/forward/run_loop.jl -------------------------------------------------------------------------------------------------
# ocn_run_loop(timestep, Prog, Diag, Tend, Setup, Stepper, clock, simulationAlarm, outputAlarm; backend)    run_loop.jl:8-22
# The clock still advances on the host; the steps between two alarm events run as ONE device-resident call (captured graphs).
function run_steps!(timestep, Prog, Diag, Tend, Setup, Stepper, clock, simulationAlarm, outputAlarm)
    bnd = bind!(Prog, Setup.mesh; Diag = Diag, Tend = Tend)
    dt = seconds(timestep)
    while !isRinging(simulationAlarm)
        n = 0
        while true
            advance!(clock)
            n += 1
            (isRinging(simulationAlarm) || isRinging(outputAlarm)) && break
        end
        step!(bnd, Stepper, dt, n)
        isRinging(outputAlarm) && reset!(outputAlarm)           # (the reference does no I/O here either, run_loop.jl:16-19)
    end
    bnd
end
function ocn_run_loop(timestep, Prog::B200Prog, Diag, Tend, Setup, Stepper, clock, simulationAlarm, outputAlarm;
                      backend = KA.get_backend(Prog.ssh[end]))
    run_steps!(timestep, Prog, Diag, Tend, Setup, Stepper, clock, simulationAlarm, outputAlarm)
    return nothing
end
# ocn_run_loop(sumCPU, sumGPU, timestep, ...): the same + sumArray of ssh^2 (run_loop.jl:26-51; the serial one-thread kernel of
# the reference becomes a two-level warp-shuffle reduction, mokab_reduce)
function ocn_run_loop(sumCPU, sumGPU, timestep, Prog::B200Prog, Diag, Tend, Setup, Stepper, clock, simulationAlarm, outputAlarm;
                      backend = KA.get_backend(Prog.ssh[end]))
    bnd = run_steps!(timestep, Prog, Diag, Tend, Setup, Stepper, clock, simulationAlarm, outputAlarm)
    r = Ref{Cdouble}(0.0)
    if bnd.comm == C_NULL
        check(ccall((:mokab_reduce, libmoka), Cint, (Ptr{Cvoid}, Cint, Ptr{Cdouble}), bnd.state, SUM_SSH2, r))
    else
        check(ccall((:mokab_reduce_decomposed, libmoka), Cint, (Ptr{Cvoid}, Cint, Ptr{Cdouble}), bnd.state, SUM_SSH2, r))
    end
    sumGPU[1] = sumGPU[1] + r[]                                  # run_loop.jl:49
    mycopyto!(sumCPU, sumGPU)
    return sumCPU[1]
end
MOKA.mycopyto!(dest::Array, src::B200Array) = copyto!(dest, src)

# ---- src/ocn entry points -------------------------------------------------------------------------------------------------------
# diagnostic_compute!(Mesh, Diag, Prog; backend)                                         DiagnosticVars.jl:108-117
function diagnostic_compute!(mesh::Mesh, Diag::DiagnosticVars, Prog::B200Prog; backend = KA.get_backend(Prog.ssh[end]))
    bnd = bind!(Prog, mesh; Diag = Diag)
    check(ccall((:mokab_diagnostic_compute, libmoka), Cint, (Ptr{Cvoid},), bnd.state))
    mark_stale!(bnd)
end
# computeNormalVelocityTendency!(Tend, Prog, Diag, Mesh, Config; backend)                normalVelocity.jl:21-53
function computeNormalVelocityTendency!(Tend::TendencyVars, Prog::B200Prog, Diag::DiagnosticVars, mesh::Mesh, Config::GlobalConfig;
                                        backend = KA.get_backend(Prog.ssh[end]))
    bnd = bind!(Prog, mesh; Diag = Diag, Tend = Tend)
    check(ccall((:mokab_compute_normal_velocity_tendency, libmoka), Cint, (Ptr{Cvoid},), bnd.state))
    mark_stale!(bnd)
end
# computeLayerThicknessTendency!(Tend, Prog, Diag, Mesh, Config; backend)                layerThickness.jl:14-28
function computeLayerThicknessTendency!(Tend::TendencyVars, Prog::B200Prog, Diag::DiagnosticVars, mesh::Mesh, Config::GlobalConfig;
                                        backend = KA.get_backend(Prog.ssh[end]))
    bnd = bind!(Prog, mesh; Diag = Diag, Tend = Tend)
    check(ccall((:mokab_compute_layer_thickness_tendency, libmoka), Cint, (Ptr{Cvoid},), bnd.state))
    mark_stale!(bnd)
end

# ---- reverse mode: what the reference asks of Enzyme ---------------------------------------------------------------------------
# `autodiff(Enzyme.Reverse, ocn_run_loop, Duplicated(sumCPU, ..), .., Duplicated(Prog, d_Prog), ..)` (test_Enzyme_end2end.jl:78-96;
# NaN on CUDA in the reference, :182-186) -> hand-written discrete adjoint in the library.  Runs the loop recording the
# trajectory, seeds d_ssh = 2 ssh (J = sum ssh^2) and sweeps back; fills d_Prog's [end] arrays, returns J.
function autodiff_reverse_run_loop!(d_Prog::PrognosticVars, timestep, Prog::B200Prog, Diag, Tend, Setup, Stepper, nsteps::Integer)
    bnd = bind!(Prog, Setup.mesh; Diag = Diag, Tend = Tend)
    check(ccall((:mokab_tape_begin, libmoka), Cint, (Ptr{Cvoid}, Int64), bnd.state, nsteps))
    step!(bnd, Stepper, seconds(timestep), nsteps)
    r = Ref{Cdouble}(0.0)
    if bnd.comm == C_NULL
        check(ccall((:mokab_reduce, libmoka), Cint, (Ptr{Cvoid}, Cint, Ptr{Cdouble}), bnd.state, SUM_SSH2, r))
    else    # decomposed state (collective): the tape was recorded by mokab_timestep_*_decomposed, the sweep exchanges its own halo copies
        check(ccall((:mokab_reduce_decomposed, libmoka), Cint, (Ptr{Cvoid}, Cint, Ptr{Cdouble}), bnd.state, SUM_SSH2, r))
    end
    check(ccall((:mokab_adjoint_seed, libmoka), Cint, (Ptr{Cvoid}, Cint), bnd.state, SUM_SSH2))
    if Stepper === ForwardEuler
        check(ccall((:mokab_adjoint_forward_euler, libmoka), Cint, (Ptr{Cvoid},), bnd.state))
    else
        check(ccall((:mokab_adjoint_rk4, libmoka), Cint, (Ptr{Cvoid},), bnd.state))
    end
    for (a, f) in ((d_Prog.ssh[end], D_SSH), (d_Prog.normalVelocity[end], D_NORMAL_VELOCITY), (d_Prog.layerThickness[end], D_LAYER_THICKNESS))
        check(ccall((:mokab_state_get, libmoka), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}), bnd.state, f, host(a)))
    end
    r[]
end

# ---- domain-decomposed runs: one Julia process per GPU (include/moka_b200.h, "domain-decomposed stepping") ----------------------
# The reference has no multi-device driver.  Rank 0 draws the communicator id and hands it to the others -- here through a file
# on a shared path; with MPI.jl (already a dependency of the reference, Project.toml) `MPI.Bcast!(id, 0, comm)` does the same.
function comm_init(b::B200, rank::Integer, nranks::Integer, id_path::AbstractString)
    id = Vector{UInt8}(undef, 128)
    if rank == 0
        check(ccall((:mokab_comm_get_unique_id, libmoka), Cint, (Ptr{UInt8},), id))
        write(id_path * ".tmp", id); mv(id_path * ".tmp", id_path; force = true)
    else
        while !isfile(id_path); sleep(0.05); end
        id = read(id_path)
    end
    r = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:mokab_comm_init, libmoka), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Cint, Cint, Ptr{Ptr{Cvoid}}), context(b), id, rank, nranks, r))
    r[]
end

# `Prog` / `mesh` hold THIS rank's local mesh (owned entities first, then one halo layer: moka_b200/partition.py computes such
# meshes and their halo lists in the reference's array layouts; the reference itself has no partitioner); send_idx / recv_idx
# are the halo lists in the combined [cells | edges] index space, ordered by rank, send_counts / recv_counts how many entries
# go to / come from every rank.
function decompose!(Prog::B200Prog, mesh::Mesh, comm::Ptr{Cvoid}, nCellsOwned, nEdgesOwned,
                    send_idx::Vector{Int32}, recv_idx::Vector{Int32}, send_counts::Vector{Int64}, recv_counts::Vector{Int64};
                    Diag = nothing, Tend = nothing, halo_mode::Cint = HALO_P2P)
    haskey(BINDINGS, Prog) && error("decompose!: the state is already on the device")
    b = KA.get_backend(Prog.ssh[end])
    m = create_mesh(b, mesh; nCellsOwned = nCellsOwned, nEdgesOwned = nEdgesOwned)
    check(ccall((:mokab_halo_setup, libmoka), Cint, (Ptr{Cvoid}, Int64, Ptr{Int32}, Int64, Ptr{Int32}),
                m, length(send_idx), send_idx, length(recv_idx), recv_idx))
    r = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:mokab_state_create, libmoka), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ptr{Ptr{Cvoid}}), context(b), m, F64, r))
    bnd = Binding(b, m, r[], B200Array[], comm)
    BINDINGS[Prog] = bnd
    for (a, f) in ((Prog.ssh[end], SSH), (Prog.normalVelocity[end], NORMAL_VELOCITY), (Prog.layerThickness[end], LAYER_THICKNESS),
                   (Prog.ssh[1], SSH_PREV), (Prog.normalVelocity[1], NORMAL_VELOCITY_PREV), (Prog.layerThickness[1], LAYER_THICKNESS_PREV))
        attach!(bnd, a, f; upload = true)
    end
    check(ccall((:mokab_decomp_setup, libmoka), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Cint, UInt32),
                bnd.state, comm, send_counts, recv_counts, halo_mode, 0))                        # flags 0: overlap + captured graphs
    bind!(Prog, mesh; Diag = Diag, Tend = Tend)
end

end # module
