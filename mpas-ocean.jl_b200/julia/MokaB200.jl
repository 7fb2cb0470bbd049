# MokaB200.jl -- the `B200` architecture for MOKA (jlk9/MPAS-Ocean.jl): the reference's entry points for the
# forward hot path, specialised on `backend::B200`, forwarding to libmoka_b200.so with `ccall`.
#
# STATUS: source only.  Julia is not installed in the build image (SURVEY.md Appendix A), so this file
# has never been executed; the same C ABI (include/moka_b200.h) is exercised from Python ctypes
# (moka_b200/api.py), which mirrors this file one to one.  INTEGRATION.md shows where it is included.
#
# No KernelAbstractions / CUDA.jl dispatch happens on this path: `B200` is deliberately NOT a
# `KA.Backend`; every method below that the reference defines with `backend = ...` keyword gets a
# positional-`B200` sibling.

module MokaB200

using MOKA
using MOKA: Mesh, HorzMesh, VerticalMesh, GlobalConfig, ModelSetup, ForwardEuler, RungeKutta4
import MOKA: diagnostic_compute!, computeNormalVelocityTendency!, computeLayerThicknessTendency!,
             ocn_timestep, ocn_run_loop

const libmoka = get(ENV, "LIBMOKA_B200", "libmoka_b200.so")

# ---- error convention: nonzero status -> error(msg) (src/Architectures.jl:23,31,39) ----------------------
function check(rc::Cint)
    rc == 0 && return nothing
    error(unsafe_string(ccall((:mokab_last_error, libmoka), Cstring, ())))
end

# ---- src/Architectures.jl: the architecture object ---------------------------------------------------------
mutable struct B200
    ctx::Ptr{Cvoid}
    function B200(device::Integer = 0)
        r = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:mokab_init, libmoka), Cint, (Cint, Ptr{Ptr{Cvoid}}), device, r))
        b = new(r[])
        finalizer(x -> ccall((:mokab_finalize, libmoka), Cint, (Ptr{Cvoid},), x.ctx), b)
    end
end
synchronize(b::B200) = check(ccall((:mokab_synchronize, libmoka), Cint, (Ptr{Cvoid},), b.ctx))

# field ids / enums of include/moka_b200.h
const F64, F32 = Cint(0), Cint(1)
const SSH, NORMAL_VELOCITY, LAYER_THICKNESS = Cint(0), Cint(1), Cint(2)
const LAYER_THICKNESS_EDGE, THICKNESS_FLUX, VELOCITY_DIV_CELL, RELATIVE_VORTICITY = Cint(6), Cint(7), Cint(8), Cint(9)
const TEND_NORMAL_VELOCITY, TEND_LAYER_THICKNESS = Cint(10), Cint(11)
const D_SSH, D_NORMAL_VELOCITY, D_LAYER_THICKNESS = Cint(12), Cint(13), Cint(14)      # shadow state d_Prog
const SUM_SSH2 = Cint(0)
const RK4_FUSED, RK4_UNFUSED = Cint(0), Cint(1)

# struct mokab_mesh_desc, field for field
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

# ---- Adapt.adapt_structure(backend, mesh): upload a host Mesh (HorzMesh.jl:334-355, VertMesh.jl:46-82) ---------
mutable struct B200Mesh
    handle::Ptr{Cvoid}
    backend::B200
    host::Mesh                      # the KA.CPU() mesh it was built from (dims, for array sizes)
end

function on_architecture(b::B200, mesh::Mesh)
    E, P, D = mesh.HorzMesh.Edges, mesh.HorzMesh.PrimaryCells, mesh.HorzMesh.DualCells
    H = vec(mesh.VertMesh.restingThicknessSum)
    GC.@preserve E P D H begin
        d = MeshDesc(P.nCells, E.nEdges, D.nVertices, P.maxEdges, size(E.edgesOnEdge, 1), D.vertexDegree,
                     pointer(E.cellsOnEdge), pointer(E.verticesOnEdge), pointer(E.edgesOnEdge), pointer(E.nEdgesOnEdge),
                     pointer(E.weightsOnEdge), pointer(E.dcEdge), pointer(E.dvEdge), pointer(E.fᵉ),
                     pointer(E.xᵉ), pointer(E.yᵉ), pointer(E.zᵉ),
                     pointer(P.edgesOnCell), pointer(P.nEdgesOnCell), pointer(P.edgeSignOnCell), pointer(P.areaCell),
                     pointer(P.xᶜ), pointer(P.yᶜ), pointer(P.zᶜ),
                     pointer(D.edgesOnVertex), pointer(D.edgeSignOnVertex), pointer(D.areaTriangle),
                     pointer(H), C_NULL, 0, 0)
        r = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:mokab_mesh_create, libmoka), Cint, (Ptr{Cvoid}, Ref{MeshDesc}, UInt32, Ptr{Ptr{Cvoid}}), b.ctx, Ref(d), 1, r))
    end
    m = B200Mesh(r[], b, mesh)
    finalizer(x -> ccall((:mokab_mesh_destroy, libmoka), Cint, (Ptr{Cvoid},), x.handle), m)
end

# ---- PrognosticVars / DiagnosticVars / TendencyVars on the backend: one device state behind three views ------
mutable struct B200State
    handle::Ptr{Cvoid}
    mesh::B200Mesh
end

function B200State(mesh::B200Mesh, Prog::MOKA.PrognosticVars)     # Prog on KA.CPU(), PrognosticVars.jl:59-106
    r = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:mokab_state_create, libmoka), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ptr{Ptr{Cvoid}}), mesh.backend.ctx, mesh.handle, F64, r))
    s = B200State(r[], mesh)
    finalizer(x -> ccall((:mokab_state_destroy, libmoka), Cint, (Ptr{Cvoid},), x.handle), s)
    set!(s, SSH, Prog.ssh[end]); set!(s, NORMAL_VELOCITY, vec(Prog.normalVelocity[end])); set!(s, LAYER_THICKNESS, vec(Prog.layerThickness[end]))
    s
end
set!(s::B200State, field::Cint, a::Array{Float64}) = check(ccall((:mokab_state_set, libmoka), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}), s.handle, field, a))
function get!(a::Array{Float64}, s::B200State, field::Cint)       # write_netcdf adapts back to the CPU: OutPut.jl:7-10
    check(ccall((:mokab_state_get, libmoka), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}), s.handle, field, a)); a
end

# Pipelined transfers: `a` must be page-locked (mokab_host_alloc) and stay valid until synchronize(s).
set_async!(s::B200State, field::Cint, a::Array{Float64}) = check(ccall((:mokab_state_set_async, libmoka), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}), s.handle, field, a))
get_async!(a::Array{Float64}, s::B200State, field::Cint) = check(ccall((:mokab_state_get_async, libmoka), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}), s.handle, field, a))
synchronize(s::B200State) = check(ccall((:mokab_state_synchronize, libmoka), Cint, (Ptr{Cvoid},), s.handle))

# ---- src/ocn entry points ---------------------------------------------------------------------------------------
diagnostic_compute!(s::B200State) = check(ccall((:mokab_diagnostic_compute, libmoka), Cint, (Ptr{Cvoid},), s.handle))                      # DiagnosticVars.jl:108
computeNormalVelocityTendency!(s::B200State) = check(ccall((:mokab_compute_normal_velocity_tendency, libmoka), Cint, (Ptr{Cvoid},), s.handle))  # normalVelocity.jl:21
computeLayerThicknessTendency!(s::B200State) = check(ccall((:mokab_compute_layer_thickness_tendency, libmoka), Cint, (Ptr{Cvoid},), s.handle))  # layerThickness.jl:14

# ---- src/forward entry points -----------------------------------------------------------------------------------------
# ocn_timestep(timestep, Prog, Diag, Tend, Setup, ::Type{ForwardEuler}; backend)  time_integration.jl:150
ocn_timestep(dt::Float64, s::B200State, ::Type{ForwardEuler}; nsteps = 1) =
    check(ccall((:mokab_timestep_forward_euler, libmoka), Cint, (Ptr{Cvoid}, Cdouble, Int64), s.handle, dt, nsteps))
# ocn_timestep(Prog, Diag, Tend, Setup, ::Type{RungeKutta4}; backend)             time_integration.jl:61
ocn_timestep(dt::Float64, s::B200State, ::Type{RungeKutta4}; nsteps = 1, fused = true) =
    check(ccall((:mokab_timestep_rk4, libmoka), Cint, (Ptr{Cvoid}, Cdouble, Int64, Cint), s.handle, dt, nsteps, fused ? RK4_FUSED : RK4_UNFUSED))

# ocn_run_loop(timestep, Prog, Diag, Tend, Setup, Stepper, clock, simulationAlarm, outputAlarm; backend)  run_loop.jl:8-22.
# The clock still advances on the host; the steps between two alarm events run as one device-resident call.
function ocn_run_loop(dt::Float64, s::B200State, Stepper, clock, simulationAlarm, outputAlarm)
    while !MOKA.isRinging(simulationAlarm)
        n = 0
        while !MOKA.isRinging(simulationAlarm) && !MOKA.isRinging(outputAlarm)
            MOKA.advance!(clock); n += 1
        end
        ocn_timestep(dt, s, Stepper; nsteps = n)
        MOKA.isRinging(outputAlarm) && MOKA.reset!(outputAlarm)
    end
    nothing
end

# sumArray replacement (run_loop.jl:47-51)
function sum_ssh2(s::B200State)
    r = Ref{Cdouble}(0.0)
    check(ccall((:mokab_reduce, libmoka), Cint, (Ptr{Cvoid}, Cint, Ptr{Cdouble}), s.handle, 0, r)); r[]
end

# ---- reverse mode: stands in for `autodiff(Enzyme.Reverse, ocn_run_loop, ..., Duplicated(Prog, d_Prog), ...)` --------------
# (test/enzyme/test_Enzyme_end2end.jl:78-96).  Runs `nsteps` steps of `Stepper` recording the trajectory, seeds
# d_ssh = 2 ssh (J = sum ssh^2, run_loop.jl:47-51) and sweeps back; returns J and fills the shadow arrays.
# ForwardEuler is the stepper the reference differentiates (its lagged thickness flux included); `d_ssh`, when given,
# receives d_Prog.ssh[end] (ssh is an input of its own for ForwardEuler; zero for RungeKutta4, where it is folded into h).
function autodiff_reverse_run_loop!(d_normalVelocity::Array{Float64}, d_layerThickness::Array{Float64},
                                    dt::Float64, s::B200State, nsteps::Integer;
                                    Stepper = RungeKutta4, d_ssh::Union{Nothing, Array{Float64}} = nothing)
    check(ccall((:mokab_tape_begin, libmoka), Cint, (Ptr{Cvoid}, Int64), s.handle, nsteps))
    ocn_timestep(dt, s, Stepper; nsteps = nsteps)
    J = sum_ssh2(s)
    check(ccall((:mokab_adjoint_seed, libmoka), Cint, (Ptr{Cvoid}, Cint), s.handle, SUM_SSH2))
    if Stepper === ForwardEuler
        check(ccall((:mokab_adjoint_forward_euler, libmoka), Cint, (Ptr{Cvoid},), s.handle))
    else
        check(ccall((:mokab_adjoint_rk4, libmoka), Cint, (Ptr{Cvoid},), s.handle))
    end
    get!(d_normalVelocity, s, D_NORMAL_VELOCITY); get!(d_layerThickness, s, D_LAYER_THICKNESS)
    d_ssh === nothing || get!(d_ssh, s, D_SSH)
    J
end
# adjoints of the two operators test/enzyme/test_Enzyme_Operators.jl differentiates
GradientOnEdge_vjp!(d_scalar::Array{Float64}, d_grad::Array{Float64}, m::B200Mesh) =
    check(ccall((:mokab_gradient_on_edge_vjp, libmoka), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), m.backend.ctx, m.handle, d_grad, d_scalar))
DivergenceOnCell_vjp!(d_vec::Array{Float64}, d_div::Array{Float64}, m::B200Mesh) =
    check(ccall((:mokab_divergence_on_cell_vjp, libmoka), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), m.backend.ctx, m.handle, d_div, d_vec))

export B200, B200Mesh, B200State, on_architecture, synchronize, sum_ssh2, autodiff_reverse_run_loop!
end # module
