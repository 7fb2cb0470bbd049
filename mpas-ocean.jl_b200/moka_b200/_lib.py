"""ctypes binding of libmoka_b200.so (include/moka_b200.h).  No CPU fallback: a missing library
or a missing CUDA device raises."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# MOKAB_LIB selects a tuning variant built next to the default one (Makefile: libmoka_b200_mb8.so); a file name, never a path
LIB_PATH = os.path.join(os.path.dirname(_HERE), os.path.basename(os.environ.get("MOKAB_LIB", "libmoka_b200.so")))

F64, F32 = 0, 1
(SSH, NORMAL_VELOCITY, LAYER_THICKNESS, SSH_PREV, NORMAL_VELOCITY_PREV, LAYER_THICKNESS_PREV,
 LAYER_THICKNESS_EDGE, THICKNESS_FLUX, VELOCITY_DIV_CELL, RELATIVE_VORTICITY,
 TEND_NORMAL_VELOCITY, TEND_LAYER_THICKNESS, D_SSH, D_NORMAL_VELOCITY, D_LAYER_THICKNESS) = range(15)
SUM_SSH2, SUM_MASS, SUM_ENERGY = range(3)
CELLS, EDGES, VERTICES = range(3)
RK4_FUSED, RK4_UNFUSED = 0, 1
PART_ALL, PART_INTERIOR, PART_BOUNDARY, PART_BOUNDARY_PUSH, PART_ALL_PUSH = 0, 1, 2, 3, 4
MESH_RENUMBER, MESH_EXPLICIT_EOE, MESH_KEEP_WIDTHS, MESH_EDGES_BY_CELL = 1, 2, 4, 8
HALO_NCCL, HALO_P2P, HALO_P2P_FUSED, HALO_P2P_LL = 0, 1, 2, 3
DECOMP_NO_OVERLAP, DECOMP_NO_GRAPH = 1, 2
COMM_ID_BYTES = 128

_I32P, _F64P = C.POINTER(C.c_int32), C.POINTER(C.c_double)

# name -> pointer type, in the order of struct mokab_mesh_desc
_DESC_PTRS = [
    ("cellsOnEdge", _I32P), ("verticesOnEdge", _I32P), ("edgesOnEdge", _I32P), ("nEdgesOnEdge", _I32P),
    ("weightsOnEdge", _F64P), ("dcEdge", _F64P), ("dvEdge", _F64P), ("fEdge", _F64P),
    ("xEdge", _F64P), ("yEdge", _F64P), ("zEdge", _F64P),
    ("edgesOnCell", _I32P), ("nEdgesOnCell", _I32P), ("edgeSignOnCell", _I32P), ("areaCell", _F64P),
    ("xCell", _F64P), ("yCell", _F64P), ("zCell", _F64P),
    ("edgesOnVertex", _I32P), ("edgeSignOnVertex", _I32P), ("areaTriangle", _F64P),
    ("restingThicknessSum", _F64P), ("boundaryEdge", _I32P),
]


class MeshDesc(C.Structure):
    _fields_ = ([(n, C.c_int64) for n in ("nCells", "nEdges", "nVertices", "maxEdges", "maxEdges2", "vertexDegree")] + _DESC_PTRS
                + [("nCellsOwned", C.c_int64), ("nEdgesOwned", C.c_int64)])


class MokaError(RuntimeError):
    """What the Julia shim raises with `error(msg)` (reference convention src/Architectures.jl:23)."""


_lib = None

# every symbol include/moka_b200.h declares
SYMBOLS = [
    "mokab_init", "mokab_finalize", "mokab_synchronize", "mokab_set_stream", "mokab_timer_start", "mokab_timer_stop",
    "mokab_launch_count", "mokab_host_alloc", "mokab_host_free", "mokab_last_error", "mokab_version",
    "mokab_mesh_create", "mokab_mesh_destroy", "mokab_mesh_get_perm", "mokab_mesh_device_bytes",
    "mokab_state_create", "mokab_state_create_levels", "mokab_state_levels", "mokab_state_destroy", "mokab_state_set", "mokab_state_get",
    "mokab_state_set_async", "mokab_state_get_async", "mokab_state_synchronize",
    "mokab_diagnostic_compute", "mokab_diagnostic_compute_consistent", "mokab_compute_normal_velocity_tendency", "mokab_compute_layer_thickness_tendency",
    "mokab_gradient_on_edge", "mokab_divergence_on_cell", "mokab_curl_on_vertex", "mokab_interpolate_cell2edge",
    "mokab_gradient_on_edge_vjp", "mokab_divergence_on_cell_vjp",
    "mokab_timestep_forward_euler", "mokab_timestep_forward_euler_unfused", "mokab_timestep_rk4", "mokab_reduce",
    "mokab_tape_begin", "mokab_tape_length", "mokab_adjoint_seed", "mokab_adjoint_rk4", "mokab_adjoint_forward_euler",
    "mokab_halo_setup", "mokab_halo_pack", "mokab_halo_unpack", "mokab_rk4_stage", "mokab_rk4_finish_step",
    "mokab_forward_euler_stage", "mokab_forward_euler_finish_step",
    "mokab_refresh_ssh", "mokab_mesh_block_counts", "mokab_mesh_derived_blocks",
    "mokab_halo_recv_device_indices", "mokab_p2p_blob_size", "mokab_p2p_export", "mokab_p2p_setup", "mokab_halo_push",
    "mokab_halo_wait", "mokab_halo_wait_arrivals", "mokab_p2p_error", "mokab_p2p_close",
    "mokab_set_option", "mokab_get_option", "mokab_trace_begin", "mokab_trace_read",
    "mokab_comm_get_unique_id", "mokab_comm_init", "mokab_comm_destroy", "mokab_comm_rank", "mokab_comm_barrier",
    "mokab_comm_allreduce_f64", "mokab_comm_allgather_bytes", "mokab_decomp_setup", "mokab_decomp_set_flags",
    "mokab_timestep_rk4_decomposed", "mokab_timestep_forward_euler_decomposed", "mokab_reduce_decomposed",
    "mokab_decomp_synchronize", "mokab_decomp_close",
]


def bind(L):
    """Declare the argument / result types of every entry point on the loaded library `L` and make it THE library of
    this process.  `lib()` calls this with libmoka_b200.so; the simulation tests (tests/sim) call it with their host
    build of the same sources."""
    global _lib
    L.mokab_last_error.restype = C.c_char_p
    vp, i64, dbl = C.c_void_p, C.c_int64, C.c_double
    sig = {
        "mokab_init": [C.c_int, C.POINTER(vp)], "mokab_finalize": [vp], "mokab_synchronize": [vp],
        "mokab_set_stream": [vp, vp], "mokab_timer_start": [vp], "mokab_timer_stop": [vp, C.POINTER(dbl)],
        "mokab_launch_count": [vp, C.POINTER(i64)], "mokab_host_alloc": [C.POINTER(vp), i64], "mokab_host_free": [vp],
        "mokab_mesh_create": [vp, C.POINTER(MeshDesc), C.c_uint32, C.POINTER(vp)], "mokab_mesh_destroy": [vp],
        "mokab_mesh_get_perm": [vp, C.c_int, _I32P], "mokab_mesh_device_bytes": [vp, C.POINTER(i64)],
        "mokab_state_create": [vp, vp, C.c_int, C.POINTER(vp)], "mokab_state_destroy": [vp],
        "mokab_state_create_levels": [vp, vp, C.c_int, C.c_int, C.POINTER(vp)], "mokab_state_levels": [vp, C.POINTER(C.c_int)],
        "mokab_state_set": [vp, C.c_int, vp], "mokab_state_get": [vp, C.c_int, vp],
        "mokab_state_set_async": [vp, C.c_int, vp], "mokab_state_get_async": [vp, C.c_int, vp],
        "mokab_state_synchronize": [vp],
        "mokab_diagnostic_compute": [vp], "mokab_diagnostic_compute_consistent": [vp], "mokab_compute_normal_velocity_tendency": [vp],
        "mokab_compute_layer_thickness_tendency": [vp],
        "mokab_gradient_on_edge": [vp, vp, _F64P, _F64P], "mokab_divergence_on_cell": [vp, vp, _F64P, _F64P],
        "mokab_gradient_on_edge_vjp": [vp, vp, _F64P, _F64P], "mokab_divergence_on_cell_vjp": [vp, vp, _F64P, _F64P],
        "mokab_curl_on_vertex": [vp, vp, _F64P, _F64P], "mokab_interpolate_cell2edge": [vp, vp, _F64P, _F64P],
        "mokab_timestep_forward_euler": [vp, dbl, i64], "mokab_timestep_forward_euler_unfused": [vp, dbl, i64], "mokab_timestep_rk4": [vp, dbl, i64, C.c_int],
        "mokab_reduce": [vp, C.c_int, C.POINTER(dbl)],
        "mokab_tape_begin": [vp, i64], "mokab_tape_length": [vp, C.POINTER(i64)],
        "mokab_adjoint_seed": [vp, C.c_int], "mokab_adjoint_rk4": [vp], "mokab_adjoint_forward_euler": [vp],
        "mokab_halo_setup": [vp, i64, _I32P, i64, _I32P], "mokab_halo_pack": [vp, C.c_int, vp, vp],
        "mokab_halo_unpack": [vp, C.c_int, vp, vp], "mokab_rk4_stage": [vp, dbl, C.c_int, C.c_int, vp],
        "mokab_rk4_finish_step": [vp], "mokab_refresh_ssh": [vp, vp],
        "mokab_forward_euler_stage": [vp, C.c_double, C.c_int, vp], "mokab_forward_euler_finish_step": [vp],
        "mokab_mesh_block_counts": [vp, C.POINTER(i64), C.POINTER(i64)],
        "mokab_mesh_derived_blocks": [vp, C.POINTER(i64), C.POINTER(i64)],
        "mokab_halo_recv_device_indices": [vp, _I32P], "mokab_p2p_blob_size": [C.POINTER(i64)],
        "mokab_p2p_export": [vp, C.c_int, vp],
        "mokab_p2p_setup": [vp, C.c_int, C.c_int, vp, C.c_int, _I32P, C.POINTER(i64), _I32P, C.c_int, _I32P],
        "mokab_comm_get_unique_id": [vp], "mokab_comm_init": [vp, vp, C.c_int, C.c_int, C.POINTER(vp)], "mokab_comm_destroy": [vp],
        "mokab_comm_rank": [vp, C.POINTER(C.c_int), C.POINTER(C.c_int)], "mokab_comm_barrier": [vp],
        "mokab_comm_allreduce_f64": [vp, _F64P, i64, C.c_int], "mokab_comm_allgather_bytes": [vp, vp, i64, vp],
        "mokab_decomp_setup": [vp, vp, C.POINTER(i64), C.POINTER(i64), C.c_int, C.c_uint32], "mokab_decomp_set_flags": [vp, C.c_uint32],
        "mokab_timestep_rk4_decomposed": [vp, dbl, i64], "mokab_timestep_forward_euler_decomposed": [vp, dbl, i64],
        "mokab_reduce_decomposed": [vp, C.c_int, C.POINTER(dbl)], "mokab_decomp_synchronize": [vp], "mokab_decomp_close": [vp],
        "mokab_set_option": [C.c_char_p, i64], "mokab_get_option": [C.c_char_p, C.POINTER(i64)],
        "mokab_trace_begin": [vp, i64], "mokab_trace_read": [vp, vp, i64, C.POINTER(i64)],
        "mokab_halo_push": [vp, C.c_int, vp], "mokab_halo_wait": [vp, vp], "mokab_halo_wait_arrivals": [vp, vp], "mokab_p2p_error": [vp, C.POINTER(C.c_int)], "mokab_p2p_close": [vp],
    }
    for name, args in sig.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = C.c_int
    _lib = L
    return L


def lib():
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MokaError(f"{LIB_PATH} is missing: build it with `make -C {os.path.dirname(LIB_PATH)}` "
                            "(python __graft_entry__.py build); there is no CPU fallback")
        bind(C.CDLL(LIB_PATH))
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise MokaError(lib().mokab_last_error().decode("utf-8", "replace"))


def set_option(name: str, value: int) -> None:
    """mokab_set_option: tuning switches of the fused stage kernel (include/moka_b200.h)."""
    check(lib().mokab_set_option(name.encode(), int(value)))


def get_option(name: str) -> int:
    v = C.c_int64()
    check(lib().mokab_get_option(name.encode(), C.byref(v)))
    return v.value


def fptr(a: np.ndarray):
    return a.ctypes.data_as(_F64P)
