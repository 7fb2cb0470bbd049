"""MPAS NetCDF mesh / initial-state input and model output, mirror of the reference's file interface.

  ReadHorzMesh(path; backend)          src/infra/MPASMesh/HorzMesh.jl:166-355  (dims + variables read)
  VerticalMesh(path, hmesh; backend)   src/infra/MPASMesh/VertMesh.jl:46-82    (needs global attr is_periodic == "YES")
  PrognosticVars(config, mesh)         src/ocn/PrognosticVars.jl:59-106        (ssh, normalVelocity, layerThickness at Time 1)
  write_netcdf(Setup, Diag, Prog)      src/infra/OutPut.jl:117-215             (output schema)

The reference reads NetCDF-4 through NCDatasets; this image has no netCDF4/HDF5 library,
so files are NetCDF-3 (classic / 64-bit offset) through scipy.io.netcdf_file -- same dimension and variable
names, same on-disk index order (a Julia array `(maxEdges, nCells)` is the NetCDF variable `(nCells, maxEdges)`).
`write_mesh_netcdf` writes such a file from a mesh dictionary (planar_hex.py), so every config of BASELINE.json can
be staged on disk and driven through `ocn_init` exactly as the reference's driver does.
"""
from __future__ import annotations

import numpy as np
from scipy.io import netcdf_file

from ._lib import MokaError

# variable -> (dims, dtype); dims in NetCDF (C) order
_MESH_VARS = {
    "xCell": (("nCells",), "f8"), "yCell": (("nCells",), "f8"), "zCell": (("nCells",), "f8"), "fCell": (("nCells",), "f8"),
    "areaCell": (("nCells",), "f8"), "nEdgesOnCell": (("nCells",), "i4"),
    "edgesOnCell": (("nCells", "maxEdges"), "i4"), "verticesOnCell": (("nCells", "maxEdges"), "i4"),
    "cellsOnCell": (("nCells", "maxEdges"), "i4"),
    "xVertex": (("nVertices",), "f8"), "yVertex": (("nVertices",), "f8"), "zVertex": (("nVertices",), "f8"),
    "fVertex": (("nVertices",), "f8"), "areaTriangle": (("nVertices",), "f8"),
    "edgesOnVertex": (("nVertices", "vertexDegree"), "i4"), "cellsOnVertex": (("nVertices", "vertexDegree"), "i4"),
    "xEdge": (("nEdges",), "f8"), "yEdge": (("nEdges",), "f8"), "zEdge": (("nEdges",), "f8"), "fEdge": (("nEdges",), "f8"),
    "nEdgesOnEdge": (("nEdges",), "i4"), "cellsOnEdge": (("nEdges", "TWO"), "i4"), "verticesOnEdge": (("nEdges", "TWO"), "i4"),
    "edgesOnEdge": (("nEdges", "maxEdges2"), "i4"), "weightsOnEdge": (("nEdges", "maxEdges2"), "f8"),
    "dvEdge": (("nEdges",), "f8"), "dcEdge": (("nEdges",), "f8"), "angleEdge": (("nEdges",), "f8"),
    "minLevelCell": (("nCells",), "i4"), "maxLevelCell": (("nCells",), "i4"),
    "boundaryEdge": (("nEdges",), "i4"),
}
_OPTIONAL = ("fCell", "fVertex", "fEdge", "boundaryEdge")            # HorzMesh.jl:176-181,225-230,257-262


def write_mesh_netcdf(path: str, fields: dict, state=None) -> None:
    """Write an MPAS mesh (+ optionally the initial state `(ssh, normalVelocity, layerThickness)`) in the layout the
    reference reads: restingThickness (Time, nCells, nVertLevels) -- `ds["restingThickness"][:,:,1]`, VertMesh.jl:57 --
    ssh (Time, nCells), normalVelocity (Time, nEdges, nVertLevels), layerThickness (Time, nCells, nVertLevels)
    (PrognosticVars.jl:95-99), global attribute is_periodic (VertMesh.jl:50)."""
    nVL = int(fields.get("nVertLevels", 1))
    with netcdf_file(path, "w", version=2) as ds:
        dims = {"nCells": fields["nCells"], "nEdges": fields["nEdges"], "nVertices": fields["nVertices"],
                "maxEdges": fields["maxEdges"], "maxEdges2": fields["maxEdges2"], "vertexDegree": fields["vertexDegree"],
                "TWO": 2, "nVertLevels": nVL, "Time": 1}
        for k, v in dims.items():
            ds.createDimension(k, int(v))
        ds.is_periodic = str(fields.get("is_periodic", "YES"))
        for k in ("x_period", "y_period"):
            if k in fields:
                setattr(ds, k, float(fields[k]))
        for name, (vd, dt) in _MESH_VARS.items():
            if name not in fields:
                if name in _OPTIONAL:
                    continue
                raise MokaError(f"write_mesh_netcdf: mesh field {name} is missing")
            v = ds.createVariable(name, dt, vd)
            v[:] = np.asarray(fields[name]).reshape([int(dims[d]) for d in vd])
        v = ds.createVariable("restingThickness", "f8", ("Time", "nCells", "nVertLevels"))
        v[:] = np.asarray(fields["restingThickness"], np.float64).reshape(1, int(dims["nCells"]), nVL)
        if state is not None:
            ssh, u, h = state
            ds.createVariable("ssh", "f8", ("Time", "nCells"))[:] = np.asarray(ssh, np.float64).reshape(1, -1)
            ds.createVariable("normalVelocity", "f8", ("Time", "nEdges", "nVertLevels"))[:] = np.asarray(u, np.float64).reshape(1, -1, nVL)
            ds.createVariable("layerThickness", "f8", ("Time", "nCells", "nVertLevels"))[:] = np.asarray(h, np.float64).reshape(1, -1, nVL)


def read_mesh_fields(path: str) -> dict:
    """Host part of ReadHorzMesh + VerticalMesh: dims (HorzMesh.jl:169-170,209-211,247), every variable of
    readPrimaryMesh / readDualMesh / readEdgeInfo, and the vertical-mesh variables (VertMesh.jl:54-57,73)."""
    with netcdf_file(path, "r", mmap=False) as ds:
        attr = getattr(ds, "is_periodic", None)
        if attr is None:
            raise MokaError("mesh file has no global attribute is_periodic (VertMesh.jl:50)")
        is_periodic = attr.decode() if isinstance(attr, bytes) else str(attr)
        f = {k: int(ds.dimensions[k]) for k in ("nCells", "nEdges", "nVertices", "maxEdges", "maxEdges2", "vertexDegree")}
        f["nVertLevels"] = int(ds.dimensions["nVertLevels"]) if "nVertLevels" in ds.dimensions else 1
        f["is_periodic"] = is_periodic
        for name, (vd, dt) in _MESH_VARS.items():
            if name in ds.variables:
                f[name] = np.array(ds.variables[name][:], dtype=np.int32 if dt == "i4" else np.float64)
            elif name in ("fCell", "fVertex", "fEdge"):                # coriolis defaults to zero (HorzMesh.jl:179,228,260)
                f[name] = np.zeros(f[{"fCell": "nCells", "fVertex": "nVertices", "fEdge": "nEdges"}[name]])
            elif name not in _OPTIONAL and name not in ("minLevelCell", "maxLevelCell"):
                raise MokaError(f"mesh file has no variable {name}")
        if "restingThickness" in ds.variables:
            rt = np.array(ds.variables["restingThickness"][:], np.float64)
            f["restingThickness"] = rt.reshape(-1, f["nCells"], f["nVertLevels"])[0]      # [:,:,1], VertMesh.jl:57
        for k in ("x_period", "y_period"):
            if hasattr(ds, k):
                f[k] = float(getattr(ds, k))
    return f


def ReadHorzMesh(meshPath: str, backend=None) -> dict:
    """ReadHorzMesh(meshPath; backend) (HorzMesh.jl:334-355).  Returns the host fields; the sign fields of
    signIndexField! (HorzMesh.jl:292-332) are derived by the library at upload (csrc/mesh.cuh), and the device
    copy (Adapt.adapt_structure) happens when `VerticalMesh` completes the mesh."""
    return read_mesh_fields(meshPath)


def VerticalMesh(mesh_fp, hmesh: dict | None = None, backend=None, nVertLevels: int = 1, renumber: bool = True):
    """VerticalMesh(mesh_fp, mesh; backend) (VertMesh.jl:46-82) or, with a fields dict as first argument, the unit-test
    constructor VerticalMesh(mesh; nVertLevels, backend) (VertMesh.jl:92-117: unit resting thickness).  Returns the
    complete `api.Mesh` (= Mesh(HorzMesh, VertMesh), MPASMesh.jl:19-29) on the backend."""
    from . import api
    if isinstance(mesh_fp, dict):                                      # VertMesh.jl:92-117
        f = dict(mesh_fp)
        f["nVertLevels"] = nVertLevels
        f["restingThickness"] = np.ones((f["nCells"], 1))
        f["restingThicknessSum"] = np.ones(f["nCells"])
        return api.Mesh(f, backend, renumber=renumber)
    f = dict(hmesh) if hmesh is not None else read_mesh_fields(mesh_fp)
    if str(f.get("is_periodic", "")).upper() != "YES" and "boundaryEdge" not in f:
        raise MokaError("Support for non-periodic meshes is not yet implemented")         # VertMesh.jl:50-52
    if "restingThickness" not in f:
        raise MokaError("mesh file has no variable restingThickness")
    if "maxLevelCell" in f and not np.all(f["maxLevelCell"] == f["nVertLevels"]):
        import warnings
        warnings.warn("Vertical Mesh is not stacked. Must implement vertical masking before this mesh can be used")  # :61-66
    return api.Mesh(f, backend, renumber=renumber)


def read_initial_state(input_filename: str, nCells: int, nEdges: int, nVertLevels: int = 1):
    """PrognosticVars(config, mesh) file part (PrognosticVars.jl:85-99): time index 1 of ssh, normalVelocity,
    layerThickness."""
    with netcdf_file(input_filename, "r", mmap=False) as ds:
        for k in ("ssh", "normalVelocity", "layerThickness"):
            if k not in ds.variables:
                raise MokaError(f"input file has no variable {k}")
        ssh = np.array(ds.variables["ssh"][:], np.float64).reshape(-1, nCells)[0]
        u = np.array(ds.variables["normalVelocity"][:], np.float64).reshape(-1, nEdges, nVertLevels)[0]
        h = np.array(ds.variables["layerThickness"][:], np.float64).reshape(-1, nCells, nVertLevels)[0]
    if nVertLevels != 1:
        raise MokaError("only nVertLevels == 1 is supported")
    return ssh, u[:, 0].copy(), h[:, 0].copy()


def write_netcdf(Setup, Diag, Prog, d_Prog=None, time_level: str = "reference") -> None:
    """write_netcdf(Setup, Diag, Prog[, d_Prog]) (OutPut.jl:1-115 with the shadow, :117-215 without).

    Same dimensions (time, nCells, nEdges, nVertices, nVertLevels, maxEdges, TWO), global attribute `dt`, the same
    variables defined and the same subset written (the connectivity variables the reference only defines --
    edgeSignOnCell, cellsOnEdge, verticesOnCell, verticesOnEdge, angleEdge, OutPut.jl:173-204 -- stay unwritten).
    `time_level="reference"` reproduces what the reference writes: `Adapt.adapt_structure(CPU(), Prog)` rebuilds Prog
    from time level 1 (PrognosticVars.jl:108-113), so the file holds the state one step BEFORE the last;
    `time_level="end"` writes Prog.*[end]."""
    from .config import ConfigGet
    from .time_manager import Period
    f, clock, cfg = Setup.mesh_fields, Setup.timeManager, Setup.config
    out = ConfigGet(ConfigGet(cfg.streams, "output"), "filename_template")
    prev = time_level == "reference"
    ssh = Prog.ssh_prev if prev else Prog.ssh
    h = Prog.layerThickness_prev if prev else Prog.layerThickness
    u = Prog.normalVelocity_prev if prev else Prog.normalVelocity
    nVL = int(f.get("nVertLevels", 1))
    with netcdf_file(out, "w", version=2) as ds:
        for k, v in (("time", 1), ("nCells", f["nCells"]), ("nEdges", f["nEdges"]), ("nVertices", f["nVertices"]),
                     ("nVertLevels", nVL), ("maxEdges", f["maxEdges"]), ("TWO", 2)):
            ds.createDimension(k, int(v))
        ts = clock.timeStep
        ds.dt = ts.seconds() if isinstance(ts, Period) else float(ts)
        ds.createVariable("time", "f8", ("time",))[:] = (clock.currTime - clock.startTime).total_seconds()
        for name, dim in (("xCell", "nCells"), ("yCell", "nCells"), ("xEdge", "nEdges"), ("yEdge", "nEdges"),
                          ("xVertex", "nVertices"), ("yVertex", "nVertices"), ("dcEdge", "nEdges"), ("areaCell", "nCells"),
                          ("areaTriangle", "nVertices")):
            ds.createVariable(name, "f8", (dim,))[:] = f[name]
        ds.createVariable("angleEdge", "f8", ("nEdges",))
        ds.createVariable("edgeSignOnCell", "i4", ("nCells", "maxEdges"))
        ds.createVariable("nEdgesOnCell", "i4", ("nCells",))[:] = f["nEdgesOnCell"]
        ds.createVariable("nEdgesOnEdge", "i4", ("nEdges",))[:] = f["nEdgesOnEdge"]
        ds.createVariable("cellsOnEdge", "i4", ("nEdges", "TWO"))
        ds.createVariable("verticesOnCell", "i4", ("nCells", "maxEdges"))
        ds.createVariable("verticesOnEdge", "i4", ("nEdges", "TWO"))
        if d_Prog is None:                                             # OutPut.jl:183-211
            ds.createVariable("ssh", "f8", ("nCells",))[:] = ssh
            ds.createVariable("layerThickness", "f8", ("nVertLevels", "nCells"))[:] = h.reshape(nVL, -1)
            ds.createVariable("normalVelocity", "f8", ("nVertLevels", "nEdges"))[:] = u.reshape(nVL, -1)
        else:                                                          # OutPut.jl:71-112
            ds.createVariable("ssh", "f8", ("time", "nCells"))[:] = ssh.reshape(1, -1)
            ds.createVariable("layerThickness", "f8", ("time", "nVertLevels", "nCells"))[:] = h.reshape(1, nVL, -1)
            ds.createVariable("normalVelocity", "f8", ("time", "nVertLevels", "nEdges"))[:] = u.reshape(1, nVL, -1)
            ds.createVariable("d_ssh", "f8", ("time", "nCells"))[:] = d_Prog.ssh.reshape(1, -1)
            ds.createVariable("d_layerThickness", "f8", ("time", "nVertLevels", "nCells"))[:] = d_Prog.layerThickness.reshape(1, nVL, -1)
            ds.createVariable("d_normalVelocity", "f8", ("time", "nVertLevels", "nEdges"))[:] = d_Prog.normalVelocity.reshape(1, nVL, -1)
