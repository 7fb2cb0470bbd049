"""ORACLE (test infrastructure): ctypes binding of oracle/libmoka_oracle.so (moka_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this.  Takes the mesh dict (reference names/layouts) produced by the generators.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_I32P, _F64P = C.POINTER(C.c_int32), C.POINTER(C.c_double)


class _Mesh(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ("nCells", "nEdges", "nVertices", "maxEdges", "maxEdges2", "vertexDegree")] + [
        ("cellsOnEdge", _I32P), ("edgesOnEdge", _I32P), ("nEdgesOnEdge", _I32P), ("weightsOnEdge", _F64P),
        ("dcEdge", _F64P), ("dvEdge", _F64P), ("fEdge", _F64P),
        ("edgesOnCell", _I32P), ("edgeSignOnCell", _I32P), ("nEdgesOnCell", _I32P), ("areaCell", _F64P),
        ("edgesOnVertex", _I32P), ("edgeSignOnVertex", _I32P), ("areaTriangle", _F64P),
        ("maxLevelEdgeTop", _I32P), ("restingThicknessSum", _F64P)]


class _State(C.Structure):
    _fields_ = [("ssh", _F64P * 2), ("normalVelocity", _F64P * 2), ("layerThickness", _F64P * 2)] + [
        (n, _F64P) for n in ("layerThicknessEdge", "thicknessFlux", "velocityDivCell", "relativeVorticity",
                             "tendNormalVelocity", "tendLayerThickness",
                             "uProvis", "hProvis", "sshProvis", "uNew", "hNew")]


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libmoka_oracle.so")
    src = os.path.join(_HERE, "moka_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libmoka_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.ora_sum_array.restype = C.c_double
        _LIB.ora_sum_array.argtypes = [_F64P, C.c_int64]
        _LIB.ora_num_threads.restype = C.c_int
    return _LIB


def set_num_threads(n: int) -> None:
    """OpenMP team size of the C oracle (all host cores for the benchmark's reference arm, whatever OMP_NUM_THREADS says)."""
    lib().ora_set_num_threads(int(n))


def _p(a):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_I32P if a.dtype == np.int32 else _F64P)


def sign_index_fields(m: dict) -> None:
    """signIndexField!, HorzMesh.jl:292-332, vectorised (for meshes too big for the loop version)."""
    nC = m["nCells"]
    eoc = m["edgesOnCell"].astype(np.int64) - 1
    act = np.arange(m["maxEdges"])[None, :] < m["nEdgesOnCell"][:, None]
    c1 = m["cellsOnEdge"][np.where(act, eoc, 0), 0]
    m["edgeSignOnCell"] = np.where(act, np.where(c1 == (np.arange(nC) + 1)[:, None], -1, 1), 0).astype(np.int32)
    if m.get("nVertices", 0) and "edgesOnVertex" in m:
        nV = m["nVertices"]
        eov = m["edgesOnVertex"].astype(np.int64) - 1
        v1 = m["verticesOnEdge"][eov, 0]
        s = np.zeros((nV, m["maxEdges"]), np.int32)
        s[:, :m["vertexDegree"]] = np.where(v1 == (np.arange(nV) + 1)[:, None], -1, 1)
        m["edgeSignOnVertex"] = s


def apply_boundary_mask(m: dict) -> dict:
    """Project-defined solid-wall treatment, as mesh preprocessing so the plain
    reference kernels need no branch: a `boundaryEdge` gets cellsOnEdge[2] := cellsOnEdge[1] (zero
    SSH gradient, hEdge = h of the one cell) and a zero `weightsOnEdge` row (no Coriolis), so its
    tendency is exactly 0 and u stays 0.  Returns a shallow copy with the two arrays replaced."""
    out = dict(m)
    b = m["boundaryEdge"] != 0
    coe = m["cellsOnEdge"].copy()
    coe[b, 1] = coe[b, 0]
    w = m["weightsOnEdge"].copy()
    w[b, :] = 0.0
    out["cellsOnEdge"], out["weightsOnEdge"] = coe, w
    return out


class OracleModel:
    """Mesh + state living in numpy arrays, stepped by the C restatement."""

    def __init__(self, m: dict, ssh, u, h):
        self.L = lib()
        if "edgeSignOnCell" not in m:
            sign_index_fields(m)
        self.m = m
        nC, nE, nV = m["nCells"], m["nEdges"], m.get("nVertices", 0)
        self.keep = {
            "top": np.ones(nE, np.int32),                                 # VertMesh.jl:31-36
            "H": np.ascontiguousarray(m["restingThickness"].sum(axis=1)),  # VertMesh.jl:73
        }
        cm = _Mesh()
        cm.nCells, cm.nEdges, cm.nVertices = nC, nE, nV
        cm.maxEdges, cm.maxEdges2, cm.vertexDegree = m["maxEdges"], m["maxEdges2"], m["vertexDegree"]
        for k in ("cellsOnEdge", "edgesOnEdge", "nEdgesOnEdge", "weightsOnEdge", "dcEdge", "dvEdge", "fEdge",
                  "edgesOnCell", "edgeSignOnCell", "nEdgesOnCell", "areaCell"):
            self.keep[k] = np.ascontiguousarray(m[k])
            setattr(cm, k, _p(self.keep[k]))
        if nV:
            for k in ("edgesOnVertex", "edgeSignOnVertex", "areaTriangle"):
                self.keep[k] = np.ascontiguousarray(m[k])
                setattr(cm, k, _p(self.keep[k]))
        cm.maxLevelEdgeTop = _p(self.keep["top"])
        cm.restingThicknessSum = _p(self.keep["H"])
        self.cm = cm
        f8 = lambda n: np.zeros(n)
        self.ssh = [np.array(ssh, dtype=np.float64), np.array(ssh, dtype=np.float64)]
        self.normalVelocity = [np.array(u, dtype=np.float64), np.array(u, dtype=np.float64)]
        self.layerThickness = [np.array(h, dtype=np.float64), np.array(h, dtype=np.float64)]
        self.layerThicknessEdge, self.thicknessFlux = f8(nE), f8(nE)
        self.velocityDivCell, self.relativeVorticity = f8(nC), f8(max(nV, 1))
        self.tendNormalVelocity, self.tendLayerThickness = f8(nE), f8(nC)
        self._work = [f8(nE), f8(nC), f8(nC), f8(nE), f8(nC)]
        cs = _State()
        for k in ("ssh", "normalVelocity", "layerThickness"):
            arr = getattr(self, k)
            setattr(cs, k, (_F64P * 2)(_p(arr[0]), _p(arr[1])))
        for k in ("layerThicknessEdge", "thicknessFlux", "velocityDivCell", "relativeVorticity",
                  "tendNormalVelocity", "tendLayerThickness"):
            setattr(cs, k, _p(getattr(self, k)))
        for k, a in zip(("uProvis", "hProvis", "sshProvis", "uNew", "hNew"), self._work):
            setattr(cs, k, _p(a))
        self.cs = cs

    # driver-level entry points -------------------------------------------------------------
    def run_loop(self, dt: float, nsteps: int, stepper: str = "RungeKutta4") -> None:
        self.L.ora_run_loop(C.byref(self.cm), C.byref(self.cs), C.c_double(dt), C.c_int64(nsteps),
                            C.c_int(0 if stepper == "ForwardEuler" else 1))

    def diagnostic_compute(self) -> None:
        self.L.ora_diagnostic_compute(C.byref(self.cm), C.byref(self.cs), _p(self.normalVelocity[1]),
                                      _p(self.layerThickness[1]))

    def compute_normal_velocity_tendency(self) -> np.ndarray:
        self.L.ora_compute_normal_velocity_tendency(C.byref(self.cm), _p(self.tendNormalVelocity),
                                                    _p(self.ssh[1]), _p(self.normalVelocity[1]))
        return self.tendNormalVelocity

    def compute_layer_thickness_tendency(self) -> np.ndarray:
        self.L.ora_compute_layer_thickness_tendency(C.byref(self.cm), _p(self.tendLayerThickness),
                                                    _p(self.thicknessFlux))
        return self.tendLayerThickness

    def sum_ssh2(self) -> float:
        return float(self.L.ora_sum_array(_p(self.ssh[1]), C.c_int64(self.m["nCells"])))

    def num_threads(self) -> int:
        return int(self.L.ora_num_threads())
