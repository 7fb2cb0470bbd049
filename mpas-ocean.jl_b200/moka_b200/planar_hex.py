"""Synthetic MPAS-spec planar hexagonal meshes (vectorised numpy).

The reference has no mesh generator: `ReadHorzMesh` (src/infra/MPASMesh/HorzMesh.jl:334-355)
only reads MPAS NetCDF files, and every mesh its tests use is a network download
(test/ocn/test_Operators.jl:12-15, test/Artifacts.toml:1-6).  This module produces the
same variables, with the same names, dtypes and (slot, entity) column-major layouts the
readers at HorzMesh.jl:166-290 and VertMesh.jl:46-82 return, for the regular meshes
BASELINE.json's configs name:

  * ``periodic_hex(nx, ny, dc)``       doubly periodic (inertial-gravity-wave configs)
  * ``channel_hex(nx, ny, dc)``        periodic in y, solid walls in x (coastal Kelvin wave)

Array convention: numpy arrays are C-ordered with shape ``(nEntities, nSlots)``, which is
byte-identical to the Julia ``(nSlots, nEntities)`` column-major arrays of the reference.
Connectivity is Int32, 1-based, 0 = absent, exactly as the NetCDF files hold it.

Layout (0-based here):
  cell (i, j) -> id j*nx+i, centre x = dc*(i + (j%2)/2) + dc/2, y = (j+1)*dc*sqrt(3)/2
  edges 3c+t, t = 0:E (angle 0), 1:NE (pi/3), 2:NW (2pi/3), cellsOnEdge = (c, neighbour)
  vertices 2c (at 90 deg) and 2c+1 (at 30 deg), distance dc/sqrt(3)
"""
from __future__ import annotations

import numpy as np

SQRT3 = np.sqrt(3.0)


def _nbrs(nx: int, ny: int):
    j, i = np.divmod(np.arange(nx * ny, dtype=np.int64), nx)
    o = j & 1

    def cid(ii, jj):
        return (jj % ny) * nx + (ii % nx)

    return {
        "i": i, "j": j, "o": o,
        "E": cid(i + 1, j), "W": cid(i - 1, j),
        "NE": cid(i + o, j + 1), "NW": cid(i - 1 + o, j + 1),
        "SE": cid(i + o, j - 1), "SW": cid(i - 1 + o, j - 1),
    }


def periodic_hex(nx: int, ny: int, dc: float, f0: float = 1.0e-4,
                 resting_thickness: float = 1000.0, with_dual: bool = True) -> dict:
    """Doubly periodic regular hex mesh with the field names of the MPAS mesh spec.

    Returns a dict of numpy arrays (reference layouts/dtypes).  `with_dual=False` skips the
    vertex arrays (only the curl diagnostic uses them) to save host memory on 4096x4096.
    """
    if ny % 2:
        raise ValueError("ny must be even for a periodic hex mesh")
    nC = nx * ny
    nb = _nbrs(nx, ny)
    c = np.arange(nC, dtype=np.int64)
    i, j = nb["i"], nb["j"]

    m: dict = {"nCells": nC, "nEdges": 3 * nC, "nVertices": 2 * nC if with_dual else 0, "maxEdges": 6,
               "maxEdges2": 10, "vertexDegree": 3, "nVertLevels": 1, "is_periodic": "YES",
               "x_period": nx * dc, "y_period": ny * dc * SQRT3 / 2.0, "dc": float(dc),
               "nx": nx, "ny": ny}

    xC = dc * (i + 0.5 * (j & 1)) + 0.5 * dc
    yC = (j + 1) * (dc * SQRT3 / 2.0)
    m["xCell"], m["yCell"], m["zCell"] = xC, yC, np.zeros(nC)
    m["fCell"] = np.full(nC, f0)
    m["areaCell"] = np.full(nC, SQRT3 / 2.0 * dc * dc)
    m["nEdgesOnCell"] = np.full(nC, 6, np.int32)

    # --- edges -----------------------------------------------------------------------
    nE = 3 * nC
    ang = np.array([0.0, np.pi / 3.0, 2.0 * np.pi / 3.0])
    coe = np.empty((nC, 3, 2), np.int32)
    coe[:, :, 0] = (c + 1)[:, None]
    coe[:, 0, 1] = nb["E"] + 1
    coe[:, 1, 1] = nb["NE"] + 1
    coe[:, 2, 1] = nb["NW"] + 1
    m["cellsOnEdge"] = coe.reshape(nE, 2)
    m["angleEdge"] = np.tile(ang, nC)
    m["xEdge"] = (xC[:, None] + 0.5 * dc * np.cos(ang)[None, :]).reshape(nE)
    m["yEdge"] = (yC[:, None] + 0.5 * dc * np.sin(ang)[None, :]).reshape(nE)
    m["zEdge"] = np.zeros(nE)
    m["fEdge"] = np.full(nE, f0)
    m["dcEdge"] = np.full(nE, float(dc))
    m["dvEdge"] = np.full(nE, dc / SQRT3)

    # edgesOnCell counter-clockwise from east; cellsOnCell across the same edge
    eoc0 = np.stack([3 * c, 3 * c + 1, 3 * c + 2, 3 * nb["W"], 3 * nb["SW"] + 1,
                     3 * nb["SE"] + 2], axis=1)
    m["edgesOnCell"] = (eoc0 + 1).astype(np.int32)
    m["cellsOnCell"] = (np.stack([nb["E"], nb["NE"], nb["NW"], nb["W"], nb["SW"], nb["SE"]],
                                 axis=1) + 1).astype(np.int32)
    m["verticesOnCell"] = (np.stack([2 * c + 1, 2 * c, 2 * nb["W"] + 1, 2 * nb["SW"],
                                     2 * nb["SW"] + 1, 2 * nb["SE"]], axis=1) + 1).astype(np.int32)

    # --- TRiSK edgesOnEdge / weightsOnEdge (SURVEY.md Appendix B recipe) -----------------
    # slots 0..4: the other five edges of cell 1 walked counter-clockwise from e,
    # slots 5..9: the other five edges of cell 2.  Regular hex: kite fraction r = k/6.
    # An edge 3c+t sits at position t of edgesOnCell[c] (cell 1) and at position t+3 of
    # edgesOnCell[cell 2]; the owner flag n_{e',c} is +1 for positions 0..2, -1 for 3..5.
    eoe = np.empty((nC, 3, 10), np.int32)
    woe = np.empty((nC, 3, 10))
    c2 = (m["cellsOnEdge"][:, 1].astype(np.int64) - 1).reshape(nC, 3)
    dvdc = (dc / SQRT3) / dc
    for t in range(3):
        for k in range(1, 6):
            p1 = (t + k) % 6
            eoe[:, t, k - 1] = eoc0[:, p1] + 1
            woe[:, t, k - 1] = +1.0 * (0.5 - k / 6.0) * (1.0 if p1 < 3 else -1.0) * dvdc
            p2 = (t + 3 + k) % 6
            eoe[:, t, 4 + k] = eoc0[c2[:, t], p2] + 1
            woe[:, t, 4 + k] = -1.0 * (0.5 - k / 6.0) * (1.0 if p2 < 3 else -1.0) * dvdc
    m["edgesOnEdge"] = eoe.reshape(nE, 10)
    m["weightsOnEdge"] = woe.reshape(nE, 10)
    m["nEdgesOnEdge"] = np.full(nE, 10, np.int32)

    # --- dual mesh (curl diagnostic only) -------------------------------------------------
    if with_dual:
        nV = 2 * nC
        r = dc / SQRT3
        vang = np.array([np.pi / 2.0, np.pi / 6.0])
        m["xVertex"] = (xC[:, None] + r * np.cos(vang)[None, :]).reshape(nV)
        m["yVertex"] = (yC[:, None] + r * np.sin(vang)[None, :]).reshape(nV)
        m["zVertex"] = np.zeros(nV)
        m["fVertex"] = np.full(nV, f0)
        m["areaTriangle"] = np.full(nV, SQRT3 / 4.0 * dc * dc)
        eov = np.empty((nC, 2, 3), np.int64)
        eov[:, 0] = np.stack([3 * c + 1, 3 * c + 2, 3 * nb["NW"]], axis=1)
        eov[:, 1] = np.stack([3 * c, 3 * c + 1, 3 * nb["E"] + 2], axis=1)
        m["edgesOnVertex"] = (eov.reshape(nV, 3) + 1).astype(np.int32)
        cov = np.empty((nC, 2, 3), np.int64)
        cov[:, 0] = np.stack([c, nb["NE"], nb["NW"]], axis=1)
        cov[:, 1] = np.stack([c, nb["E"], nb["NE"]], axis=1)
        m["cellsOnVertex"] = (cov.reshape(nV, 3) + 1).astype(np.int32)
        voe = np.empty((nC, 3, 2), np.int64)
        voe[:, 0] = np.stack([2 * nb["SE"], 2 * c + 1], axis=1)
        voe[:, 1] = np.stack([2 * c + 1, 2 * c], axis=1)
        voe[:, 2] = np.stack([2 * c, 2 * nb["W"] + 1], axis=1)
        m["verticesOnEdge"] = (voe.reshape(nE, 2) + 1).astype(np.int32)
        m["kiteAreasOnVertex"] = np.full((nV, 3), SQRT3 / 12.0 * dc * dc)

    # --- vertical mesh (VertMesh.jl:46-82: single stacked layer) -----------------------------
    m["minLevelCell"] = np.ones(nC, np.int32)
    m["maxLevelCell"] = np.ones(nC, np.int32)
    m["restingThickness"] = np.full((nC, 1), float(resting_thickness))
    m["boundaryEdge"] = np.zeros(nE, np.int32)
    return m


def channel_hex(nx: int, ny: int, dc: float, f0: float = 1.0e-4,
                resting_thickness: float = 1000.0) -> dict:
    """Hex mesh periodic in y with solid walls at the west and east ends of every row.

    Built from `periodic_hex` by cutting every edge that wraps around in x.  A cut edge keeps
    its first cell and gets ``cellsOnEdge[2] = 0`` (``boundaryEdge = 1``); its TRiSK stencil
    keeps only the five edges of the remaining cell (``nEdgesOnEdge = 5``).  The reference
    itself rejects such meshes (VertMesh.jl:50-52); the masked treatment follows the glossary
    in the legacy src/infra/Mesh.jl:110-114 and is project-defined.
    Vertex arrays are not produced (the curl diagnostic is defined on periodic meshes only).
    """
    m = periodic_hex(nx, ny, dc, f0, resting_thickness, with_dual=False)
    nC, nE = m["nCells"], m["nEdges"]
    coe = m["cellsOnEdge"].astype(np.int64) - 1
    x1, x2 = m["xCell"][coe[:, 0]], m["xCell"][coe[:, 1]]
    cut = np.abs(x2 - x1) > 2.0 * dc            # the edge wraps around the x period
    # a cut edge is seen by two cells: it stays with cell 1 (the owner, cell id = e//3) and
    # cell 2 gets a fresh boundary edge appended at the end of the edge list.
    cut_ids = np.nonzero(cut)[0]
    nNew = cut_ids.size
    new_ids = nE + np.arange(nNew, dtype=np.int64)
    other = coe[cut_ids, 1]

    def grow(a, fill):
        out = np.empty((nE + nNew,) + a.shape[1:], a.dtype)
        out[:nE] = a
        out[nE:] = fill
        return out

    ang = m["angleEdge"][cut_ids] + np.pi       # new edge: normal points out of `other`
    xe = m["xCell"][other] + 0.5 * dc * np.cos(ang)
    ye = m["yCell"][other] + 0.5 * dc * np.sin(ang)
    m["angleEdge"] = grow(m["angleEdge"], ang)
    m["xEdge"], m["yEdge"] = grow(m["xEdge"], xe), grow(m["yEdge"], ye)
    m["zEdge"] = grow(m["zEdge"], 0.0)
    m["fEdge"] = grow(m["fEdge"], f0)
    m["dcEdge"] = grow(m["dcEdge"], float(dc))
    m["dvEdge"] = grow(m["dvEdge"], dc / SQRT3)
    coe_new = np.stack([other + 1, np.zeros_like(other)], axis=1).astype(np.int32)
    m["cellsOnEdge"] = grow(m["cellsOnEdge"], coe_new)
    m["cellsOnEdge"][cut_ids, 1] = 0
    m["boundaryEdge"] = grow(m["boundaryEdge"], 1)
    m["boundaryEdge"][cut_ids] = 1

    # redirect `other`'s edgesOnCell slot from the cut edge to its new private edge
    eoc = m["edgesOnCell"].astype(np.int64) - 1
    coc = m["cellsOnCell"].copy()
    for slot in range(6):
        e_here = eoc[other, slot]
        hit = e_here == cut_ids
        eoc[other[hit], slot] = new_ids[hit]
        coc[other[hit], slot] = 0
    owner = coe[cut_ids, 0]
    for slot in range(6):
        hit = eoc[owner, slot] == cut_ids
        coc[owner[hit], slot] = 0
    m["edgesOnCell"] = (eoc + 1).astype(np.int32)
    m["cellsOnCell"] = coc
    nE2 = nE + nNew
    m["nEdges"] = nE2

    # rebuild TRiSK stencils with the general recipe (regular hex: kite fraction 1/6)
    coe = m["cellsOnEdge"].astype(np.int64) - 1
    eoe = np.zeros((nE2, 10), np.int32)
    woe = np.zeros((nE2, 10))
    nEoE = np.zeros(nE2, np.int32)
    e_all = np.arange(nE2, dtype=np.int64)
    dvdc = (dc / SQRT3) / dc
    for s, sigma in ((0, 1.0), (1, -1.0)):
        cs = coe[:, s]
        ok = cs >= 0
        rows = eoc[np.where(ok, cs, 0)]                     # (nE2, 6)
        j0 = np.argmax(rows == e_all[:, None], axis=1)
        for k in range(1, 6):
            ep = rows[e_all, (j0 + k) % 6]
            n_own = np.where(coe[ep, 0] == cs, 1.0, -1.0)
            slot = nEoE.astype(np.int64)
            w = sigma * (0.5 - k / 6.0) * n_own * dvdc
            eoe[e_all[ok], slot[ok]] = (ep[ok] + 1).astype(np.int32)
            woe[e_all[ok], slot[ok]] = w[ok]
            nEoE[ok] += 1
    m["edgesOnEdge"], m["weightsOnEdge"], m["nEdgesOnEdge"] = eoe, woe, nEoE
    m["is_periodic"] = "NO"
    m["x_period"] = 0.0
    m["nVertices"] = 0
    return m
