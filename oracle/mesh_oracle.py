"""ORACLE (test infrastructure, not product code): loop-based planar hex mesh + TRiSK weights.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.

The reference reads meshes from NetCDF (src/infra/MPASMesh/HorzMesh.jl:166-290) and computes
only the two sign fields itself (HorzMesh.jl:292-332); its meshes are network downloads that
are not on disk.  This file builds the mesh independently of the product generator
(moka_b200/planar_hex.py): plain Python loops, connectivity found by matching coordinates
instead of index formulas, and `weightsOnEdge` from the general TRiSK recipe (kite areas from
the geometry, Thuburn et al. 2009 / Ringler et al. 2010 as used by the MPAS mesh tools;
SURVEY.md Appendix B).  Small meshes only (pure Python; a 64x64 mesh takes a few seconds).

Arrays: numpy, shape (nEntities, nSlots) C-order == Julia (nSlots, nEntities) column-major;
Int32 1-based connectivity, 0 = absent.
"""
from __future__ import annotations

import math

import numpy as np


def _key(x, y, px, py, q):
    if px:
        x = x % px
        if abs(x - px) < q * 0.25:
            x = 0.0
    if py:
        y = y % py
        if abs(y - py) < q * 0.25:
            y = 0.0
    return (int(round(x / q * 4)), int(round(y / q * 4)))


def build_periodic_hex(nx: int, ny: int, dc: float, f0: float = 1e-4, H: float = 1000.0) -> dict:
    assert ny % 2 == 0
    s3 = math.sqrt(3.0)
    px, py = nx * dc, ny * dc * s3 / 2.0
    nC = nx * ny
    xC = np.empty(nC)
    yC = np.empty(nC)
    for j in range(ny):
        for i in range(nx):
            xC[j * nx + i] = dc * (i + 0.5 * (j % 2)) + 0.5 * dc
            yC[j * nx + i] = (j + 1) * dc * s3 / 2.0
    q = dc / 8.0
    cell_at = {_key(xC[c], yC[c], px, py, q): c for c in range(nC)}
    ang6 = [k * math.pi / 3.0 for k in range(6)]           # E NE NW W SW SE

    def nbr(c, a):
        return cell_at[_key(xC[c] + dc * math.cos(a), yC[c] + dc * math.sin(a), px, py, q)]

    # edges: owned by the cell on the tail side of the normal, 3 per cell (E, NE, NW)
    nE = 3 * nC
    xE, yE, aE = np.empty(nE), np.empty(nE), np.empty(nE)
    coe = np.zeros((nE, 2), np.int32)
    edge_at = {}
    for c in range(nC):
        for t in range(3):
            e = 3 * c + t
            a = ang6[t]
            xE[e] = xC[c] + 0.5 * dc * math.cos(a)
            yE[e] = yC[c] + 0.5 * dc * math.sin(a)
            aE[e] = a
            coe[e] = (c + 1, nbr(c, a) + 1)
            edge_at[_key(xE[e], yE[e], px, py, q)] = e
    # vertices: 2 per cell (90 deg, 30 deg)
    nV = 2 * nC
    r = dc / s3
    xV, yV = np.empty(nV), np.empty(nV)
    vert_at = {}
    for c in range(nC):
        for t, a in enumerate((math.pi / 2.0, math.pi / 6.0)):
            v = 2 * c + t
            xV[v] = xC[c] + r * math.cos(a)
            yV[v] = yC[c] + r * math.sin(a)
            vert_at[_key(xV[v], yV[v], px, py, q)] = v

    eoc = np.zeros((nC, 6), np.int32)
    coc = np.zeros((nC, 6), np.int32)
    voc = np.zeros((nC, 6), np.int32)
    for c in range(nC):
        for k in range(6):
            a = ang6[k]
            eoc[c, k] = edge_at[_key(xC[c] + 0.5 * dc * math.cos(a), yC[c] + 0.5 * dc * math.sin(a), px, py, q)] + 1
            coc[c, k] = nbr(c, a) + 1
            av = math.pi / 6.0 + k * math.pi / 3.0
            voc[c, k] = vert_at[_key(xC[c] + r * math.cos(av), yC[c] + r * math.sin(av), px, py, q)] + 1

    # verticesOnEdge ordered along t = k x n (tail, head); edgesOnVertex / cellsOnVertex by search
    voe = np.zeros((nE, 2), np.int32)
    for e in range(nE):
        tx, ty = -math.sin(aE[e]), math.cos(aE[e])
        half = 0.5 * dc / s3
        voe[e, 0] = vert_at[_key(xE[e] - half * tx, yE[e] - half * ty, px, py, q)] + 1
        voe[e, 1] = vert_at[_key(xE[e] + half * tx, yE[e] + half * ty, px, py, q)] + 1
    eov = np.zeros((nV, 3), np.int32)
    cov = np.zeros((nV, 3), np.int32)
    for v in range(nV):
        t = v % 2
        # edges leave a vertex at 120-degree spacing; up-pointing (t=0) and down-pointing stars
        eang = (-math.pi / 6.0, 7 * math.pi / 6.0, math.pi / 2.0) if t == 0 else \
               (3 * math.pi / 2.0, 5 * math.pi / 6.0, math.pi / 6.0)
        cang = (-math.pi / 2.0, math.pi / 6.0, 5 * math.pi / 6.0) if t == 0 else \
               (7 * math.pi / 6.0, -math.pi / 6.0, math.pi / 2.0)
        half = 0.5 * dc / s3
        for k in range(3):
            eov[v, k] = edge_at[_key(xV[v] + half * math.cos(eang[k]), yV[v] + half * math.sin(eang[k]), px, py, q)] + 1
            cov[v, k] = cell_at[_key(xV[v] + r * math.cos(cang[k]), yV[v] + r * math.sin(cang[k]), px, py, q)] + 1

    m = {
        "nCells": nC, "nEdges": nE, "nVertices": nV, "maxEdges": 6, "maxEdges2": 10,
        "vertexDegree": 3, "nVertLevels": 1, "is_periodic": "YES", "dc": dc,
        "x_period": px, "y_period": py,
        "xCell": xC, "yCell": yC, "zCell": np.zeros(nC), "fCell": np.full(nC, f0),
        "xEdge": xE, "yEdge": yE, "zEdge": np.zeros(nE), "fEdge": np.full(nE, f0),
        "xVertex": xV, "yVertex": yV, "zVertex": np.zeros(nV), "fVertex": np.full(nV, f0),
        "angleEdge": aE, "cellsOnEdge": coe, "verticesOnEdge": voe,
        "edgesOnCell": eoc, "cellsOnCell": coc, "verticesOnCell": voc,
        "edgesOnVertex": eov, "cellsOnVertex": cov,
        "nEdgesOnCell": np.full(nC, 6, np.int32),
        "dcEdge": np.full(nE, float(dc)), "dvEdge": np.full(nE, dc / s3),
        "areaCell": np.full(nC, s3 / 2.0 * dc * dc),
        "areaTriangle": np.full(nV, s3 / 4.0 * dc * dc),
        "minLevelCell": np.ones(nC, np.int32), "maxLevelCell": np.ones(nC, np.int32),
        "restingThickness": np.full((nC, 1), float(H)),
        "boundaryEdge": np.zeros(nE, np.int32),
    }
    trisk_weights(m)
    return m


def _kite_area(xc, yc, xa, ya, xv, yv, xb, yb):
    """Area of the quadrilateral cell-centre -> edge-a midpoint -> vertex -> edge-b midpoint."""
    xs = (xc, xa, xv, xb)
    ys = (yc, ya, yv, yb)
    s = 0.0
    for k in range(4):
        s += xs[k] * ys[(k + 1) % 4] - xs[(k + 1) % 4] * ys[k]
    return abs(s) * 0.5


def trisk_weights(m: dict) -> None:
    """General TRiSK edgesOnEdge / weightsOnEdge, in place.

    For edge e and side s in {1, 2} with c = cellsOnEdge[s, e]: walk counter-clockwise round c
    starting after e, accumulate r += kite(c, vertex passed) / areaCell[c], and emit
    w = sigma_s * (1/2 - r) * n_{e',c} * dvEdge[e'] / dcEdge[e], sigma_1 = +1, sigma_2 = -1,
    n_{e',c} = +1 if cellsOnEdge[1, e'] == c else -1.
    """
    nE, nC = m["nEdges"], m["nCells"]
    px, py = m["x_period"], m["y_period"]
    eoc, coe, voc = m["edgesOnCell"], m["cellsOnEdge"], m["verticesOnCell"]
    nEoC = m["nEdgesOnCell"]
    eoe = np.zeros((nE, m["maxEdges2"]), np.int32)
    woe = np.zeros((nE, m["maxEdges2"]))
    nEoE = np.zeros(nE, np.int32)

    def near(x, x0, p):            # periodic image of x nearest x0
        if p:
            x = x - p * round((x - x0) / p)
        return x

    for e in range(nE):
        slot = 0
        for s, sigma in ((0, 1.0), (1, -1.0)):
            c = coe[e, s] - 1
            if c < 0:
                continue
            n = nEoC[c]
            row = [eoc[c, k] - 1 for k in range(n)]
            j0 = row.index(e)
            xc, yc = m["xCell"][c], m["yCell"][c]
            rsum = 0.0
            for k in range(1, n):
                ja, jb = (j0 + k - 1) % n, (j0 + k) % n
                ea, eb = row[ja], row[jb]
                # the vertex between consecutive edges ja and jb (counter-clockwise): the one they share
                va = {m["verticesOnEdge"][ea, 0], m["verticesOnEdge"][ea, 1]}
                vb = {m["verticesOnEdge"][eb, 0], m["verticesOnEdge"][eb, 1]}
                (v1,) = va & vb
                v = v1 - 1
                assert v + 1 in set(voc[c])
                kite = _kite_area(xc, yc,
                                  near(m["xEdge"][ea], xc, px), near(m["yEdge"][ea], yc, py),
                                  near(m["xVertex"][v], xc, px), near(m["yVertex"][v], yc, py),
                                  near(m["xEdge"][eb], xc, px), near(m["yEdge"][eb], yc, py))
                rsum += kite / m["areaCell"][c]
                n_own = 1.0 if coe[eb, 0] - 1 == c else -1.0
                eoe[e, slot] = eb + 1
                woe[e, slot] = sigma * (0.5 - rsum) * n_own * m["dvEdge"][eb] / m["dcEdge"][e]
                slot += 1
        nEoE[e] = slot
    m["edgesOnEdge"], m["weightsOnEdge"], m["nEdgesOnEdge"] = eoe, woe, nEoE


def sign_index_fields(m: dict) -> None:
    """edgeSignOnCell / edgeSignOnVertex exactly as HorzMesh.jl:292-332 (loops, 1-based)."""
    nC, nV = m["nCells"], m["nVertices"]
    coe, voe = m["cellsOnEdge"], m.get("verticesOnEdge")
    esc = np.zeros((nC, m["maxEdges"]), np.int32)
    for iCell in range(1, nC + 1):
        for i in range(1, m["nEdgesOnCell"][iCell - 1] + 1):
            iEdge = m["edgesOnCell"][iCell - 1, i - 1]
            esc[iCell - 1, i - 1] = -1 if iCell == coe[iEdge - 1, 0] else 1   # :302-306
    m["edgeSignOnCell"] = esc
    if nV and voe is not None:
        esv = np.zeros((nV, m["maxEdges"]), np.int32)
        for iVertex in range(1, nV + 1):
            for i in range(1, m["vertexDegree"] + 1):
                iEdge = m["edgesOnVertex"][iVertex - 1, i - 1]
                esv[iVertex - 1, i - 1] = -1 if iVertex == voe[iEdge - 1, 0] else 1   # :323-327
        m["edgeSignOnVertex"] = esv
