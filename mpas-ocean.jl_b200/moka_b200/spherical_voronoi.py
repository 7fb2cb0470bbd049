"""Quasi-uniform SPHERICAL Voronoi mesh in the MPAS layout: what MPAS-Ocean meshes actually are.

Generators on a Fibonacci lattice, the Voronoi diagram read off their convex hull (on a sphere the hull's triangles are the
Delaunay triangulation; a triangle's circumcentre direction is a Voronoi vertex), great-circle lengths, spherical kite /
cell / triangle areas, fEdge = 2 Omega sin(lat) -- so the Coriolis weights are NOT uniform and the library's "folded"
kernels (weightsOnEdge * fEdge[eoe]) run, cells are a mix of pentagons, hexagons and heptagons (plus a few squares at the
poles), coordinates are three-dimensional (the renumbering takes its Morton branch), and nothing is periodic.  Arrays follow the
MPAS mesh specification as the reference reads it (HorzMesh.jl:166-290), conventions as in planar_voronoi.py:

  * cellsOnEdge[e] = (c1, c2) with c1 < c2, the normal points from c1 to c2; angleEdge = its angle from local east at the edge;
  * edgesOnCell / cellsOnCell / verticesOnCell counter-clockwise seen from outside;  verticesOnEdge along t = k x n;
  * edgesOnEdge / weightsOnEdge: TRiSK with spherical kite areas.

Vectorised (sorts and segment operations); host-side tool for tests and benchmarks.
"""
from __future__ import annotations

import numpy as np

EARTH_RADIUS, EARTH_OMEGA = 6371220.0, 7.292e-5


def _unit(v: np.ndarray) -> np.ndarray:
    return v / np.linalg.norm(v, axis=1)[:, None]


def _tri_area(a: np.ndarray, b: np.ndarray, c: np.ndarray) -> np.ndarray:
    """Signed area of the spherical triangles (a, b, c) on the unit sphere (Van Oosterom & Strackee), positive when
    counter-clockwise seen from outside."""
    det = np.einsum("ij,ij->i", a, np.cross(b, c))
    den = 1.0 + np.einsum("ij,ij->i", a, b) + np.einsum("ij,ij->i", b, c) + np.einsum("ij,ij->i", c, a)
    return 2.0 * np.arctan2(det, den)


def _arc(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    return 2.0 * np.arcsin(np.clip(0.5 * np.linalg.norm(a - b, axis=1), 0.0, 1.0))


def fibonacci_sphere(n: int) -> np.ndarray:
    i = np.arange(n) + 0.5
    phi, th = np.arccos(1.0 - 2.0 * i / n), np.pi * (1.0 + 5.0 ** 0.5) * i
    return np.stack([np.cos(th) * np.sin(phi), np.sin(th) * np.sin(phi), np.cos(phi)], axis=1)


def spherical_voronoi(n: int, radius: float = EARTH_RADIUS, omega: float = EARTH_OMEGA, resting_thickness: float = 1000.0,
                      with_dual: bool = True) -> dict:
    """`n` cells on a sphere of `radius`; Coriolis parameter 2 `omega` sin(latitude)."""
    from scipy.spatial import ConvexHull
    if n < 32:
        raise ValueError("spherical_voronoi: at least 32 cells")
    p = fibonacci_sphere(n)
    tri = ConvexHull(p).simplices.astype(np.int64)
    a, b, c = p[tri[:, 0]], p[tri[:, 1]], p[tri[:, 2]]
    flip = np.einsum("ij,ij->i", a, np.cross(b, c)) < 0
    tri[flip] = tri[flip][:, [0, 2, 1]]                                             # counter-clockwise seen from outside
    a, b, c = p[tri[:, 0]], p[tri[:, 1]], p[tri[:, 2]]
    vert = _unit(np.cross(b - a, c - a))                                            # circumcentre direction = Voronoi vertex
    nV = len(tri)
    # ---- directed sides: a -> b of a counter-clockwise triangle has that triangle's vertex on its LEFT -------------------
    da = np.concatenate([tri[:, 0], tri[:, 1], tri[:, 2]])
    db = np.concatenate([tri[:, 1], tri[:, 2], tri[:, 0]])
    vleft = np.concatenate([np.arange(nV)] * 3)
    # angle of the side in the tangent plane of its start cell (east, north), to order the sides counter-clockwise
    pa, pb = p[da], p[db]
    east = np.cross(np.array([0.0, 0.0, 1.0]), pa)
    pole = np.linalg.norm(east, axis=1) < 1e-12
    east[pole] = np.array([1.0, 0.0, 0.0])
    east = _unit(east)
    north = np.cross(pa, east)
    d = pb - np.einsum("ij,ij->i", pb, pa)[:, None] * pa
    ang = np.arctan2(np.einsum("ij,ij->i", d, north), np.einsum("ij,ij->i", d, east))
    o = np.lexsort((ang, da))
    da, db, vleft = da[o], db[o], vleft[o]
    nD = len(da)
    nEoC = np.bincount(da, minlength=n).astype(np.int32)
    start = np.concatenate([[0], np.cumsum(nEoC)]).astype(np.int64)
    posn = np.arange(nD) - start[da]
    nxt = start[da] + (posn + 1) % nEoC[da]
    # ---- edges -----------------------------------------------------------------------------------------------------------
    lo, hi = np.minimum(da, db), np.maximum(da, db)
    ekeys, eid = np.unique(lo * n + hi, return_inverse=True)
    nE = len(ekeys)
    if nV - nE + n != 2:
        raise RuntimeError(f"spherical_voronoi: Euler characteristic of the sphere violated (V - E + F = {nV - nE + n})")
    fwd = da < db
    ef = eid[fwd]
    coe = np.zeros((nE, 2), np.int32)
    coe[ef, 0], coe[ef, 1] = da[fwd] + 1, db[fwd] + 1
    c1, c2 = p[coe[:, 0] - 1], p[coe[:, 1] - 1]
    mid = _unit(c1 + c2)                                                            # the edge point: great-circle midpoint of the two centres
    voe = np.zeros((nE, 2), np.int32)
    voe[ef, 1] = vleft[fwd] + 1                                                     # left of c1 -> c2
    voe[eid[~fwd], 0] = vleft[~fwd] + 1                                             # left of c2 -> c1 = right of c1 -> c2
    dcE = radius * _arc(c1, c2)
    dvE = radius * _arc(vert[voe[:, 0] - 1], vert[voe[:, 1] - 1])
    # normal at the edge point: the direction from c1 to c2, tangent to the sphere; its angle from local east
    nrm = _unit(c2 - c1 - np.einsum("ij,ij->i", c2 - c1, mid)[:, None] * mid)
    e_e = np.cross(np.array([0.0, 0.0, 1.0]), mid)
    e_e = _unit(np.where((np.linalg.norm(e_e, axis=1) < 1e-12)[:, None], np.array([1.0, 0.0, 0.0]), e_e))
    e_n = np.cross(mid, e_e)
    angE = np.arctan2(np.einsum("ij,ij->i", nrm, e_n), np.einsum("ij,ij->i", nrm, e_e))
    # ---- per cell rows, kites ------------------------------------------------------------------------------------------
    S = int(nEoC.max())
    eoc, coc, voc = (np.zeros((n, S), np.int32) for _ in range(3))
    eoc[da, posn], coc[da, posn], voc[da, posn] = eid + 1, db + 1, vleft + 1
    ctr, m0, m1, rv = p[da], mid[eid], mid[eid[nxt]], vert[vleft]
    kite = _tri_area(ctr, m0, rv) + _tri_area(ctr, rv, m1)                          # (centre, edge point k, vertex, edge point k + 1)
    if kite.min() <= 0.0:
        raise RuntimeError("spherical_voronoi: non-positive kite area")
    area1 = np.bincount(da, weights=kite, minlength=n)                              # on the unit sphere
    # ---- TRiSK -----------------------------------------------------------------------------------------------------------
    S2 = 2 * S - 2
    eoe = np.zeros((nE, S2), np.int32)
    woe = np.zeros((nE, S2))
    n_c = nEoC[da].astype(np.int64)
    n1 = nEoC[coe[:, 0] - 1].astype(np.int64)
    slot0 = np.where(fwd, 0, n1[eid] - 1)
    sigma = np.where(fwd, 1.0, -1.0)
    rsum = np.zeros(nD)
    for kk in range(1, S):
        act = kk < n_c
        passed = start[da] + (posn + kk - 1) % n_c
        rsum = rsum + kite[passed] / area1[da]
        tgt = start[da] + (posn + kk) % n_c
        e2 = eid[tgt]
        owner = np.where(coe[e2, 0] - 1 == da, 1.0, -1.0)
        wv = sigma * (0.5 - rsum) * owner * dvE[e2] / dcE[eid]
        sl = slot0 + kk - 1
        eoe[eid[act], sl[act]] = e2[act] + 1
        woe[eid[act], sl[act]] = wv[act]
    nEoE = (n1 + nEoC[coe[:, 1] - 1] - 2).astype(np.int32)

    lat = lambda q: np.arcsin(np.clip(q[:, 2], -1.0, 1.0))                            # noqa: E731
    lon = lambda q: np.arctan2(q[:, 1], q[:, 0])                                      # noqa: E731
    m: dict = {"nCells": n, "nEdges": nE, "nVertices": nV if with_dual else 0, "maxEdges": S, "maxEdges2": S2, "vertexDegree": 3,
               "nVertLevels": 1, "is_periodic": "NO", "on_a_sphere": "YES", "sphere_radius": float(radius),
               "dc": float(np.mean(dcE))}
    m["xCell"], m["yCell"], m["zCell"] = radius * p[:, 0], radius * p[:, 1], radius * p[:, 2]
    m["latCell"], m["lonCell"] = lat(p), lon(p)
    m["fCell"], m["areaCell"], m["nEdgesOnCell"] = 2.0 * omega * p[:, 2], radius * radius * area1, nEoC
    m["cellsOnEdge"], m["angleEdge"] = coe, angE
    m["xEdge"], m["yEdge"], m["zEdge"] = radius * mid[:, 0], radius * mid[:, 1], radius * mid[:, 2]
    m["latEdge"], m["lonEdge"], m["fEdge"] = lat(mid), lon(mid), 2.0 * omega * mid[:, 2]
    m["dcEdge"], m["dvEdge"] = dcE, dvE
    m["edgesOnCell"], m["cellsOnCell"] = eoc, coc
    m["edgesOnEdge"], m["weightsOnEdge"], m["nEdgesOnEdge"] = eoe, woe, nEoE
    if with_dual:
        m["verticesOnEdge"], m["verticesOnCell"] = voe, voc
        m["xVertex"], m["yVertex"], m["zVertex"] = radius * vert[:, 0], radius * vert[:, 1], radius * vert[:, 2]
        m["fVertex"] = 2.0 * omega * vert[:, 2]
        m["cellsOnVertex"] = (tri + 1).astype(np.int32)
        kv = np.zeros((nV, 3))
        eov = np.zeros((nV, 3), np.int32)
        for j in range(3):
            sel = np.nonzero(tri[vleft, j] == da)[0]                                 # the side leaving corner j with the vertex on its left
            kv[vleft[sel], j] = kite[sel]
            eov[vleft[sel], j] = eid[sel] + 1
        m["kiteAreasOnVertex"], m["edgesOnVertex"], m["areaTriangle"] = radius * radius * kv, eov, radius * radius * kv.sum(axis=1)
    m["minLevelCell"] = np.ones(n, np.int32)
    m["maxLevelCell"] = np.ones(n, np.int32)
    m["restingThickness"] = np.full((n, 1), float(resting_thickness))
    m["boundaryEdge"] = np.zeros(nE, np.int32)
    return m


def geostrophic_zonal_flow(m: dict, u0: float = 20.0, h0: float = 1000.0, omega: float = EARTH_OMEGA, g: float = 9.80616):
    """A steady state of the linear rotating shallow-water equations the library integrates: solid-body zonal wind
    u = u0 cos(lat) in geostrophic balance with h = h0 - (R Omega u0 / g) sin^2(lat)  (Williamson et al. 1992, test case 2,
    without the u0^2 / 2 term of the nonlinear equations).  Returns (ssh, normalVelocity, layerThickness)."""
    R = m["sphere_radius"]
    ssh = -(R * omega * u0 / g) * np.sin(m["latCell"]) ** 2
    un = u0 * np.cos(m["latEdge"]) * np.cos(m["angleEdge"])                        # eastward wind projected on the edge normal
    return ssh, un, h0 + ssh
