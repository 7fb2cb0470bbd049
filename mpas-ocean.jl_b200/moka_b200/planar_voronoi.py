"""Doubly periodic planar VORONOI mesh generator: a genuinely unstructured MPAS C-grid for the tests.

The regular generators (planar_hex.py) only ever produce hexagons with identical metrics, so every parity test on them
exercises one stencil shape.  Here the cell centres are a jittered hex lattice and the mesh is the periodic Voronoi
diagram of those points (scipy.spatial.Voronoi on a 3x3 tiling): pentagons, hexagons and heptagons, every dcEdge /
dvEdge / areaCell different, edgesOnEdge rows of 8..12 entries, TRiSK weights from real kite areas.  The arrays follow the
MPAS mesh specification as the reference reads it (HorzMesh.jl:166-290):

  * cellsOnEdge[e] = (c1, c2), the normal points from c1 to c2, angleEdge its angle;
  * edgesOnCell / verticesOnCell / cellsOnCell counter-clockwise, vertex i between edges i and i+1;
  * verticesOnEdge ordered along t = k x n;  edgesOnVertex / cellsOnVertex / kiteAreasOnVertex per vertex;
  * edgesOnEdge[:, e] = the other edges of cell 1 counter-clockwise starting after e, then the same for cell 2, and
    weightsOnEdge = sigma (1/2 - R) n_{e',c} dvEdge[e'] / dcEdge[e] with R the running kite-area fraction (TRiSK,
    Thuburn et al. 2009 / Ringler et al. 2010: the recipe of SURVEY.md Appendix B).

Host-side tool (numpy + scipy), meant for meshes of up to ~10^4 cells.
"""
from __future__ import annotations

import numpy as np

SQRT3 = float(np.sqrt(3.0))


def _area(poly: np.ndarray) -> float:
    x, y = poly[:, 0], poly[:, 1]
    return 0.5 * float(np.sum(x * np.roll(y, -1) - np.roll(x, -1) * y))


def _lloyd(pts: np.ndarray, Lx: float, Ly: float, iterations: int) -> np.ndarray:
    """Move every generator to the centroid of its (periodic) Voronoi cell, `iterations` times: towards a centroidal
    tessellation, which is what MPAS meshes are (SCVT) -- well-shaped cells, the topological defects of the start survive."""
    from scipy.spatial import Voronoi
    N = len(pts)
    for _ in range(iterations):
        allp = np.concatenate([pts + np.array([ox * Lx, oy * Ly]) for oy in (-1, 0, 1) for ox in (-1, 0, 1)])
        vor = Voronoi(allp)
        new = np.empty_like(pts)
        for c in range(N):
            poly = vor.vertices[vor.regions[vor.point_region[4 * N + c]]]
            ctr = poly.mean(axis=0)
            poly = poly[np.argsort(np.arctan2(poly[:, 1] - ctr[1], poly[:, 0] - ctr[0]))]
            x, y = poly[:, 0], poly[:, 1]
            cr = x * np.roll(y, -1) - np.roll(x, -1) * y
            a = 0.5 * cr.sum()
            new[c] = (((x + np.roll(x, -1)) * cr).sum() / (6.0 * a), ((y + np.roll(y, -1)) * cr).sum() / (6.0 * a))
        pts = new
    return pts


def _generators(nx: int, ny: int, dc: float, jitter: float, seed: int, lloyd: int):
    if ny % 2 or nx < 4 or ny < 4:
        raise ValueError("periodic_voronoi: nx, ny >= 4 and ny even")
    N = nx * ny
    Lx, Ly = nx * dc, ny * dc * SQRT3 / 2.0
    j, i = np.divmod(np.arange(N), nx)
    base = np.stack([dc * (i + 0.5 * (j & 1)) + 0.5 * dc, (j + 1) * (dc * SQRT3 / 2.0)], axis=1)
    rng = np.random.default_rng(seed)
    r, a = jitter * dc * np.sqrt(rng.random(N)), 2.0 * np.pi * rng.random(N)
    pts = base + np.stack([r * np.cos(a), r * np.sin(a)], axis=1)
    if lloyd:
        pts = _lloyd(pts, Lx, Ly, lloyd)
    return pts, Lx, Ly


def periodic_voronoi_loops(nx: int, ny: int, dc: float, jitter: float = 0.25, seed: int = 0, f0: float = 1.0e-4,
                           resting_thickness: float = 1000.0, lloyd: int = 0) -> dict:
    """The same mesh as `periodic_voronoi`, built entity by entity in Python loops from scipy's Voronoi diagram of a 3x3
    tiling: slow, obvious, independent of the vectorised construction below -- the tests hold the two against each other.
    (Edge / vertex numbering and edge orientation differ between the two; cells are numbered alike.)"""
    from scipy.spatial import Voronoi
    pts, Lx, Ly = _generators(nx, ny, dc, jitter, seed, lloyd)
    N = nx * ny
    tiles = [(ox, oy) for oy in (-1, 0, 1) for ox in (-1, 0, 1)]
    allp = np.concatenate([pts + np.array([ox * Lx, oy * Ly]) for ox, oy in tiles])
    vor = Voronoi(allp)
    central = lambda p: 4 * N <= p < 5 * N                                   # noqa: E731

    # ---- edges: every ridge that touches a central cell, once ----------------------------------------------------
    ridges = []
    for (p, q), rv in zip(vor.ridge_points, vor.ridge_vertices):
        if central(q) and not central(p):
            p, q = q, p
        if not central(p):
            continue
        if central(q):
            if q < p:
                p, q = q, p
        elif p % N > q % N:          # the mirrored copy (q central, an image of p outside) is the one that is kept
            continue
        if rv[0] < 0 or rv[1] < 0:
            raise RuntimeError("periodic_voronoi: open ridge inside the central tile")
        ridges.append((int(p), int(q), int(rv[0]), int(rv[1])))
    nE = len(ridges)
    # the three cells around every Voronoi vertex of the tiling identify it across the periodic copies
    cells_of_vertex: dict = {}
    for (p, q), rv in zip(vor.ridge_points, vor.ridge_vertices):
        for v in rv:
            if v >= 0:
                cells_of_vertex.setdefault(int(v), set()).update((int(p), int(q)))
    vid: dict = {}

    def vertex_id(v: int) -> int:
        key = tuple(sorted(p % N for p in cells_of_vertex[v]))
        if len(key) != 3 or len(set(key)) != 3:
            raise RuntimeError("periodic_voronoi: degenerate vertex (more than three cells, or a cell meeting its own image)")
        return vid.setdefault(key, len(vid))

    coe = np.zeros((nE, 2), np.int32)
    voe = np.zeros((nE, 2), np.int32)
    xE, yE, ang, dcE, dvE = (np.zeros(nE) for _ in range(5))
    # per cell: (angle of the edge seen from the centre, edge, is this cell c1?, the two vertex positions relative to the centre)
    per_cell: list = [[] for _ in range(N)]
    vpos: dict = {}
    for e, (p, q, va, vb) in enumerate(ridges):
        x1, x2 = allp[p], allp[q]
        d = x2 - x1
        n = d / np.linalg.norm(d)
        t = np.array([-n[1], n[0]])
        pa, pb = vor.vertices[va], vor.vertices[vb]
        if np.dot(pb - pa, t) < 0:
            va, vb, pa, pb = vb, va, pb, pa
        coe[e] = (p % N + 1, q % N + 1)
        ia, ib = vertex_id(va), vertex_id(vb)
        voe[e] = (ia + 1, ib + 1)
        mid = 0.5 * (x1 + x2)
        xE[e], yE[e] = mid[0] % Lx, mid[1] % Ly
        ang[e], dcE[e], dvE[e] = np.arctan2(d[1], d[0]), np.linalg.norm(d), np.linalg.norm(pb - pa)
        vpos.setdefault(ia, np.array([pa[0] % Lx, pa[1] % Ly]))
        vpos.setdefault(ib, np.array([pb[0] % Lx, pb[1] % Ly]))
        per_cell[p % N].append((np.arctan2(d[1], d[0]), e, True, pa - x1, pb - x1, ia, ib, 0.5 * d))
        per_cell[q % N].append((np.arctan2(-d[1], -d[0]), e, False, pa - x2, pb - x2, ia, ib, -0.5 * d))
    nV = len(vid)
    if nV - nE + N != 0:
        raise RuntimeError(f"periodic_voronoi: Euler characteristic of the torus violated (V - E + F = {nV - nE + N})")

    S = max(len(c) for c in per_cell)
    eoc, voc, coc = (np.zeros((N, S), np.int32) for _ in range(3))
    nEoC = np.zeros(N, np.int32)
    area = np.zeros(N)
    kite_cv: dict = {}                       # (cell, vertex) -> kite area
    for c in range(N):
        lst = sorted(per_cell[c], key=lambda z: z[0])          # counter-clockwise by the angle of the edge midpoint
        n = len(lst)
        nEoC[c] = n
        for k, (_, e, first, ra, rb, ia, ib, rm) in enumerate(lst):
            eoc[c, k] = e + 1
            coc[c, k] = coe[e, 1] if first else coe[e, 0]
            # the counter-clockwise end of this edge as seen from the cell: vertex b on the c1 side (t = k x n), a on the c2 side
            v_ccw, r_ccw = (ib, rb) if first else (ia, ra)
            voc[c, k] = v_ccw + 1
            rm_next = lst[(k + 1) % n][7]
            kite = _area(np.array([[0.0, 0.0], rm, r_ccw, rm_next]))
            kite_cv[(c, v_ccw)] = kite
            area[c] += kite
        ends = [((ia, ib) if first else (ib, ia)) for _, _, first, _, _, ia, ib, _ in lst]      # (clockwise end, counter-clockwise end)
        if any(ends[k][1] != ends[(k + 1) % n][0] for k in range(n)):
            raise RuntimeError("periodic_voronoi: the edges of a cell do not close counter-clockwise")
    if min(kite_cv.values()) <= 0.0:
        raise RuntimeError("periodic_voronoi: non-positive kite area (an obtuse Delaunay triangle: reduce the jitter)")

    # ---- dual mesh ---------------------------------------------------------------------------------------------------
    eov_l: list = [[] for _ in range(nV)]
    for e in range(nE):
        eov_l[voe[e, 0] - 1].append(e)
        eov_l[voe[e, 1] - 1].append(e)
    eov = np.array(eov_l, np.int32) + 1
    cov = np.zeros((nV, 3), np.int32)
    kites = np.zeros((nV, 3))
    for key, v in vid.items():
        cov[v] = np.array(key) + 1
        kites[v] = [kite_cv[(c, v)] for c in key]
    area_tri = kites.sum(axis=1)

    # ---- TRiSK edgesOnEdge / weightsOnEdge -------------------------------------------------------------------------
    S2 = 2 * S - 2
    eoe = np.zeros((nE, S2), np.int32)
    woe = np.zeros((nE, S2))
    nEoE = np.zeros(nE, np.int32)
    pos = {(c, int(eoc[c, k]) - 1): k for c in range(N) for k in range(nEoC[c])}
    for e in range(nE):
        slot = 0
        for side, sigma in ((0, 1.0), (1, -1.0)):
            c = int(coe[e, side]) - 1
            n, j0, rsum = int(nEoC[c]), pos[(c, e)], 0.0
            for k in range(1, n):
                rsum += kite_cv[(c, int(voc[c, (j0 + k - 1) % n]) - 1)] / area[c]     # the vertex passed on the way
                e2 = int(eoc[c, (j0 + k) % n]) - 1
                owner = 1.0 if coe[e2, 0] - 1 == c else -1.0
                eoe[e, slot] = e2 + 1
                woe[e, slot] = sigma * (0.5 - rsum) * owner * dvE[e2] / dcE[e]
                slot += 1
        nEoE[e] = slot

    m: dict = {"nCells": N, "nEdges": nE, "nVertices": nV, "maxEdges": S, "maxEdges2": S2, "vertexDegree": 3, "nVertLevels": 1,
               "is_periodic": "YES", "x_period": Lx, "y_period": Ly, "dc": float(dc), "nx": nx, "ny": ny}
    m["xCell"], m["yCell"], m["zCell"] = pts[:, 0] % Lx, pts[:, 1] % Ly, np.zeros(N)
    m["fCell"], m["areaCell"], m["nEdgesOnCell"] = np.full(N, f0), area, nEoC
    m["cellsOnEdge"], m["verticesOnEdge"], m["angleEdge"] = coe, voe, ang
    m["xEdge"], m["yEdge"], m["zEdge"], m["fEdge"] = xE, yE, np.zeros(nE), np.full(nE, f0)
    m["dcEdge"], m["dvEdge"] = dcE, dvE
    m["edgesOnCell"], m["cellsOnCell"], m["verticesOnCell"] = eoc, coc, voc
    m["edgesOnEdge"], m["weightsOnEdge"], m["nEdgesOnEdge"] = eoe, woe, nEoE
    vx = np.array([vpos[v] for v in range(nV)])
    m["xVertex"], m["yVertex"], m["zVertex"], m["fVertex"] = vx[:, 0], vx[:, 1], np.zeros(nV), np.full(nV, f0)
    m["areaTriangle"], m["edgesOnVertex"], m["cellsOnVertex"], m["kiteAreasOnVertex"] = area_tri, eov, cov, kites
    m["minLevelCell"] = np.ones(N, np.int32)
    m["maxLevelCell"] = np.ones(N, np.int32)
    m["restingThickness"] = np.full((N, 1), float(resting_thickness))
    m["boundaryEdge"] = np.zeros(nE, np.int32)
    return m


def periodic_voronoi(nx: int, ny: int, dc: float, jitter: float = 0.25, seed: int = 0, f0: float = 1.0e-4,
                     resting_thickness: float = 1000.0, lloyd: int = 0, allow_obtuse: bool = False, with_dual: bool = True) -> dict:
    """`nx` x `ny` cells (ny even, both >= 4) with mean spacing `dc`; `jitter` = maximal displacement of a centre from the
    hex lattice in units of dc (0.3 gives a few per cent of pentagons and heptagons; 0 reproduces the hexagons); `lloyd`
    relaxation sweeps afterwards (large jitter + a few sweeps: many defects, well-shaped cells).  `allow_obtuse`: keep
    meshes in which a Delaunay triangle is obtuse (its circumcentre falls outside, a kite area turns negative -- the
    identities the weights rest on still hold, the cells are just badly shaped); needed for large un-relaxed meshes.

    Vectorised: the Voronoi diagram is read off the Delaunay triangulation of the generators plus a margin of periodic
    images (a triangle = a vertex at its circumcentre, a triangle side = an edge), every array is built with sorts and
    segment operations -- a million cells in about a minute."""
    from scipy.spatial import Delaunay
    pts, Lx, Ly = _generators(nx, ny, dc, jitter, seed, lloyd)
    N = nx * ny
    # ---- generators + the periodic images within a margin of the tile ---------------------------------------------------
    w = 3.0 * dc
    P, B = [pts], [np.arange(N)]
    for oy in (-1, 0, 1):
        for ox in (-1, 0, 1):
            if ox == 0 and oy == 0:
                continue
            q = pts + np.array([ox * Lx, oy * Ly])
            keep = (q[:, 0] > -w) & (q[:, 0] < Lx + w) & (q[:, 1] > -w) & (q[:, 1] < Ly + w)
            P.append(q[keep])
            B.append(np.nonzero(keep)[0])
    allp, base = np.concatenate(P), np.concatenate(B)
    tri = Delaunay(allp).simplices.astype(np.int64)
    a, b, c = allp[tri[:, 0]], allp[tri[:, 1]], allp[tri[:, 2]]
    flip = (b[:, 0] - a[:, 0]) * (c[:, 1] - a[:, 1]) - (b[:, 1] - a[:, 1]) * (c[:, 0] - a[:, 0]) < 0
    tri[flip] = tri[flip][:, [0, 2, 1]]                                            # counter-clockwise
    tri = tri[(tri < N).any(axis=1)]                                                # touching the tile
    a, b, c = allp[tri[:, 0]], allp[tri[:, 1]], allp[tri[:, 2]]
    # circumcentre relative to corner a
    ab, ac = b - a, c - a
    d2 = 2.0 * (ab[:, 0] * ac[:, 1] - ab[:, 1] * ac[:, 0])
    lab, lac = (ab * ab).sum(axis=1), (ac * ac).sum(axis=1)
    cc = np.stack([(ac[:, 1] * lab - ab[:, 1] * lac) / d2, (ab[:, 0] * lac - ac[:, 0] * lab) / d2], axis=1)
    # ---- vertices: triangles up to periodic images (the sorted triple of cells identifies one) ------------------------
    tb = base[tri]
    key = np.sort(tb, axis=1)
    if (key[:, 0] == key[:, 1]).any() or (key[:, 1] == key[:, 2]).any():
        raise RuntimeError("periodic_voronoi: a cell meets its own image (mesh too small)")
    k1 = (key[:, 0] * N + key[:, 1]) * N + key[:, 2]
    uk, first, vid_of_tri = np.unique(k1, return_index=True, return_inverse=True)
    nV = len(uk)
    # ---- directed sides: a -> b of a counter-clockwise triangle has that triangle's vertex on its LEFT -------------------
    da = np.concatenate([tb[:, 0], tb[:, 1], tb[:, 2]])
    db = np.concatenate([tb[:, 1], tb[:, 2], tb[:, 0]])
    disp = np.concatenate([b - a, c - b, a - c])                                    # x(b image) - x(a)
    vleft = np.concatenate([vid_of_tri] * 3)
    rleft = np.concatenate([cc, cc - ab, cc - ac])                                  # the vertex relative to the side's start cell
    _, fi = np.unique(da * N + db, return_index=True)                               # once per physical directed side
    da, db, disp, vleft, rleft = da[fi], db[fi], disp[fi], vleft[fi], rleft[fi]
    # sort the sides by (start cell, angle): the neighbours of every cell counter-clockwise
    ang = np.arctan2(disp[:, 1], disp[:, 0])
    o = np.lexsort((ang, da))
    da, db, disp, vleft, rleft, ang = da[o], db[o], disp[o], vleft[o], rleft[o], ang[o]
    nD = len(da)
    nEoC = np.bincount(da, minlength=N).astype(np.int32)
    start = np.concatenate([[0], np.cumsum(nEoC)]).astype(np.int64)
    if nEoC.min() < 3:
        raise RuntimeError("periodic_voronoi: a cell with fewer than three sides")
    posn = np.arange(nD) - start[da]                                                # position of the side in its cell's row
    nxt = start[da] + (posn + 1) % nEoC[da]                                         # the next side counter-clockwise
    # ---- edges: undirected sides, cellsOnEdge = (smaller id, larger id) --------------------------------------------------
    lo, hi = np.minimum(da, db), np.maximum(da, db)
    ekeys, eid = np.unique(lo * N + hi, return_inverse=True)                        # edge of every directed side
    nE = len(ekeys)
    if nV - nE + N != 0:
        raise RuntimeError(f"periodic_voronoi: Euler characteristic of the torus violated (V - E + F = {nV - nE + N})")
    fwd = da < db                                                                   # the side runs c1 -> c2
    e_of_fwd = eid[fwd]
    coe = np.zeros((nE, 2), np.int32)
    coe[e_of_fwd, 0], coe[e_of_fwd, 1] = da[fwd] + 1, db[fwd] + 1
    dvec = np.zeros((nE, 2))
    dvec[e_of_fwd] = disp[fwd]
    dcE = np.hypot(dvec[:, 0], dvec[:, 1])
    angE = np.arctan2(dvec[:, 1], dvec[:, 0])
    mid = pts[coe[:, 0] - 1] + 0.5 * dvec
    # verticesOnEdge along t = k x n: from the vertex on the right of c1 -> c2 (= left of c2 -> c1) to the one on its left
    voe = np.zeros((nE, 2), np.int32)
    voe[e_of_fwd, 1] = vleft[fwd] + 1
    voe[eid[~fwd], 0] = vleft[~fwd] + 1
    vL, vR = np.zeros((nE, 2)), np.zeros((nE, 2))
    vL[e_of_fwd] = rleft[fwd]                                                       # relative to c1
    # the right vertex seen from c1: the left vertex of c2 -> c1 is given relative to c2; c2 sits at c1 + dvec
    vR[eid[~fwd]] = rleft[~fwd] + dvec[eid[~fwd]]
    dvE = np.hypot(vL[:, 0] - vR[:, 0], vL[:, 1] - vR[:, 1])
    # ---- per cell rows ---------------------------------------------------------------------------------------------------
    S = int(nEoC.max())
    eoc, coc, voc = (np.zeros((N, S), np.int32) for _ in range(3))
    eoc[da, posn], coc[da, posn], voc[da, posn] = eid + 1, db + 1, vleft + 1
    # kite of (cell, vertex to the left of side k): (centre, midpoint of side k, vertex, midpoint of side k + 1)
    m0, m1, rv = 0.5 * disp, 0.5 * disp[nxt], rleft
    kite = 0.5 * ((m0[:, 0] * rv[:, 1] - rv[:, 0] * m0[:, 1]) + (rv[:, 0] * m1[:, 1] - m1[:, 0] * rv[:, 1]))
    if not allow_obtuse and kite.min() <= 0.0:
        raise RuntimeError("periodic_voronoi: non-positive kite area (an obtuse Delaunay triangle: reduce the jitter, relax, or allow_obtuse)")
    area = np.bincount(da, weights=kite, minlength=N)
    # ---- TRiSK: for the edge of side (c, position k), walk c's other sides counter-clockwise ------------------------------
    S2 = 2 * S - 2
    eoe = np.zeros((nE, S2), np.int32)
    woe = np.zeros((nE, S2))
    n_c = nEoC[da].astype(np.int64)
    n1 = nEoC[coe[:, 0] - 1].astype(np.int64)
    slot0 = np.where(fwd, 0, n1[eid] - 1)                                           # cell 1's entries first, then cell 2's
    sigma = np.where(fwd, 1.0, -1.0)
    rsum = np.zeros(nD)
    for kk in range(1, S):
        act = kk < n_c
        passed = start[da] + (posn + kk - 1) % n_c                                  # the side whose left vertex is passed on the way
        rsum = rsum + kite[passed] / area[da]
        tgt = start[da] + (posn + kk) % n_c                                         # the side of the edge e'
        e2 = eid[tgt]
        owner = np.where(coe[e2, 0] - 1 == da, 1.0, -1.0)
        wv = sigma * (0.5 - rsum) * owner * dvE[e2] / dcE[eid]
        sl = slot0 + kk - 1
        eoe[eid[act], sl[act]] = e2[act] + 1
        woe[eid[act], sl[act]] = wv[act]
    nEoE = (n1 + nEoC[coe[:, 1] - 1] - 2).astype(np.int32)

    m: dict = {"nCells": N, "nEdges": nE, "nVertices": nV if with_dual else 0, "maxEdges": S, "maxEdges2": S2, "vertexDegree": 3,
               "nVertLevels": 1, "is_periodic": "YES", "x_period": Lx, "y_period": Ly, "dc": float(dc), "nx": nx, "ny": ny}
    m["xCell"], m["yCell"], m["zCell"] = pts[:, 0] % Lx, pts[:, 1] % Ly, np.zeros(N)
    m["fCell"], m["areaCell"], m["nEdgesOnCell"] = np.full(N, f0), area, nEoC
    m["cellsOnEdge"], m["angleEdge"] = coe, angE
    m["xEdge"], m["yEdge"], m["zEdge"], m["fEdge"] = mid[:, 0] % Lx, mid[:, 1] % Ly, np.zeros(nE), np.full(nE, f0)
    m["dcEdge"], m["dvEdge"] = dcE, dvE
    m["edgesOnCell"], m["cellsOnCell"] = eoc, coc
    m["edgesOnEdge"], m["weightsOnEdge"], m["nEdgesOnEdge"] = eoe, woe, nEoE
    if with_dual:
        m["verticesOnEdge"], m["verticesOnCell"] = voe, voc
        t0 = tri[first]                                                             # one triangle per vertex
        vx = allp[t0[:, 0]] + cc[first]
        m["xVertex"], m["yVertex"], m["zVertex"], m["fVertex"] = vx[:, 0] % Lx, vx[:, 1] % Ly, np.zeros(nV), np.full(nV, f0)
        cov = base[t0]                                                              # counter-clockwise
        m["cellsOnVertex"] = (cov + 1).astype(np.int32)
        # kiteAreasOnVertex[v, j] belongs to cellsOnVertex[v, j]; edgesOnVertex[v, j] joins cellsOnVertex[v, j] and [v, j + 1]
        kv = np.zeros((nV, 3))
        eov = np.zeros((nV, 3), np.int32)
        for j in range(3):
            sel = np.nonzero(cov[vleft, j] == da)[0]                                 # the side leaving corner j with v on its left
            kv[vleft[sel], j] = kite[sel]
            eov[vleft[sel], j] = eid[sel] + 1
        m["kiteAreasOnVertex"], m["edgesOnVertex"], m["areaTriangle"] = kv, eov, kv.sum(axis=1)
    m["minLevelCell"] = np.ones(N, np.int32)
    m["maxLevelCell"] = np.ones(N, np.int32)
    m["restingThickness"] = np.full((N, 1), float(resting_thickness))
    m["boundaryEdge"] = np.zeros(nE, np.int32)
    return m
