"""Domain decomposition of an MPAS mesh for one-process-per-GPU runs (host logic, numpy).

The reference has no multi-device code at all, so the algorithm is defined
here and pinned bit-exactly by an independent loop implementation in oracle/partition_oracle.py:

cell -> part   recursive coordinate bisection of the cell centres: split the longest extent among
               x, y (and z when the mesh is not planar; ties -> x, then y), order by (coordinate, global id),
               left gets floor(n*pl/p) cells with pl = p // 2 parts, recurse.
edge owner     owner of cellsOnEdge[1, e].
local sets     cells = [owned (by global id) | halo = not-owned cells sharing an edge with an owned
               cell (by global id)]; edges = [owned (by global id) | halo = every other edge of a
               local cell (by global id)].
dependencies   one RK stage for the owned edges/cells reads h on local cells and u on local edges
               only, so one exchange of (h on halo cells, u on halo edges)
               per stage suffices.
halo lists     recv[q] = halo entities owned by q in local order; send[q] = the entities rank q
               receives from this rank, in q's recv order.  Entities are addressed in one combined
               local index space [cells | edges] (edge k -> nCellsLocal + k), the order of the packed
               message is cells first, then edges.
"""
from __future__ import annotations

import numpy as np


def rcb_partition(x: np.ndarray, y: np.ndarray, nparts: int, z: np.ndarray | None = None) -> np.ndarray:
    """cell -> part (int32) by recursive coordinate bisection (`z`: spherical meshes -- without it the two hemispheres
    would be cut as one disc and every part would come in two far-apart pieces)."""
    n = x.shape[0]
    if not 1 <= int(nparts) <= n:
        raise ValueError(f"rcb_partition: nparts = {nparts} must be between 1 and the number of cells ({n}): every part owns at least one cell")
    part = np.zeros(n, np.int32)
    coords = [x, y] + ([z] if z is not None else [])

    def rec(ids: np.ndarray, p: int, base: int) -> None:
        if p == 1:
            part[ids] = base
            return
        vals = [c[ids] for c in coords]
        ext = [v.max() - v.min() for v in vals]
        axis_vals = vals[int(np.argmax(ext))]         # the first of equal extents: x before y before z
        order = np.lexsort((ids, axis_vals))          # by coordinate, ties by global id
        pl = p // 2
        nleft = (ids.shape[0] * pl) // p
        srt = ids[order]
        rec(srt[:nleft], pl, base)
        rec(srt[nleft:], p - pl, base + pl)

    rec(np.arange(n, dtype=np.int64), int(nparts), 0)
    return part


def build_local_mesh(m: dict, part: np.ndarray, rank: int) -> dict:
    """Local mesh (reference layouts, 1-based LOCAL ids, 0 = not local) + global id maps for `rank`."""
    nC, nE = m["nCells"], m["nEdges"]
    S, S2 = m["maxEdges"], m["maxEdges2"]
    coe = m["cellsOnEdge"].astype(np.int64) - 1            # (nE, 2), -1 absent
    eoc = m["edgesOnCell"].astype(np.int64) - 1            # (nC, S)
    nEoC = m["nEdgesOnCell"].astype(np.int64)
    slot_ok = np.arange(S)[None, :] < nEoC[:, None]

    owned_c = np.nonzero(part == rank)[0]                  # ascending global id
    # halo cells: the other cell of every edge of an owned cell
    e_of_owned = eoc[owned_c][slot_ok[owned_c]]
    nb = coe[e_of_owned].ravel()
    nb = nb[nb >= 0]
    halo_c = np.unique(nb[part[nb] != rank])
    cells = np.concatenate([owned_c, halo_c])
    # edges of all local cells; owned = cellsOnEdge[1] is an owned cell
    e_all = np.unique(eoc[cells][slot_ok[cells]])
    e_owner = part[coe[e_all, 0]]
    owned_e = e_all[e_owner == rank]
    halo_e = e_all[e_owner != rank]
    edges = np.concatenate([owned_e, halo_e])

    g2l_c = np.zeros(nC + 1, np.int32)                     # index 0 <- global -1 (absent); value 0 = not local
    g2l_c[cells + 1] = np.arange(1, cells.size + 1, dtype=np.int32)
    g2l_e = np.zeros(nE + 1, np.int32)
    g2l_e[edges + 1] = np.arange(1, edges.size + 1, dtype=np.int32)

    loc = {"nCells": int(cells.size), "nEdges": int(edges.size), "nVertices": 0,
           "nCellsOwned": int(owned_c.size), "nEdgesOwned": int(owned_e.size),
           "maxEdges": S, "maxEdges2": S2, "vertexDegree": m.get("vertexDegree", 3), "nVertLevels": int(m.get("nVertLevels", 1)),
           "dc": m.get("dc"), "x_period": m.get("x_period"), "y_period": m.get("y_period")}
    loc["cellsOnEdge"] = g2l_c[m["cellsOnEdge"][edges].astype(np.int64)]
    eoe_g = m["edgesOnEdge"][edges].astype(np.int64)
    loc["edgesOnEdge"] = g2l_e[eoe_g]
    loc["weightsOnEdge"] = np.ascontiguousarray(m["weightsOnEdge"][edges])
    loc["nEdgesOnEdge"] = np.ascontiguousarray(m["nEdgesOnEdge"][edges])
    loc["edgesOnCell"] = g2l_e[m["edgesOnCell"][cells].astype(np.int64)]
    loc["nEdgesOnCell"] = np.ascontiguousarray(m["nEdgesOnCell"][cells])
    for k in ("dcEdge", "dvEdge", "fEdge", "xEdge", "yEdge", "zEdge", "angleEdge", "boundaryEdge"):
        if m.get(k) is not None:
            loc[k] = np.ascontiguousarray(m[k][edges])
    for k in ("areaCell", "xCell", "yCell", "zCell", "fCell", "restingThickness"):
        if m.get(k) is not None:
            loc[k] = np.ascontiguousarray(m[k][cells])
    loc["cellsGlobal"] = cells.astype(np.int64)
    loc["edgesGlobal"] = edges.astype(np.int64)
    loc["cellOwner"] = part[cells].astype(np.int32)
    loc["edgeOwner"] = part[coe[edges, 0]].astype(np.int32)
    return loc


def recv_lists(loc: dict, nparts: int):
    """Per source rank q: (combined local indices, global cell ids, global edge ids) of the halo entities q owns."""
    nCl, nCo, nEo = loc["nCells"], loc["nCellsOwned"], loc["nEdgesOwned"]
    out = {}
    hc = np.arange(nCo, nCl)
    he = np.arange(nEo, loc["nEdges"])
    oc, oe = loc["cellOwner"][hc], loc["edgeOwner"][he]
    for q in range(nparts):
        c, e = hc[oc == q], he[oe == q]
        if c.size or e.size:
            out[q] = (np.concatenate([c, nCl + e]).astype(np.int32), loc["cellsGlobal"][c], loc["edgesGlobal"][e])
    return out


def decompose(m: dict, nparts: int, part: np.ndarray | None = None) -> list[dict]:
    """All local meshes with their halo send/recv lists (`halo` key) -- what rank 0 prepares."""
    if part is None:
        z = m.get("zCell")
        part = rcb_partition(m["xCell"], m["yCell"], nparts, z if z is not None and np.ptp(z) > 0 else None)
    locs = [build_local_mesh(m, part, r) for r in range(nparts)]
    recvs = [recv_lists(loc, nparts) for loc in locs]
    for r, loc in enumerate(locs):
        # global -> local lookup restricted to this rank's owned entities (sorted by global id)
        oc_g, oe_g = loc["cellsGlobal"][:loc["nCellsOwned"]], loc["edgesGlobal"][:loc["nEdgesOwned"]]
        send = {}
        for q in range(nparts):
            if q == r or r not in recvs[q]:
                continue
            _, gc, ge = recvs[q][r]
            lc = np.searchsorted(oc_g, gc)
            le = np.searchsorted(oe_g, ge)
            assert np.array_equal(oc_g[lc], gc) and np.array_equal(oe_g[le], ge)
            send[q] = np.concatenate([lc, loc["nCells"] + le]).astype(np.int32)
        recv = {q: v[0] for q, v in recvs[r].items()}
        peers = sorted(set(send) | set(recv))
        loc["halo"] = {
            "peers": peers,
            "send": {q: send.get(q, np.zeros(0, np.int32)) for q in peers},
            "recv": {q: recv.get(q, np.zeros(0, np.int32)) for q in peers},
        }
        loc["rank"], loc["nparts"] = r, nparts
    return locs


def _local_sets(m: dict, part: np.ndarray, rank: int):
    """(halo cells, halo edges) of `rank` as sorted global ids -- the index sets of build_local_mesh without its arrays."""
    S = m["maxEdges"]
    coe = m["cellsOnEdge"].astype(np.int64) - 1
    eoc = m["edgesOnCell"].astype(np.int64) - 1
    slot_ok = np.arange(S)[None, :] < m["nEdgesOnCell"].astype(np.int64)[:, None]
    owned_c = np.nonzero(part == rank)[0]
    nb = coe[eoc[owned_c][slot_ok[owned_c]]].ravel()
    nb = nb[nb >= 0]
    halo_c = np.unique(nb[part[nb] != rank])
    cells = np.concatenate([owned_c, halo_c])
    e_all = np.unique(eoc[cells][slot_ok[cells]])
    halo_e = e_all[part[coe[e_all, 0]] != rank]
    return halo_c, halo_e


def decompose_one(m: dict, nparts: int, rank: int, part: np.ndarray | None = None, loc: dict | None = None, halo_sets=None) -> dict:
    """decompose(m, nparts)[rank] without building the other ranks' meshes: what every rank of a job computes for itself
    from the (deterministic) global mesh (`loc`: its build_local_mesh result, when already at hand).  Only the other ranks' halo index sets are needed, to know what they expect from
    this one -- including the corner case of a rank that needs an edge of this rank without sharing a cell with it."""
    if part is None:
        z = m.get("zCell")
        part = rcb_partition(m["xCell"], m["yCell"], nparts, z if z is not None and np.ptp(z) > 0 else None)
    if loc is None:
        loc = build_local_mesh(m, part, rank)
    recv = {q: v[0] for q, v in recv_lists(loc, nparts).items()}
    coe0 = m["cellsOnEdge"][:, 0].astype(np.int64) - 1
    oc_g, oe_g = loc["cellsGlobal"][:loc["nCellsOwned"]], loc["edgesGlobal"][:loc["nEdgesOwned"]]
    send = {}
    for q in range(nparts):
        if q == rank:
            continue
        # (`halo_sets`: every rank's (halo cells, halo edges) as sorted global ids, e.g. all-gathered in a job where every rank
        #  has built its own local mesh -- otherwise recomputed here)
        hc, he = halo_sets[q] if halo_sets is not None else _local_sets(m, part, q)
        gc, ge = hc[part[hc] == rank], he[part[coe0[he]] == rank]      # what q holds as halo copies of this rank's entities, in q's recv order
        if gc.size or ge.size:
            lc, le = np.searchsorted(oc_g, gc), np.searchsorted(oe_g, ge)
            assert np.array_equal(oc_g[lc], gc) and np.array_equal(oe_g[le], ge)
            send[q] = np.concatenate([lc, loc["nCells"] + le]).astype(np.int32)
    peers = sorted(set(send) | set(recv))
    loc["halo"] = {"peers": peers, "send": {q: send.get(q, np.zeros(0, np.int32)) for q in peers},
                   "recv": {q: recv.get(q, np.zeros(0, np.int32)) for q in peers}}
    loc["rank"], loc["nparts"] = rank, nparts
    return loc


def flat_halo(loc: dict, nparts: int):
    """Flatten the per-peer lists into the all-to-all layout: (send_idx, send_counts, recv_idx, recv_counts)."""
    h = loc["halo"]
    z = np.zeros(0, np.int32)
    send = [h["send"].get(q, z) for q in range(nparts)]
    recv = [h["recv"].get(q, z) for q in range(nparts)]
    return (np.concatenate(send).astype(np.int32), [int(a.size) for a in send],
            np.concatenate(recv).astype(np.int32), [int(a.size) for a in recv])
