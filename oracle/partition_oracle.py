"""ORACLE (test infrastructure, not product code): loop restatement of the domain decomposition.

Only tests/ may import this.  The reference has no partitioner (grep for
MPI|halo|partition hits only Project.toml:18-20), so the algorithm is project-defined
(moka_b200/partition.py docstring) and this file restates it independently with plain Python
loops, sets and `sorted`, so that cell->part, local<->global maps and the per-neighbour send/recv
lists can be compared bit for bit.  Small meshes only.
"""
from __future__ import annotations


def rcb_partition(x, y, nparts, z=None):
    n = len(x)
    part = [0] * n
    coords = [x, y] + ([z] if z is not None else [])

    def rec(ids, p, base):
        if p == 1:
            for i in ids:
                part[i] = base
            return
        best, best_ext = 0, None
        for k, c in enumerate(coords):                     # the longest extent, the earliest coordinate among equals
            vals = [c[i] for i in ids]
            ext = max(vals) - min(vals)
            if best_ext is None or ext > best_ext:
                best, best_ext = k, ext
        c = coords[best]
        srt = sorted(ids, key=lambda i: (c[i], i))
        pl = p // 2
        nleft = (len(ids) * pl) // p
        rec(srt[:nleft], pl, base)
        rec(srt[nleft:], p - pl, base + pl)

    rec(list(range(n)), nparts, 0)
    return part


def local_sets(m, part, rank):
    """(cells, nOwnedCells, edges, nOwnedEdges) as lists of global 0-based ids."""
    nC = m["nCells"]
    coe, eoc, nEoC = m["cellsOnEdge"], m["edgesOnCell"], m["nEdgesOnCell"]
    owned = [c for c in range(nC) if part[c] == rank]
    halo = set()
    for c in owned:
        for i in range(nEoC[c]):
            e = eoc[c][i] - 1
            for s in range(2):
                o = coe[e][s] - 1
                if o >= 0 and part[o] != rank:
                    halo.add(o)
    cells = owned + sorted(halo)
    eset = set()
    for c in cells:
        for i in range(nEoC[c]):
            eset.add(int(eoc[c][i]) - 1)
    owned_e = sorted(e for e in eset if part[coe[e][0] - 1] == rank)
    halo_e = sorted(e for e in eset if part[coe[e][0] - 1] != rank)
    return cells, len(owned), owned_e + halo_e, len(owned_e)


def halo_lists(m, part, nparts):
    """For every rank: {"send": {q: [combined local idx]}, "recv": {q: [...]}}, combined index space
    [cells | edges]; recv in local order, send in the receiver's recv order."""
    sets = [local_sets(m, part, r) for r in range(nparts)]
    coe = m["cellsOnEdge"]
    out = []
    for r in range(nparts):
        cells, nco, edges, neo = sets[r]
        recv = {}
        for k in range(nco, len(cells)):
            recv.setdefault(part[cells[k]], []).append(k)
        for k in range(neo, len(edges)):
            recv.setdefault(part[coe[edges[k]][0] - 1], []).append(len(cells) + k)
        out.append({"recv": {q: v for q, v in recv.items()}, "send": {}})
    for r in range(nparts):
        cells_r, _, edges_r, _ = sets[r]
        for q, lst in out[r]["recv"].items():
            cells_q, _, edges_q, _ = sets[q]
            pos_c = {g: i for i, g in enumerate(cells_q)}
            pos_e = {g: i for i, g in enumerate(edges_q)}
            send = []
            for k in lst:
                if k < len(cells_r):
                    send.append(pos_c[cells_r[k]])
                else:
                    send.append(len(cells_q) + pos_e[edges_r[k - len(cells_r)]])
            out[q]["send"][r] = send
    return sets, out
