"""CPU: the periodic Voronoi mesh generator (moka_b200/planar_voronoi.py) -- a genuinely unstructured MPAS C-grid with
pentagons, hexagons and heptagons -- checked through the identities a valid mesh and valid TRiSK weights satisfy, the
oracle stepping it sanely, and the domain decomposition of an irregular mesh pinned by the loop oracle."""
import numpy as np
import pytest

import moka_oracle as O
import moka_oracle_c as OC
import partition_oracle as PO
from moka_b200 import partition
from moka_b200.planar_hex import periodic_hex
from moka_b200.planar_voronoi import periodic_voronoi


@pytest.fixture(scope="module")
def vm():
    m = periodic_voronoi(20, 20, 1.0e7 / 20, jitter=0.3, seed=2)          # 12 pentagons, 376 hexagons, 12 heptagons
    OC.sign_index_fields(m)
    return m


def test_mesh_identities(vm):
    m = vm
    nC, nE, nV = m["nCells"], m["nEdges"], m["nVertices"]
    assert nV - nE + nC == 0                                                      # torus
    kinds = np.bincount(m["nEdgesOnCell"], minlength=9)
    assert kinds[5] > 0 and kinds[7] > 0 and kinds[:5].sum() == 0 and kinds[8:].sum() == 0 and kinds[5] == kinds[7]
    A = m["x_period"] * m["y_period"]
    assert abs(m["areaCell"].sum() - A) <= 1e-13 * A and abs(m["areaTriangle"].sum() - A) <= 1e-13 * A
    assert abs(m["kiteAreasOnVertex"].sum() - A) <= 1e-13 * A and m["kiteAreasOnVertex"].min() > 0
    coe, eoc, nec = m["cellsOnEdge"], m["edgesOnCell"], m["nEdgesOnCell"]
    listed = np.zeros(nE, int)
    for c in range(nC):
        e = eoc[c, :nec[c]] - 1
        assert np.all(eoc[c, nec[c]:] == 0) and np.all((coe[e] == c + 1).sum(axis=1) == 1)      # every edge of a cell lists that cell once
        listed[e] += 1
        assert np.all(m["cellsOnCell"][c, :nec[c]] == np.where(coe[e, 0] == c + 1, coe[e, 1], coe[e, 0]))
    assert np.all(listed == 2)
    assert np.all(m["nEdgesOnEdge"] == nec[coe[:, 0] - 1] + nec[coe[:, 1] - 1] - 2)
    # dv * dc / 2 summed over the edges tiles the plane as well (each edge's diamond)
    assert abs(0.5 * np.sum(m["dcEdge"] * m["dvEdge"]) - A) <= 1e-12 * A
    # edgeSignOnCell as the reference derives it (HorzMesh.jl:292-311): per edge, its two cells see opposite signs
    s = np.zeros(nE)
    for c in range(nC):
        np.add.at(s, eoc[c, :nec[c]] - 1, m["edgeSignOnCell"][c, :nec[c]])
    assert np.all(s == 0)


def test_trisk_weights_are_energy_neutral_and_reduce_to_the_hex_ones(vm):
    m = vm
    eoe, w, ne = m["edgesOnEdge"].astype(np.int64) - 1, m["weightsOnEdge"], m["nEdgesOnEdge"]
    W = {}
    for e in range(m["nEdges"]):
        for i in range(ne[e]):
            W[(e, eoe[e, i])] = W.get((e, eoe[e, i]), 0.0) + w[e, i] * m["dcEdge"][e] / m["dvEdge"][eoe[e, i]]
    assert max(abs(v + W.get((b, a), 0.0)) for (a, b), v in W.items()) < 1e-14          # w~[e,e'] = -w~[e',e]
    # without jitter the generator must reproduce the regular mesh's metrics and weights (up to numbering): exact
    # tangential reconstruction of a uniform flow is the numbering-independent statement of that
    m0 = periodic_voronoi(8, 8, 1000.0, jitter=0.0)
    h0 = periodic_hex(8, 8, 1000.0)
    assert np.allclose(np.sort(m0["dvEdge"]), np.sort(h0["dvEdge"]), rtol=1e-12) and np.allclose(m0["areaCell"], h0["areaCell"], rtol=1e-12)
    U = np.array([0.3, -0.8])
    un = np.cos(m0["angleEdge"]) * U[0] + np.sin(m0["angleEdge"]) * U[1]
    ut = -np.sin(m0["angleEdge"]) * U[0] + np.cos(m0["angleEdge"]) * U[1]
    e0 = m0["edgesOnEdge"].astype(np.int64) - 1
    rec = np.sum(m0["weightsOnEdge"] * un[e0], axis=1)
    assert np.max(np.abs(rec - ut)) < 1e-13
    assert np.allclose(np.sort(np.abs(m0["weightsOnEdge"]).ravel()), np.sort(np.abs(h0["weightsOnEdge"]).ravel()), atol=1e-13)


def test_oracle_on_the_voronoi_mesh_conserves_mass_and_nearly_energy(vm):
    m = vm
    ig = O.InertialGravityWave(m)
    ssh, u, h = ig.initial_state()
    dt = 0.25 * m["dc"] / float(np.sqrt(O.GRAVITY * 1000.0))
    om = OC.OracleModel(m, ssh, u, h)
    mass0 = float(np.sum(m["areaCell"] * h))

    def energy(uu, hh):
        c1, c2 = m["cellsOnEdge"][:, 0] - 1, m["cellsOnEdge"][:, 1] - 1
        ke = np.sum(0.5 * m["dcEdge"] * m["dvEdge"] * 0.5 * (hh[c1] + hh[c2]) * uu * uu)
        return float(ke + np.sum(m["areaCell"] * 0.5 * O.GRAVITY * (hh - 1000.0) ** 2))

    e0 = energy(u, h)
    om.run_loop(dt, 60, "RungeKutta4")
    un, hn = om.normalVelocity[1], om.layerThickness[1]
    assert abs(float(np.sum(m["areaCell"] * hn)) - mass0) <= 1e-13 * mass0
    assert abs(energy(un, hn) - e0) <= 2e-3 * e0 and np.abs(hn - 1000.0).max() < 3.0      # bounded: Coriolis does no work, RK4 barely damps
    # numpy and C oracles agree bit for bit on ragged rows too
    prog = O.new_state(m, ssh, u, h)
    for _ in range(3):
        O.timestep_rk4(m, prog, dt)
    oc = OC.OracleModel(m, ssh, u, h)
    oc.run_loop(dt, 3, "RungeKutta4")
    assert np.array_equal(prog["normalVelocity"][-1], oc.normalVelocity[1]) and np.array_equal(prog["layerThickness"][-1], oc.layerThickness[1])


@pytest.mark.parametrize("nparts", [2, 3, 8])
def test_partition_of_an_irregular_mesh_bit_exact_vs_loop_oracle(vm, nparts):
    m = vm
    part = partition.rcb_partition(m["xCell"], m["yCell"], nparts)
    part_o = PO.rcb_partition(m["xCell"].tolist(), m["yCell"].tolist(), nparts)
    assert part.tolist() == part_o
    locs = partition.decompose(m, nparts, part)
    mo = {k: (v.tolist() if isinstance(v, np.ndarray) else v) for k, v in m.items()
          if k in ("nCells", "cellsOnEdge", "edgesOnCell", "nEdgesOnCell")}
    sets, halos = PO.halo_lists(mo, part_o, nparts)
    owned_c, owned_e = np.zeros(m["nCells"], int), np.zeros(m["nEdges"], int)
    for r, loc in enumerate(locs):
        cells, nco, edges, neo = sets[r]
        assert loc["cellsGlobal"].tolist() == cells and loc["nCellsOwned"] == nco
        assert loc["edgesGlobal"].tolist() == edges and loc["nEdgesOwned"] == neo
        owned_c[loc["cellsGlobal"][:nco]] += 1
        owned_e[loc["edgesGlobal"][:neo]] += 1
        for q in loc["halo"]["peers"]:
            assert loc["halo"]["recv"][q].tolist() == halos[r]["recv"].get(q, [])
            assert loc["halo"]["send"][q].tolist() == halos[r]["send"].get(q, [])
        # the live part of every owned stencil row stays inside the local mesh
        no, ne = loc["nCellsOwned"], loc["nEdgesOwned"]
        for c in range(no):
            assert np.all(loc["edgesOnCell"][c, :loc["nEdgesOnCell"][c]] > 0)
        for e in range(ne):
            assert np.all(loc["edgesOnEdge"][e, :loc["nEdgesOnEdge"][e]] > 0) and np.all(loc["cellsOnEdge"][e] > 0)
    assert np.all(owned_c == 1) and np.all(owned_e == 1)


def test_lloyd_relaxed_mesh_keeps_its_defects_and_satisfies_the_identities():
    """Large jitter + Lloyd sweeps (towards a centroidal tessellation, which MPAS meshes are): many pentagons / heptagons,
    well-shaped cells, and the tangential reconstruction of a uniform flow becomes accurate to a few per cent."""
    m = periodic_voronoi(16, 16, 1000.0, jitter=0.4, seed=2, lloyd=6)
    kinds = np.bincount(m["nEdgesOnCell"], minlength=9)
    assert kinds[5] > 0 and kinds[5] == kinds[7] and kinds[:5].sum() == 0 and kinds[8:].sum() == 0
    A = m["x_period"] * m["y_period"]
    assert abs(m["areaCell"].sum() - A) <= 1e-13 * A and m["kiteAreasOnVertex"].min() > 0.2 * m["kiteAreasOnVertex"].mean()
    eoe, w = m["edgesOnEdge"].astype(np.int64) - 1, m["weightsOnEdge"]
    U = np.array([0.3, -0.8])
    un = np.cos(m["angleEdge"]) * U[0] + np.sin(m["angleEdge"]) * U[1]
    ut = -np.sin(m["angleEdge"]) * U[0] + np.cos(m["angleEdge"]) * U[1]
    rec = np.sum(np.where(eoe >= 0, w * un[np.maximum(eoe, 0)], 0.0), axis=1)
    assert np.sqrt(np.mean((rec - ut) ** 2)) < 0.05


def test_vectorised_and_loop_constructions_agree():
    """periodic_voronoi (Delaunay-based, vectorised) against periodic_voronoi_loops (entity by entity from scipy's Voronoi
    diagram).  Cells are numbered alike; edges are matched through their pair of cells (numbering and orientation differ)."""
    from moka_b200.planar_voronoi import periodic_voronoi_loops
    a = periodic_voronoi(14, 12, 1000.0, jitter=0.3, seed=5)
    b = periodic_voronoi_loops(14, 12, 1000.0, jitter=0.3, seed=5)
    assert (a["nCells"], a["nEdges"], a["nVertices"]) == (b["nCells"], b["nEdges"], b["nVertices"])
    assert np.array_equal(a["nEdgesOnCell"], b["nEdgesOnCell"]) and np.allclose(a["areaCell"], b["areaCell"], rtol=1e-12)
    assert np.allclose(a["xCell"], b["xCell"]) and np.allclose(a["yCell"], b["yCell"])
    for c in range(a["nCells"]):                                  # the same neighbours in the same cyclic order
        n = a["nEdgesOnCell"][c]
        ra, rb = a["cellsOnCell"][c, :n].tolist(), b["cellsOnCell"][c, :n].tolist()
        k = rb.index(ra[0])
        assert ra == rb[k:] + rb[:k]
    key = lambda m: {tuple(sorted(ce)): e for e, ce in enumerate(m["cellsOnEdge"].tolist())}      # noqa: E731
    ka, kb = key(a), key(b)
    assert ka.keys() == kb.keys()
    ea = np.array([ka[k] for k in ka])
    eb = np.array([kb[k] for k in ka])
    assert np.allclose(a["dcEdge"][ea], b["dcEdge"][eb], rtol=1e-12) and np.allclose(a["dvEdge"][ea], b["dvEdge"][eb], rtol=1e-9, atol=1e-9)
    assert np.array_equal(a["nEdgesOnEdge"][ea], b["nEdgesOnEdge"][eb])
    # weights: the same multiset of |w| per edge (orientation flips signs and swaps the two halves of a row)
    wa = np.sort(np.abs(a["weightsOnEdge"][ea]), axis=1)
    wb = np.sort(np.abs(b["weightsOnEdge"][eb]), axis=1)
    assert np.allclose(wa, wb, rtol=1e-9, atol=1e-12)
    assert np.allclose(np.sort(a["areaTriangle"]), np.sort(b["areaTriangle"]), rtol=1e-10)


def test_large_voronoi_mesh_with_obtuse_triangles_keeps_the_identities():
    m = periodic_voronoi(96, 96, 1000.0, jitter=0.3, seed=2, allow_obtuse=True)
    assert m["kiteAreasOnVertex"].min() < 0                                       # the un-relaxed mesh does have obtuse triangles at this size
    A = m["x_period"] * m["y_period"]
    assert abs(m["areaCell"].sum() - A) <= 1e-12 * A and m["areaCell"].min() > 0
    eoe, w, ne = m["edgesOnEdge"].astype(np.int64) - 1, m["weightsOnEdge"], m["nEdgesOnEdge"]
    nE = m["nEdges"]
    rows = np.repeat(np.arange(nE), eoe.shape[1]).reshape(eoe.shape)
    live = np.arange(eoe.shape[1])[None, :] < ne[:, None]
    wt = w * m["dcEdge"][:, None] / m["dvEdge"][np.maximum(eoe, 0)]
    import scipy.sparse as sp
    W = sp.coo_matrix((wt[live], (rows[live], eoe[live])), shape=(nE, nE)).tocsr()
    assert abs(W + W.T).max() < 1e-12                                             # energy-neutral Coriolis on any such mesh
