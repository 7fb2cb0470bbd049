"""CPU: the spherical Voronoi mesh generator (moka_b200/spherical_voronoi.py) -- quasi-uniform cells on the sphere, variable
Coriolis parameter: what MPAS-Ocean meshes are -- through the identities of a valid mesh, a physical known answer (a zonal wind
in geostrophic balance is a steady state of the equations the library integrates) and the partition oracle."""
import numpy as np
import pytest
import scipy.sparse as sp

import moka_oracle as O
import moka_oracle_c as OC
import partition_oracle as PO
from moka_b200 import partition
from moka_b200.spherical_voronoi import geostrophic_zonal_flow, spherical_voronoi


@pytest.fixture(scope="module")
def sm():
    m = spherical_voronoi(1500)
    OC.sign_index_fields(m)
    return m


def test_sphere_identities(sm):
    m = sm
    nC, nE, nV = m["nCells"], m["nEdges"], m["nVertices"]
    assert nV - nE + nC == 2 and nV == 2 * nC - 4 and nE == 3 * nC - 6
    kinds = np.bincount(m["nEdgesOnCell"], minlength=9)
    assert kinds[5] > 0 and kinds[7] > 0 and kinds[6] > kinds[5] + kinds[7] and kinds[8:].sum() == 0
    assert np.sum(6 - m["nEdgesOnCell"]) == 12                                   # Euler: the defects add up to twelve pentagons
    A = 4.0 * np.pi * m["sphere_radius"] ** 2
    assert abs(m["areaCell"].sum() - A) <= 1e-13 * A and abs(m["areaTriangle"].sum() - A) <= 1e-13 * A
    assert m["kiteAreasOnVertex"].min() > 0
    coe, eoc, nec = m["cellsOnEdge"], m["edgesOnCell"], m["nEdgesOnCell"]
    listed = np.zeros(nE, int)
    for c in range(nC):
        e = eoc[c, :nec[c]] - 1
        assert np.all((coe[e] == c + 1).sum(axis=1) == 1)
        listed[e] += 1
    assert np.all(listed == 2)
    assert np.all(m["nEdgesOnEdge"] == nec[coe[:, 0] - 1] + nec[coe[:, 1] - 1] - 2)
    assert np.ptp(m["fEdge"]) > 2.5e-4 and abs(m["fEdge"]).max() <= 2 * 7.292e-5     # 2 Omega sin(lat): far from uniform
    eoe, w, ne = m["edgesOnEdge"].astype(np.int64) - 1, m["weightsOnEdge"], m["nEdgesOnEdge"]
    rows = np.repeat(np.arange(nE), eoe.shape[1]).reshape(eoe.shape)
    live = np.arange(eoe.shape[1])[None, :] < ne[:, None]
    wt = w * m["dcEdge"][:, None] / m["dvEdge"][np.maximum(eoe, 0)]
    W = sp.coo_matrix((wt[live], (rows[live], eoe[live])), shape=(nE, nE)).tocsr()
    assert abs(W + W.T).max() < 1e-14                                             # TRiSK: the Coriolis term does no work


def test_geostrophic_zonal_flow_is_nearly_steady(sm):
    """u = u0 cos(lat), h = h0 - (R Omega u0 / g) sin^2(lat): tendencies vanish up to the discretisation error, and a day of
    RK4 leaves the state where it was -- this pins the overall sign and scale of weightsOnEdge, fEdge, the gradient and the
    divergence on a mesh whose every row is different."""
    m = sm
    ssh, u, h = geostrophic_zonal_flow(m)
    fu = float(np.abs(m["fEdge"]).max() * 20.0)
    tu = O.compute_normal_velocity_tendency(m, ssh, u)
    th = O.compute_layer_thickness_tendency(m, u * O.interpolate_cell2edge(m, h))
    assert np.sqrt(np.mean(tu ** 2)) < 0.01 * fu and np.abs(tu).max() < 0.05 * fu
    assert np.abs(th).max() < 1e-3 * np.ptp(h) / 86400.0 * 100                    # thickness changes by far less than its range per day
    # the same wind with the Coriolis sign flipped would be out of balance by 2 f u
    tu_wrong = O.compute_normal_velocity_tendency(m, -ssh, u)
    assert np.sqrt(np.mean(tu_wrong ** 2)) > 0.5 * np.sqrt(np.mean((m["fEdge"] * u) ** 2))
    dt = 0.25 * float(m["dcEdge"].min()) / float(np.sqrt(O.GRAVITY * 1000.0))
    om = OC.OracleModel(m, ssh, u, h)
    om.run_loop(dt, int(86400.0 / dt), "RungeKutta4")
    assert np.abs(om.layerThickness[1] - h).max() < 0.01 * np.ptp(h) and np.abs(om.normalVelocity[1] - u).max() < 0.05 * 20.0
    mass0 = float(np.sum(m["areaCell"] * h))
    assert abs(float(np.sum(m["areaCell"] * om.layerThickness[1])) - mass0) <= 1e-13 * mass0


@pytest.mark.parametrize("nparts", [2, 8])
def test_partition_of_the_sphere_bit_exact_vs_loop_oracle(sm, nparts):
    m = sm
    part = partition.rcb_partition(m["xCell"], m["yCell"], nparts, m["zCell"])
    part_o = PO.rcb_partition(m["xCell"].tolist(), m["yCell"].tolist(), nparts, m["zCell"].tolist())
    assert part.tolist() == part_o
    flat = partition.rcb_partition(m["xCell"], m["yCell"], nparts)              # cutting the projected disc instead
    locs = partition.decompose(m, nparts)                                        # decompose() takes z by itself on a sphere
    assert all(np.array_equal(loc["cellsGlobal"][:loc["nCellsOwned"]], np.nonzero(part == r)[0]) for r, loc in enumerate(locs))
    halo3 = sum(loc["nCells"] - loc["nCellsOwned"] for loc in locs)
    halo2 = sum(loc["nCells"] - loc["nCellsOwned"] for loc in partition.decompose(m, nparts, flat))
    assert halo3 <= halo2 and (nparts == 2 or halo3 < halo2)                      # compact parts: fewer halo cells (one cut is a plane either way)
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
    assert np.all(owned_c == 1) and np.all(owned_e == 1)
