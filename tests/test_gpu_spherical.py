"""GPU: parity on a spherical Voronoi mesh with fEdge = 2 Omega sin(lat) -- the kind of mesh MPAS-Ocean runs on
(moka_b200/spherical_voronoi.py).  The Coriolis parameter varies, so the fused kernels take their "folded" form
(weightsOnEdge * fEdge[eoe] formed at upload: rel-L2 <= 1e-12 instead of bit-identity), the unfused
reference-order path stays bit-identical; coordinates are three-dimensional (Morton renumbering), cells mix pentagons,
hexagons and heptagons (the compile-time (12, 7) kernels), and the decomposition cuts a closed surface."""
import numpy as np
import pytest

import adjoint_oracle as A
import moka_b200 as mb
import moka_oracle_c as OC
from conftest import rel_l2
from moka_b200.spherical_voronoi import geostrophic_zonal_flow, spherical_voronoi

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def case():
    m = spherical_voronoi(1500)
    OC.sign_index_fields(m)
    ssh, u, h = geostrophic_zonal_flow(m)
    rng = np.random.default_rng(4)                                              # a disturbance on top, so that things move
    ssh = ssh + 2.0 * np.cos(3 * m["lonCell"]) * np.cos(m["latCell"]) ** 2
    u = u + 0.5 * rng.standard_normal(m["nEdges"])
    return m, ssh, u, 1000.0 + ssh, 0.25 * float(m["dcEdge"].min()) / float(np.sqrt(9.80616 * 1000.0))


@pytest.mark.parametrize("renumber", [True, False])
def test_rk4_and_forward_euler_on_the_sphere(backend, case, renumber):
    m, ssh, u, h, dt = case
    mesh = mb.Mesh(m, backend, renumber=renumber)
    om = OC.OracleModel(m, ssh, u, h)
    om.run_loop(dt, 20, "RungeKutta4")
    prog = mb.PrognosticVars(ssh, u, h, 2, mesh)
    mass0 = mb.reduce_sum(prog, "mass")
    mb.ocn_timestep(dt, prog, None, None, None, mb.RungeKutta4, nsteps=20)
    assert rel_l2(prog.normalVelocity, om.normalVelocity[1]) <= 1e-12 and rel_l2(prog.layerThickness, om.layerThickness[1]) <= 1e-12
    assert rel_l2(prog.ssh, om.ssh[1]) <= 1e-10                                  # ssh = h - 1000 carries the ulp of 1000
    assert abs(mb.reduce_sum(prog, "mass") - mass0) <= 1e-13 * mass0
    unf = mb.PrognosticVars(ssh, u, h, 2, mesh)
    mb.ocn_timestep(dt, unf, None, None, None, mb.RungeKutta4, nsteps=20, fused=False)
    assert np.array_equal(unf.normalVelocity, om.normalVelocity[1]) and np.array_equal(unf.layerThickness, om.layerThickness[1])
    exp = mb.PrognosticVars(ssh, u, h, 2, mb.Mesh(m, backend, renumber=renumber, explicit_eoe=True))
    mb.ocn_timestep(dt, exp, None, None, None, mb.RungeKutta4, nsteps=20)
    assert np.array_equal(exp.normalVelocity, prog.normalVelocity) and np.array_equal(exp.layerThickness, prog.layerThickness)
    pfe = mb.PrognosticVars(ssh, u, h, 2, mesh)
    mb.ocn_timestep(dt, pfe, None, None, None, mb.ForwardEuler, nsteps=10)
    ofe = OC.OracleModel(m, ssh, u, h)
    ofe.run_loop(dt, 10, "ForwardEuler")
    assert np.array_equal(pfe.normalVelocity, ofe.normalVelocity[1]) and np.array_equal(pfe.layerThickness, ofe.layerThickness[1])


def test_geostrophic_balance_holds_on_the_device(backend):
    m = spherical_voronoi(1500)
    ssh, u, h = geostrophic_zonal_flow(m)
    dt = 0.25 * float(m["dcEdge"].min()) / float(np.sqrt(9.80616 * 1000.0))
    prog = mb.PrognosticVars(ssh, u, h, 2, mb.Mesh(m, backend))
    mb.ocn_timestep(dt, prog, None, None, None, mb.RungeKutta4, nsteps=int(43200.0 / dt))
    assert np.abs(prog.layerThickness - h).max() < 0.01 * np.ptp(h) and np.abs(prog.normalVelocity - u).max() < 0.05 * 20.0


def test_both_adjoints_on_the_sphere(backend, case):
    m, ssh, u, h, dt = case
    mesh = mb.Mesh(m, backend)
    prog = mb.PrognosticVars(ssh, u, h, 2, mesh)
    d = mb.ocn_init_shadows(prog)
    mb.autodiff_reverse_run_loop(dt, prog, d, None, None, None, mb.RungeKutta4, 5)
    _, gu, gh = A.gradient_sum_ssh2(m, u, h, dt, 5)
    assert rel_l2(d.normalVelocity, gu) <= 1e-12 and rel_l2(d.layerThickness, gh) <= 1e-12
    prog = mb.PrognosticVars(ssh, u, h, 2, mesh)
    d = mb.ocn_init_shadows(prog)
    mb.autodiff_reverse_run_loop(dt, prog, d, None, None, None, mb.ForwardEuler, 5)
    _, gu, gh, gs, _ = A.gradient_sum_ssh2_fe(m, ssh, u, h, dt, 5)
    assert rel_l2(d.normalVelocity, gu) <= 1e-12 and rel_l2(d.layerThickness, gh) <= 1e-12 and rel_l2(d.ssh, gs) <= 1e-12


@pytest.mark.parametrize("nparts", [2, 8])
def test_decomposed_sphere_matches_the_single_domain_run(backend, case, nparts):
    from test_gpu_decomposed import _run_emulated
    m, ssh, u, h, dt = case
    md = {k: v for k, v in m.items() if k not in ("edgesOnVertex", "cellsOnVertex", "verticesOnEdge", "kiteAreasOnVertex",
                                                  "areaTriangle", "verticesOnCell", "edgeSignOnVertex")}
    md["nVertices"] = 0                                                            # decomposed meshes carry no vertex arrays
    gu, gh, gs, ranks = _run_emulated(backend, md, (ssh, u, h), nparts, dt, 8)
    prog = mb.PrognosticVars(ssh, u, h, 2, mb.Mesh(m, backend))
    mb.ocn_timestep(dt, prog, None, None, None, mb.RungeKutta4, nsteps=8)
    assert np.array_equal(gu, prog.normalVelocity) and np.array_equal(gh, prog.layerThickness)     # same kernel, same arithmetic per entity
    om = OC.OracleModel(m, ssh, u, h)
    om.run_loop(dt, 8, "RungeKutta4")
    assert rel_l2(gu, om.normalVelocity[1]) <= 1e-12 and rel_l2(gh, om.layerThickness[1]) <= 1e-12
