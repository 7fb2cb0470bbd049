"""GPU: parity on a genuinely unstructured mesh -- the periodic Voronoi diagram of a jittered lattice (pentagons, hexagons,
heptagons; every metric different; edgesOnEdge rows of 8 to 12 entries; moka_b200/planar_voronoi.py).  Every other parity
test runs on regular hexagons (or on hexagon rows made ragged artificially); this one takes the run-time-width kernels
through real irregular connectivity, the renumbering through non-lattice coordinates, the decomposition through an irregular
cell graph and both adjoints through an irregular transposed stencil.  Bit-exact against the oracle where the hexagon
tests are (same operation order), rel-L2 <= 1e-12 for the adjoints."""
import numpy as np
import pytest

import adjoint_oracle as A
import moka_b200 as mb
import moka_oracle_c as OC
from conftest import rel_l2
from moka_b200.planar_voronoi import periodic_voronoi

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def case():
    m = periodic_voronoi(24, 24, 1.0e7 / 24, jitter=0.3, seed=2)
    OC.sign_index_fields(m)
    kinds = np.bincount(m["nEdgesOnCell"], minlength=8)
    assert kinds[5] > 0 and kinds[7] > 0
    ssh, u, h = mb.inertialGravityWave(m).initial_state()
    return m, ssh, u, h, 0.25 * mb.cfl_dt(m["dc"])


@pytest.mark.parametrize("renumber", [True, False])
def test_rk4_and_forward_euler_bit_exact_on_a_voronoi_mesh(backend, case, renumber):
    m, ssh, u, h, dt = case
    mesh = mb.Mesh(m, backend, renumber=renumber)
    assert (mesh.maxEdges, mesh.maxEdges2) == (7, 12)             # the compile-time (12, 7) kernels; rows of 8..12 / 5..7 live entries
    nblk, nder = mesh.derived_blocks()
    assert nder == nblk > 0                                       # the generator follows the MPAS edgesOnEdge ordering: rebuilt everywhere
    prog = mb.PrognosticVars(ssh, u, h, 2, mesh)
    mass0 = mb.reduce_sum(prog, "mass")
    mb.ocn_timestep(dt, prog, None, None, None, mb.RungeKutta4, nsteps=25)
    om = OC.OracleModel(m, ssh, u, h)
    om.run_loop(dt, 25, "RungeKutta4")
    assert np.array_equal(prog.normalVelocity, om.normalVelocity[1]) and np.array_equal(prog.layerThickness, om.layerThickness[1])
    assert np.array_equal(prog.ssh, om.ssh[1])
    assert abs(mb.reduce_sum(prog, "mass") - mass0) <= 1e-13 * mass0
    unf = mb.PrognosticVars(ssh, u, h, 2, mesh)
    mb.ocn_timestep(dt, unf, None, None, None, mb.RungeKutta4, nsteps=25, fused=False)
    assert np.array_equal(unf.normalVelocity, prog.normalVelocity) and np.array_equal(unf.layerThickness, prog.layerThickness)
    # reading edgesOnEdge instead of rebuilding it, and the run-time-width kernels (wider device rows kept): the same bits
    for kw in (dict(explicit_eoe=True), dict(keep_widths=True)):
        mk = dict(m)
        if "keep_widths" in kw:
            from test_gpu_parity import _padded
            mk = _padded(m, S=8, S2=14)
        other = mb.Mesh(mk, backend, renumber=renumber, **kw)
        assert other.derived_blocks()[1] == 0
        p2 = mb.PrognosticVars(ssh, u, h, 2, other)
        mb.ocn_timestep(dt, p2, None, None, None, mb.RungeKutta4, nsteps=25)
        assert np.array_equal(p2.normalVelocity, prog.normalVelocity) and np.array_equal(p2.layerThickness, prog.layerThickness)
    pfe = mb.PrognosticVars(ssh, u, h, 2, mesh)
    diag, tend = mb.DiagnosticVars(pfe), mb.TendencyVars(pfe)
    mb.ocn_timestep(dt, pfe, diag, tend, None, mb.ForwardEuler, nsteps=12)
    ofe = OC.OracleModel(m, ssh, u, h)
    ofe.run_loop(dt, 12, "ForwardEuler")
    assert np.array_equal(pfe.normalVelocity, ofe.normalVelocity[1]) and np.array_equal(pfe.layerThickness, ofe.layerThickness[1])
    assert np.array_equal(diag.relativeVorticity, ofe.relativeVorticity[:m["nVertices"]]) and np.array_equal(tend.tendNormalVelocity, ofe.tendNormalVelocity)


def test_both_adjoints_on_a_voronoi_mesh(backend, case):
    m, ssh, u, h, dt = case
    mesh = mb.Mesh(m, backend)
    prog = mb.PrognosticVars(ssh, u, h, 2, mesh)
    d = mb.ocn_init_shadows(prog)
    mb.autodiff_reverse_run_loop(dt, prog, d, None, None, None, mb.RungeKutta4, 6)
    _, gu, gh = A.gradient_sum_ssh2(m, u, h, dt, 6)
    assert rel_l2(d.normalVelocity, gu) <= 1e-12 and rel_l2(d.layerThickness, gh) <= 1e-12
    prog = mb.PrognosticVars(ssh, u, h, 2, mesh)
    d = mb.ocn_init_shadows(prog)
    mb.autodiff_reverse_run_loop(dt, prog, d, None, None, None, mb.ForwardEuler, 6)
    _, gu, gh, gs, _ = A.gradient_sum_ssh2_fe(m, ssh, u, h, dt, 6)
    assert rel_l2(d.normalVelocity, gu) <= 1e-12 and rel_l2(d.layerThickness, gh) <= 1e-12 and rel_l2(d.ssh, gs) <= 1e-12
    k = int(np.argmax(m["nEdgesOnCell"] == 7))                    # a heptagon: finite-difference check of the reference's kind
    fd = A.finite_difference_fe(m, ssh, u, h, dt, 6, "h", k, eps=1e-7)
    assert abs(d.layerThickness[k] - fd) < 1e-4


@pytest.mark.parametrize("nparts", [3, 8])
def test_decomposed_voronoi_mesh_matches_the_single_domain_run(backend, case, nparts):
    from test_gpu_decomposed import _run_emulated
    m, ssh, u, h, dt = case
    gu, gh, gs, ranks = _run_emulated(backend, m, (ssh, u, h), nparts, dt, 9)
    om = OC.OracleModel(m, ssh, u, h)
    om.run_loop(dt, 9, "RungeKutta4")
    assert np.array_equal(gu, om.normalVelocity[1]) and np.array_equal(gh, om.layerThickness[1]) and np.array_equal(gs, om.ssh[1])
