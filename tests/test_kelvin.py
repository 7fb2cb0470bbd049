"""configs[4]: coastal Kelvin wave on a non-periodic (channel) hex mesh with boundary-edge masks.
The reference rejects non-periodic meshes (VertMesh.jl:50-52); the masked treatment is project-defined
 and identical in the oracle (mesh preprocessing) and the library (at upload)."""
import numpy as np
import pytest

import moka_b200 as mb
import moka_oracle as O
import moka_oracle_c as OC
from conftest import rel_l2


def _case(nx):
    m = mb.channel_hex(nx, nx, 1.0e7 / nx)
    mm = OC.apply_boundary_mask(m)
    OC.sign_index_fields(mm)
    return m, mm, mb.kelvinWave(m)


def test_channel_mesh_structure():
    m, mm, kw = _case(16)
    nC, nE = m["nCells"], m["nEdges"]
    assert nE == 3 * nC + int(m["boundaryEdge"].sum()) // 2
    coe, eoc = m["cellsOnEdge"], m["edgesOnCell"]
    assert np.all((coe[:, 1] == 0) == (m["boundaryEdge"] != 0))
    for c in range(nC):                                   # every edge of a cell lists that cell
        for i in range(6):
            assert c + 1 in coe[eoc[c, i] - 1]
    assert set(np.unique(m["nEdgesOnEdge"])) == {5, 10}
    # TRiSK: uniform flow along the wall is reconstructed exactly on interior edges away from the wall
    eoe, w = m["edgesOnEdge"].astype(np.int64) - 1, m["weightsOnEdge"]
    un = np.sin(m["angleEdge"])
    un[m["boundaryEdge"] != 0] = 0.0
    rec = np.sum(np.where(eoe >= 0, w * un[np.maximum(eoe, 0)], 0.0), axis=1)
    interior = (m["nEdgesOnEdge"] == 10) & np.all(m["boundaryEdge"][np.maximum(eoe, 0)] == 0, axis=1)
    assert np.max(np.abs(rec[interior] - np.cos(m["angleEdge"][interior]))) < 1e-14


def test_kelvin_oracle_converges_and_conserves():
    errs = []
    for nx in (32, 64):
        m, mm, kw = _case(nx)
        ssh, u, h = kw.initial_state()
        om = OC.OracleModel(mm, ssh, u, h)
        T = 10000.0
        n = int(np.ceil(T / mb.cfl_dt(m["dc"])))
        om.run_loop(T / n, n, "RungeKutta4")
        errs.append(O.rel_l2(om.ssh[1], kw.exact_ssh(T)))
        assert abs(np.sum(om.layerThickness[1] - h)) < 1e-9
        assert np.all(om.normalVelocity[1][m["boundaryEdge"] != 0] == 0.0)      # walls stay closed
    assert errs[0] < 0.15 and errs[1] < 0.6 * errs[0]                          # first order (staircase coast)


@pytest.mark.gpu
def test_kelvin_gpu_parity(backend):
    m, mm, kw = _case(64)
    ssh, u, h = kw.initial_state()
    dt = mb.cfl_dt(m["dc"])
    mesh = mb.Mesh(m, backend)
    # live reference path (ForwardEuler order) bit-exact
    prog = mb.PrognosticVars(ssh, u, h, 2, mesh)
    mb.ocn_run_loop(dt, prog, None, None, None, mb.ForwardEuler, 30)
    om = OC.OracleModel(mm, ssh, u, h)
    om.run_loop(dt, 30, "ForwardEuler")
    assert np.array_equal(prog.normalVelocity, om.normalVelocity[1]) and np.array_equal(prog.layerThickness, om.layerThickness[1])
    # fused RK4
    prog = mb.PrognosticVars(ssh, u, h, 2, mesh)
    mb.ocn_run_loop(dt, prog, None, None, None, mb.RungeKutta4, 60)
    om = OC.OracleModel(mm, ssh, u, h)
    om.run_loop(dt, 60, "RungeKutta4")
    assert rel_l2(prog.normalVelocity, om.normalVelocity[1]) <= 1e-12 and rel_l2(prog.layerThickness, om.layerThickness[1]) <= 1e-12
    # f-plane mesh: the fused kernel is bit-identical to the oracle (ssh = h - 1000 is too ill-conditioned here,
    # RMS ssh 0.16 m against ulp(1000 m) = 1.1e-13, for two different operation orders to agree to 1e-12)
    assert np.array_equal(prog.ssh, om.ssh[1]) and np.array_equal(prog.normalVelocity, om.normalVelocity[1])
    assert np.all(prog.normalVelocity[m["boundaryEdge"] != 0] == 0.0)
    assert abs(mb.reduce_sum(prog, "mass") - float(np.sum(m["areaCell"] * h))) <= 1e-13 * float(np.sum(m["areaCell"] * h))


@pytest.mark.gpu
def test_kelvin_decomposed_emulated(backend):
    from test_gpu_decomposed import _run_emulated
    m, mm, kw = _case(48)
    state = kw.initial_state()
    dt = mb.cfl_dt(m["dc"])
    gu, gh, gs, _ = _run_emulated(backend, m, state, 4, dt, 10)
    om = OC.OracleModel(mm, *state)
    om.run_loop(dt, 10, "RungeKutta4")
    assert rel_l2(gu, om.normalVelocity[1]) <= 1e-12 and rel_l2(gh, om.layerThickness[1]) <= 1e-12


@pytest.mark.gpu
def test_kelvin_1024_full_size(backend):
    """BASELINE.json configs[4] at its stated size: coastal Kelvin wave, 1024x1024 channel mesh with boundary-edge masks.
    Single-domain fused RK4 and ForwardEuler against the C oracle (bit for bit on this f-plane mesh, i.e. within the 1e-12 of
    the north star), walls closed, mass conserved."""
    m, mm, kw = _case(1024)
    ssh, u, h = kw.initial_state()
    dt = mb.cfl_dt(m["dc"])
    mesh = mb.Mesh(m, backend)
    prog = mb.PrognosticVars(ssh, u, h, 2, mesh)
    mb.ocn_run_loop(dt, prog, None, None, None, mb.RungeKutta4, 3)
    om = OC.OracleModel(mm, ssh, u, h)
    om.run_loop(dt, 3, "RungeKutta4")
    gu, gs = prog.normalVelocity, prog.ssh
    assert rel_l2(gu, om.normalVelocity[1]) <= 1e-12 and rel_l2(gs, om.ssh[1]) <= 1e-12
    assert np.array_equal(gu, om.normalVelocity[1]) and np.array_equal(prog.layerThickness, om.layerThickness[1])
    assert np.all(gu[m["boundaryEdge"] != 0] == 0.0)
    mass0 = float(np.sum(m["areaCell"] * h))
    assert abs(mb.reduce_sum(prog, "mass") - mass0) <= 1e-13 * mass0
    prog = mb.PrognosticVars(ssh, u, h, 2, mesh)
    mb.ocn_run_loop(dt, prog, None, None, None, mb.ForwardEuler, 3)
    om = OC.OracleModel(mm, ssh, u, h)
    om.run_loop(dt, 3, "ForwardEuler")
    assert np.array_equal(prog.normalVelocity, om.normalVelocity[1]) and np.array_equal(prog.ssh, om.ssh[1])


@pytest.mark.gpu
def test_kelvin_1024_full_size_eight_ranks(backend):
    """configs[4], "1 and 8 B200": the same mesh decomposed into 8 parts (ranks emulated on one GPU, messages moved by
    device copies), packed and direct-store halo exchange, against the single-domain C oracle."""
    from test_gpu_decomposed import _run_emulated
    m, mm, kw = _case(1024)
    state = kw.initial_state()
    dt = mb.cfl_dt(m["dc"])
    om = OC.OracleModel(mm, *state)
    om.run_loop(dt, 2, "RungeKutta4")
    for halo in ("nccl", "p2p_fused"):
        gu, gh, gs, ranks = _run_emulated(backend, m, state, 8, dt, 2, halo=halo)
        assert rel_l2(gu, om.normalVelocity[1]) <= 1e-12 and rel_l2(gs, om.ssh[1]) <= 1e-12, halo
        assert np.array_equal(gu, om.normalVelocity[1]) and np.array_equal(gh, om.layerThickness[1]), halo
        del ranks
