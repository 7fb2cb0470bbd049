"""CPU: pin the oracle against everything the reference's own tests hold for this path, and
against the project's known answers."""
import json
import os

import numpy as np
import pytest

import mesh_oracle
import moka_oracle as O
import moka_oracle_c as OC
from conftest import hex_mesh

HERE = os.path.dirname(os.path.abspath(__file__))
ATOL = 1e-8            # test/utilities.jl:10

# golden relative errors, reference test/ocn/test_Operators.jl:52-53, 72-73, 90-91 (L_inf, L_two)
GOLD = {"grad": (0.00125026071878552, 0.00134354611117257),
        "div": (0.00124886886594453, 0.00124886886590979),
        "curl": (0.16136566356969, 0.16134801689713)}


@pytest.fixture(scope="module")
def mesh48():
    m = mesh_oracle.build_periodic_hex(48, 48, 1000.0)
    mesh_oracle.sign_index_fields(m)
    return m


def _errors(m, grad, div, curl):
    f = O.planar_test_fields(m)
    out = {}
    out["grad"] = O.error_measures(grad, f["grad_h_edge"], m["dcEdge"] * m["dvEdge"] * 0.5)
    out["div"] = O.error_measures(div, f["div_F"], m["areaCell"])
    out["curl"] = O.error_measures(curl, f["curl_F"], m["areaTriangle"])
    return out


def test_reference_operator_goldens_numpy_oracle(mesh48):
    m = mesh48
    f = O.planar_test_fields(m)
    err = _errors(m, O.gradient_on_edge(m, f["h"]), O.divergence_on_cell(m, f["F_edge"])[0],
                  O.curl_on_vertex(m, f["F_edge"]))
    for k, (linf, ltwo) in GOLD.items():
        assert abs(err[k][1] - linf) < ATOL, (k, "Linf", err[k][1], linf)
        assert abs(err[k][0] - ltwo) < ATOL, (k, "L2", err[k][0], ltwo)


def test_reference_operator_goldens_product_generator_and_renumbering():
    """Same goldens from the product generator's mesh; norms are numbering independent."""
    m = hex_mesh(48, 48, 1000.0)
    f = O.planar_test_fields(m)
    err = _errors(m, O.gradient_on_edge(m, f["h"]), O.divergence_on_cell(m, f["F_edge"])[0],
                  O.curl_on_vertex(m, f["F_edge"]))
    for k, (linf, ltwo) in GOLD.items():
        assert abs(err[k][1] - linf) < ATOL and abs(err[k][0] - ltwo) < ATOL


def test_generators_agree(mesh48):
    mp = hex_mesh(48, 48, 1000.0)
    for k, a in mesh48.items():
        if not isinstance(a, np.ndarray) or k not in mp:
            continue
        b = mp[k]
        assert a.shape == b.shape, k
        if a.dtype.kind == "i":
            assert np.array_equal(a, b), k
        elif k == "weightsOnEdge":
            assert np.max(np.abs(a - b)) < 1e-11, k       # geometric kite areas vs closed form
        else:
            assert np.allclose(a, b, rtol=1e-13, atol=1e-7), k


def test_trisk_properties():
    m = hex_mesh(16, 16, 1000.0)
    w, eoe = m["weightsOnEdge"], m["edgesOnEdge"].astype(np.int64) - 1
    s3 = np.sqrt(3.0)
    vals = np.unique(np.round(np.abs(w) * s3, 12))
    assert np.allclose(vals, [0.0, 1 / 6, 1 / 3]), vals      # the two opposite-edge slots are zero
    # exact tangential reconstruction of a uniform flow: sum w*(U.n_eoe) = U.(k x n_e)
    U = np.array([0.3, -0.7])
    un = U[0] * np.cos(m["angleEdge"]) + U[1] * np.sin(m["angleEdge"])
    ut = -U[0] * np.sin(m["angleEdge"]) + U[1] * np.cos(m["angleEdge"])
    rec = np.sum(w * un[eoe], axis=1)
    assert np.max(np.abs(rec - ut)) < 1e-15
    # energy-neutral Coriolis: W + W^T = 0 for equal dc, dv
    nE = m["nEdges"]
    W = np.zeros((nE, nE))
    for i in range(10):
        W[np.arange(nE), eoe[:, i]] += w[:, i]
    assert np.max(np.abs(W + W.T)) < 1e-15
    # each edge carries opposite signs in its two cells
    tot = np.zeros(nE)
    np.add.at(tot, m["edgesOnCell"].astype(np.int64).ravel() - 1, m["edgeSignOnCell"].ravel())
    assert np.all(tot == 0)


def test_c_oracle_bit_exact_vs_numpy_oracle():
    m = hex_mesh(32)
    igw = O.InertialGravityWave(m)
    ssh, u, h = igw.initial_state()
    dt = O.reference_dt(m)
    for stepper in ("ForwardEuler", "RungeKutta4"):
        prog, diag = O.new_state(m, ssh, u, h), O.new_diag(m)
        om = OC.OracleModel(m, ssh, u, h)
        for _ in range(7):
            if stepper == "ForwardEuler":
                O.timestep_forward_euler(m, prog, diag, dt)
            else:
                O.timestep_rk4(m, prog, dt)
        om.run_loop(dt, 7, stepper)
        for k in ("ssh", "normalVelocity", "layerThickness"):
            assert np.array_equal(prog[k][-1], getattr(om, k)[1]), (stepper, k)
            assert np.array_equal(prog[k][0], getattr(om, k)[0]), (stepper, k, "prev")
        if stepper == "ForwardEuler":
            assert np.array_equal(diag["relativeVorticity"], om.relativeVorticity)
            assert np.array_equal(diag["velocityDivCell"], om.velocityDivCell)
            assert np.array_equal(diag["thicknessFlux"], om.thicknessFlux)
            assert om.sum_ssh2() == O.sum_array(prog["ssh"][-1])


def test_forward_euler_first_step_leaves_h_unchanged():
    """Q1: flux uses the hEdge of the previous call (zeros on step 1)."""
    m = hex_mesh(16)
    ssh, u, h = O.InertialGravityWave(m).initial_state()
    om = OC.OracleModel(m, ssh, u, h)
    om.run_loop(O.reference_dt(m), 1, "ForwardEuler")
    assert np.array_equal(om.layerThickness[1], h)
    assert not np.array_equal(om.normalVelocity[1], u)


@pytest.mark.parametrize("nx,nsteps,e_ssh,e_u", [(16, 11, 0.48754, 0.51372), (32, 23, 0.12497, 0.13389),
                                                  (64, 46, 0.031501, 0.033819)])
def test_rk4_igw_convergence_known_answers(nx, nsteps, e_ssh, e_u):
    """SURVEY.md Appendix B table (project-generated): IGW rel-L2 error at T = 10 h, 2nd order."""
    m = hex_mesh(nx)
    igw = O.InertialGravityWave(m)
    ssh, u, h = igw.initial_state()
    dc = 1.0e7 / nx
    assert round(36000 / (0.5 * dc / np.sqrt(O.GRAVITY * 1000.0))) == nsteps
    om = OC.OracleModel(m, ssh, u, h)
    om.run_loop(36000.0 / nsteps, nsteps, "RungeKutta4")
    assert abs(O.rel_l2(om.ssh[1], igw.exact_ssh(36000.0)) - e_ssh) < 5e-5 * max(1, e_ssh / 0.03)
    assert abs(O.rel_l2(om.normalVelocity[1], igw.exact_norm_vel(36000.0)) - e_u) < 5e-5 * max(1, e_u / 0.03)
    assert abs(np.sum(om.layerThickness[1] - h)) < 1e-9          # mass conserved to round-off


def test_mass_tendency_sums_to_zero():
    m = hex_mesh(16)
    ssh, u, h = O.InertialGravityWave(m).initial_state()
    _, th = O.tendencies_consistent(m, u, h)
    assert abs(np.sum(m["areaCell"] * th)) / np.sum(m["areaCell"] * np.abs(th)) < 1e-13


def test_committed_golden_fixture_matches_oracle():
    """tests/golden/igw16_*.npz were produced by tests/golden/make_golden.py from the C oracle."""
    path = os.path.join(HERE, "golden", "igw16_rk4_fe.npz")
    g = np.load(path)
    m = hex_mesh(16)
    ssh, u, h = O.InertialGravityWave(m).initial_state()
    meta = json.loads(str(g["meta"]))
    for stepper in ("ForwardEuler", "RungeKutta4"):
        om = OC.OracleModel(m, ssh, u, h)
        om.run_loop(meta["dt"], meta["nsteps"], stepper)
        assert np.array_equal(om.ssh[1], g[f"{stepper}_ssh"])
        assert np.array_equal(om.normalVelocity[1], g[f"{stepper}_normalVelocity"])
        assert np.array_equal(om.layerThickness[1], g[f"{stepper}_layerThickness"])
