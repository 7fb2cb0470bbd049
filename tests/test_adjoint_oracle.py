"""CPU: the adjoint oracle against the reference's own acceptance test for gradients (central finite
differences of J = sum ssh^2, test/enzyme/test_Enzyme_end2end.jl:112-180) and the dot-product identity."""
import numpy as np

import adjoint_oracle as A
import moka_b200 as mb
import moka_oracle as O
import moka_oracle_c as OC
from conftest import hex_mesh


def _case(nx=16, kelvin=False):
    if kelvin:
        m = OC.apply_boundary_mask(mb.channel_hex(nx, nx, 1.0e7 / nx))
        OC.sign_index_fields(m)
        ssh, u, h = mb.kelvinWave(m).initial_state()
    else:
        m = hex_mesh(nx)
        ssh, u, h = mb.inertialGravityWave(m).initial_state()
    return m, u, h, mb.cfl_dt(m["dc"])


def test_vjp_is_the_transpose_of_jvp():
    for kelvin in (False, True):
        m, u, h, dt = _case(12, kelvin)
        rng = np.random.default_rng(0)
        du, dh = rng.standard_normal(m["nEdges"]), rng.standard_normal(m["nCells"])
        wu, wh = rng.standard_normal(m["nEdges"]), rng.standard_normal(m["nCells"])
        tu, th = A.tendencies_jvp(m, u, h, du, dh)
        ub, hb = A.tendencies_vjp(m, u, h, wu, wh)
        lhs, rhs = tu @ wu + th @ wh, du @ ub + dh @ hb
        assert abs(lhs - rhs) <= 1e-12 * max(abs(lhs), abs(rhs))
        # jvp itself: F is quadratic, so a central difference is exact up to round-off
        e = 1e-3
        fp, fm = O.tendencies_consistent(m, u + e * du, h + e * dh), O.tendencies_consistent(m, u - e * du, h - e * dh)
        assert O.rel_l2((fp[0] - fm[0]) / (2 * e), tu) < 1e-9 and O.rel_l2((fp[1] - fm[1]) / (2 * e), th) < 1e-9


def test_gradient_matches_finite_differences_like_the_reference_test():
    m, u, h, dt = _case(16)
    J, gu, gh = A.gradient_sum_ssh2(m, u, h, dt, 5)
    assert J > 0
    for k in (4, 77, 200):                                               # the reference checks index 5 (1-based)
        fd_h = A.finite_difference(m, u, h, dt, 5, "h", k, eps=1e-7)
        fd_u = A.finite_difference(m, u, h, dt, 5, "u", k, eps=1e-4)
        assert abs(gh[k] - fd_h) < 1e-4                                  # atol of test_Enzyme_end2end.jl:176
        assert abs(gu[k] - fd_u) < 1e-2                                  # atol of :177
        assert abs(gh[k] - fd_h) < 1e-5 * abs(gh[k]) + 1e-7 and abs(gu[k] - fd_u) < 1e-5 * abs(gu[k]) + 1e-5


def test_step_vjp_dot_product_identity():
    m, u, h, dt = _case(12, kelvin=True)
    rng = np.random.default_rng(1)
    du, dh = 1e-3 * rng.standard_normal(m["nEdges"]), rng.standard_normal(m["nCells"])
    lu, lh = rng.standard_normal(m["nEdges"]), rng.standard_normal(m["nCells"])
    e = 1e-4
    up, hp = A.rk4_step(m, u + e * du, h + e * dh, dt)
    um, hm = A.rk4_step(m, u - e * du, h - e * dh, dt)
    lhs = ((up - um) / (2 * e)) @ lu + ((hp - hm) / (2 * e)) @ lh
    bu, bh = A.rk4_step_vjp(m, u, h, dt, lu, lh)
    rhs = du @ bu + dh @ bh
    assert abs(lhs - rhs) <= 1e-7 * max(abs(lhs), abs(rhs))


def _gather_form_step_vjp(m, u, h, dt, lam_u, lam_h):
    """numpy emulation of csrc/kernels_adjoint.cuh + adjoint_step (moka_b200.cu): the GATHER formulation with the
    transposed Coriolis stencil, q = invArea * kbar_h, and the fused RK bookkeeping, statement for statement."""
    nC, nE = m["nCells"], m["nEdges"]
    c1, c2 = m["cellsOnEdge"][:, 0] - 1, m["cellsOnEdge"][:, 1] - 1
    masked = c1 == c2
    gdc, dv, inv_area = O.GRAVITY * (1.0 / m["dcEdge"]), m["dvEdge"], 1.0 / m["areaCell"]
    eoe, w, ne, f = m["edgesOnEdge"], m["weightsOnEdge"], m["nEdgesOnEdge"], m["fEdge"]
    rows = [[] for _ in range(nE)]                                    # ensure_adjoint_mesh: row x <- (e, w[i,e]*f[x])
    for e in range(nE):
        for i in range(ne[e]):
            x = eoe[e, i] - 1
            if x >= 0:
                rows[x].append((e, w[e, i] * f[x]))
    eoc, n_eoc = m["edgesOnCell"], m["nEdgesOnCell"]
    a = [dt / 2.0, dt / 2.0, dt]
    b = [dt / 6.0, dt / 3.0, dt / 3.0, dt / 6.0]
    ys = A.rk4_stage_states(m, u, h, dt)
    acc_u = acc_h = ku = kq = None
    for s in (4, 3, 2, 1):
        uY, hY = ys[s - 1]
        if s == 4:
            ku, kq = b[3] * lam_u, b[3] * inv_area * lam_h
        G_e = dv * (np.where(masked, 0.0, kq[c2]) - kq[c1])
        yb_u = 0.5 * (hY[c1] + np.where(masked, hY[c1], hY[c2])) * G_e
        for x in range(nE):
            for e, wt in rows[x]:
                yb_u[x] += wt * ku[e]
        yb_h = np.zeros(nC)
        for c in range(nC):
            for i in range(n_eoc[c]):
                e = eoc[c, i] - 1
                sgn = -1.0 if c1[e] == c else 1.0
                other = c2[e] if c1[e] == c else c1[e]
                G = dv[e] * sgn * (kq[c] if masked[e] else kq[c] - kq[other])
                yb_h[c] += (1.0 if masked[e] else 0.5) * uY[e] * G
                if not masked[e]:
                    yb_h[c] -= sgn * gdc[e] * ku[e]
        acc_u = (lam_u if s == 4 else acc_u) + yb_u
        acc_h = (lam_h if s == 4 else acc_h) + yb_h
        if s > 1:
            ku = b[s - 2] * lam_u + a[s - 2] * yb_u
            kq = inv_area * (b[s - 2] * lam_h + a[s - 2] * yb_h)
    return acc_u, acc_h


def test_gather_form_of_the_cuda_adjoint_equals_the_scatter_oracle():
    for kelvin in (False, True):
        m, u, h, dt = _case(8, kelvin)
        rng = np.random.default_rng(2)
        lu, lh = rng.standard_normal(m["nEdges"]), rng.standard_normal(m["nCells"])
        gu, gh = _gather_form_step_vjp(m, u, h, dt, lu, lh)
        su, sh = A.rk4_step_vjp(m, u, h, dt, lu, lh)
        assert O.rel_l2(gu, su) < 1e-13 and O.rel_l2(gh, sh) < 1e-13


def test_committed_adjoint_fixture_is_reproduced_by_the_oracle():
    import json
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "igw16_adjoint.npz"))
    meta = json.loads(str(g["meta"]))
    m = hex_mesh(16)
    ssh, u, h = mb.inertialGravityWave(m).initial_state()
    J, gu, gh = A.gradient_sum_ssh2(m, u, h, meta["dt"], meta["nsteps"])
    assert J == float(g["J"]) and np.array_equal(gu, g["d_normalVelocity"]) and np.array_equal(gh, g["d_layerThickness"])
    k = meta["fd_index"]
    assert abs(gh[k] - float(g["fd_layerThickness"])) < 1e-4 and abs(gu[k] - float(g["fd_normalVelocity"])) < 1e-2


# ---- ForwardEuler adjoint oracle (the stepper test_Enzyme_end2end.jl differentiates) --------------------------------
def _fe_case(nx=16, kelvin=False):
    if kelvin:
        m = OC.apply_boundary_mask(mb.channel_hex(nx, nx, 1.0e7 / nx))
        OC.sign_index_fields(m)
        ssh, u, h = mb.kelvinWave(m).initial_state()
    else:
        m = hex_mesh(nx)
        ssh, u, h = mb.inertialGravityWave(m).initial_state()
    return m, ssh, u, h, mb.cfl_dt(m["dc"])


def test_fe_oracle_step_is_the_reference_order_forward_euler():
    m, ssh, u, h, dt = _fe_case(12)
    om = OC.OracleModel(m, ssh, u, h)
    om.run_loop(dt, 4, "ForwardEuler")
    u4, h4, s4, _ = A.run_forward_fe(m, ssh, u, h, dt, 4)[-1]
    assert np.array_equal(u4, om.normalVelocity[1]) and np.array_equal(h4, om.layerThickness[1]) and np.array_equal(s4, om.ssh[1])


def test_fe_step_vjp_dot_product_identity():
    for kelvin in (False, True):
        m, ssh, u, h, dt = _fe_case(12, kelvin)
        rng = np.random.default_rng(2)
        hE = O.interpolate_cell2edge(m, h)
        d = [1e-3 * rng.standard_normal(m["nEdges"]), rng.standard_normal(m["nCells"]), rng.standard_normal(m["nCells"]),
             rng.standard_normal(m["nEdges"])]
        lam = [rng.standard_normal(m["nEdges"]), rng.standard_normal(m["nCells"]), rng.standard_normal(m["nCells"]),
               rng.standard_normal(m["nEdges"])]
        e = 1e-4
        p = A.fe_step(m, u + e * d[0], h + e * d[1], ssh + e * d[2], hE + e * d[3], dt)
        q = A.fe_step(m, u - e * d[0], h - e * d[1], ssh - e * d[2], hE - e * d[3], dt)
        lhs = sum(((a - b) / (2 * e)) @ l for a, b, l in zip(p, q, lam))
        bar = A.fe_step_vjp(m, u, hE, dt, *lam)
        rhs = sum(x @ y for x, y in zip(d, bar))
        assert abs(lhs - rhs) <= 1e-7 * max(abs(lhs), abs(rhs))


def test_fe_gradient_matches_finite_differences_like_the_reference_test():
    m, ssh, u, h, dt = _fe_case(16)
    J, gu, gh, gs, ge = A.gradient_sum_ssh2_fe(m, ssh, u, h, dt, 6)
    assert J > 0 and np.all(np.isfinite(gu)) and np.all(np.isfinite(gh))
    for k in (4, 77, 200):                                               # the reference checks index 5 (1-based)
        fd_h = A.finite_difference_fe(m, ssh, u, h, dt, 6, "h", k, eps=1e-7)
        fd_u = A.finite_difference_fe(m, ssh, u, h, dt, 6, "u", k, eps=1e-4)
        fd_s = A.finite_difference_fe(m, ssh, u, h, dt, 6, "s", k, eps=1e-4)
        assert abs(gh[k] - fd_h) < 1e-4                                  # atol of test_Enzyme_end2end.jl:176
        assert abs(gu[k] - fd_u) < 1e-2                                  # atol of :177
        assert abs(gh[k] - fd_h) < 1e-5 * abs(gh[k]) + 1e-7 and abs(gu[k] - fd_u) < 1e-5 * abs(gu[k]) + 1e-5
        assert abs(gs[k] - fd_s) < 1e-5 * abs(gs[k]) + 1e-7
    # the first step's thickness flux is zero (hEdge starts as zeros), so nothing depends on u through it there; and ssh
    # enters only through the first step's pressure gradient
    assert np.linalg.norm(gs) > 0 and np.linalg.norm(ge) > 0


def test_committed_forward_euler_adjoint_fixture_is_reproduced_by_the_oracle():
    import json
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "igw16_adjoint_fe.npz"))
    meta = json.loads(str(g["meta"]))
    m = hex_mesh(16)
    ssh, u, h = mb.inertialGravityWave(m).initial_state()
    J, gu, gh, gs, ge = A.gradient_sum_ssh2_fe(m, ssh, u, h, meta["dt"], meta["nsteps"])
    assert J == float(g["J"]) and np.array_equal(gu, g["d_normalVelocity"]) and np.array_equal(gh, g["d_layerThickness"])
    assert np.array_equal(gs, g["d_ssh"]) and np.array_equal(ge, g["d_layerThicknessEdge"])
    k = meta["fd_index"]
    assert abs(gh[k] - float(g["fd_layerThickness"])) < 1e-4 and abs(gu[k] - float(g["fd_normalVelocity"])) < 1e-2


def _levels_case(nx, K):
    m = dict(hex_mesh(nx))
    ssh, u, h = mb.inertialGravityWave(m).initial_state()
    frac = np.random.default_rng(K).uniform(0.5, 1.5, K)
    frac /= frac.sum()
    rest = np.outer(np.full(m["nCells"], 1000.0), frac)
    m["restingThickness"], m["nVertLevels"] = rest, K
    uk = np.ascontiguousarray(np.outer(u, 1.0 + 0.1 * np.arange(K)).T)
    hk = np.ascontiguousarray((rest + np.outer(ssh, frac)).T)
    return m, uk, hk, mb.cfl_dt(m["dc"])


def test_multilevel_vjp_is_the_transpose_of_the_multilevel_tendencies():
    """Pins the level-axis adjoint oracle: <J v, w> = <v, J^T w> with J v from central differences of the (quadratic) K-level
    tendencies, whose single pressure gradient per column couples the levels."""
    m, u, h, dt = _levels_case(12, 3)
    rng = np.random.default_rng(2)
    du, dh = rng.standard_normal(u.shape), rng.standard_normal(h.shape)
    wu, wh = rng.standard_normal(u.shape), rng.standard_normal(h.shape)
    e = 1e-3
    fp, fm = O.tendencies_consistent(m, u + e * du, h + e * dh), O.tendencies_consistent(m, u - e * du, h - e * dh)
    tu, th = (fp[0] - fm[0]) / (2 * e), (fp[1] - fm[1]) / (2 * e)
    ub, hb = A.tendencies_vjp_levels(m, u, h, wu, wh)
    lhs, rhs = np.sum(tu * wu) + np.sum(th * wh), np.sum(du * ub) + np.sum(dh * hb)
    assert abs(lhs - rhs) <= 1e-9 * max(abs(lhs), abs(rhs))


def test_multilevel_gradient_matches_finite_differences():
    m, u, h, dt = _levels_case(12, 3)
    nsteps = 4
    J, gu, gh = A.gradient_sum_ssh2_levels(m, u, h, dt, nsteps)

    def objective(uu, hh):
        return float(np.sum(O.ssh_from_thickness(m, A.run_forward_levels(m, uu, hh, dt, nsteps)[-1][1]) ** 2))

    assert abs(J - objective(u, h)) <= 1e-12 * J
    for k, i in ((0, 5), (2, 77), (1, 100)):
        for kind, g, eps in (("h", gh, 1e-6), ("u", gu, 1e-4)):
            up, hp, um, hm = u.copy(), h.copy(), u.copy(), h.copy()
            (hp if kind == "h" else up)[k, i] += eps
            (hm if kind == "h" else um)[k, i] -= eps
            fd = (objective(up, hp) - objective(um, hm)) / (2 * eps)
            assert abs(g[k, i] - fd) <= 1e-5 * abs(g[k, i]) + (1e-6 if kind == "h" else 1e-4), (kind, k, i, g[k, i], fd)
    # with one level it is the single-level oracle
    m1, u1, h1, _ = _levels_case(12, 1)
    J1, gu1, gh1 = A.gradient_sum_ssh2_levels(m1, u1, h1, dt, 3)
    Js, gus, ghs = A.gradient_sum_ssh2(m1, u1[0], h1[0], dt, 3)
    assert abs(J1 - Js) <= 1e-13 * Js and np.allclose(gu1[0], gus, rtol=1e-12, atol=0) and np.allclose(gh1[0], ghs, rtol=1e-12, atol=1e-18)
