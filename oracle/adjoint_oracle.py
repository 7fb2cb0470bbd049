"""ORACLE (test infrastructure, not product code): reverse-mode (discrete adjoint) of the RK4 forward path.

Only tests/ may import this.

What it stands for.  The reference obtains gradients of its run loop with Enzyme:
`autodiff(Reverse, ocn_run_loop, Duplicated(Prog, d_Prog), ...)` of J = sum(ssh[end].^2)
(test/enzyme/test_Enzyme_end2end.jl:30-110; objective = sumArray, src/forward/run_loop.jl:47-51;
custom rule for the device->host copy ext/MPASEnzymeExt.jl:13-38), and checks d_Prog against central
finite differences of J at one cell and one edge (test_Enzyme_end2end.jl:112-180, atol 1e-4 / 1e-2).
It does so through ForwardEuler and reports NaN on CUDA (`@test_broken`, :182-186).

Here the same quantity -- dJ/d(normalVelocity_0, layerThickness_0) -- is restated for the project-defined
RungeKutta4 stepper (moka_oracle.timestep_rk4), by hand, in SCATTER form: every forward statement
`y[i] += c * x[j]` becomes `xbar[j] += c * ybar[i]` with np.add.at, statement by statement in reverse
order.  The CUDA adjoint uses the transposed (gather) form, so the two are independent formulations.

PARITY PINNING: unpinned by the reference (it holds no gradient values); pinned here by the reference's own
acceptance test -- agreement with central finite differences of J -- and by the dot-product identity
<J v, w> == <v, J^T w> (tests/test_adjoint_oracle.py).
"""
from __future__ import annotations

import numpy as np

import moka_oracle as O


def tendencies_vjp(m, u, h, ku_bar, kh_bar):
    """(ubar, hbar) = (dF/d(u,h))^T (ku_bar, kh_bar) at (u, h), F = moka_oracle.tendencies_consistent."""
    nC, nE = m["nCells"], m["nEdges"]
    c1 = m["cellsOnEdge"][:, 0] - 1
    c2 = m["cellsOnEdge"][:, 1] - 1
    ubar, hbar = np.zeros(nE), np.zeros(nC)
    # reverse of compute_layer_thickness_tendency(flux) (horizontal_advection.jl:64-65)
    inv_area = 1.0 / m["areaCell"]
    eoc, sgn, n = m["edgesOnCell"], m["edgeSignOnCell"], m["nEdgesOnCell"]
    flux_bar = np.zeros(nE)
    for i in range(m["maxEdges"]):
        act = i < n
        e = eoc[act, i] - 1
        np.add.at(flux_bar, e, m["dvEdge"][e] * sgn[act, i] * inv_area[act] * kh_bar[act])
    # reverse of flux = u * interpolateCell2Edge(h) (DiagnosticVars.jl:158-173, Operators.jl:201-222)
    ubar += flux_bar * (0.5 * (h[c1] + h[c2]))
    hedge_bar = flux_bar * u
    np.add.at(hbar, c1, 0.5 * hedge_bar)
    np.add.at(hbar, c2, 0.5 * hedge_bar)
    # reverse of coriolis_force_tendency (horizontal_advection_and_coriolis.jl:70-72)
    eoe, w, ne, f = m["edgesOnEdge"], m["weightsOnEdge"], m["nEdgesOnEdge"], m["fEdge"]
    for i in range(eoe.shape[1]):
        act = (i < ne) & (eoe[:, i] != 0)
        j = eoe[act, i] - 1
        np.add.at(ubar, j, w[act, i] * f[j] * ku_bar[act])
    # reverse of ssh_grad_on_edge (pressure_gradient.jl:63) with ssh = h - restingThicknessSum
    gk = O.GRAVITY * (1.0 / m["dcEdge"]) * ku_bar
    np.add.at(hbar, c2, -gk)
    np.add.at(hbar, c1, gk)
    return ubar, hbar


def tendencies_jvp(m, u, h, du, dh):
    """Directional derivative dF[(u,h)](du, dh) (F is linear in u and bilinear in (u, h))."""
    c1 = m["cellsOnEdge"][:, 0] - 1
    c2 = m["cellsOnEdge"][:, 1] - 1
    tu = O.coriolis_force_tendency(m, O.ssh_grad_on_edge(m, np.zeros(m["nEdges"]), dh), du)
    dflux = du * (0.5 * (h[c1] + h[c2])) + u * (0.5 * (dh[c1] + dh[c2]))
    return tu, O.compute_layer_thickness_tendency(m, dflux)


def rk4_stage_states(m, u, h, dt):
    """The four states the tendencies of one RK4 step are evaluated at (moka_oracle.timestep_rk4)."""
    a = [dt / 2.0, dt / 2.0, dt]
    ys = [(u, h)]
    for s in range(3):
        tu, th = O.tendencies_consistent(m, *ys[-1])
        ys.append((u + a[s] * tu, h + a[s] * th))
    return ys


def rk4_step(m, u, h, dt):
    prog = O.new_state(m, h - O.resting_thickness_sum(m), u, h)
    O.timestep_rk4(m, prog, dt)
    return prog["normalVelocity"][-1], prog["layerThickness"][-1]


def rk4_step_vjp(m, u, h, dt, lam_u, lam_h):
    """Adjoint of one RK4 step x' = x + sum_s b_s k_s, k_s = F(y_s), y_1 = x, y_{s+1} = x + a_s k_s:
    kbar_4 = b_4 lam', ybar_s = J_s^T kbar_s, kbar_{s-1} = b_{s-1} lam' + a_{s-1} ybar_s,
    lam = lam' + sum_s ybar_s."""
    a = [dt / 2.0, dt / 2.0, dt]
    b = [dt / 6.0, dt / 3.0, dt / 3.0, dt / 6.0]
    ys = rk4_stage_states(m, u, h, dt)
    out_u, out_h = lam_u.copy(), lam_h.copy()
    kbu, kbh = b[3] * lam_u, b[3] * lam_h
    for s in (3, 2, 1, 0):
        ybu, ybh = tendencies_vjp(m, ys[s][0], ys[s][1], kbu, kbh)
        out_u, out_h = out_u + ybu, out_h + ybh
        if s > 0:
            kbu, kbh = b[s - 1] * lam_u + a[s - 1] * ybu, b[s - 1] * lam_h + a[s - 1] * ybh
    return out_u, out_h


def objective_sum_ssh2(m, h):
    """sumArray over ssh = h - restingThicknessSum (run_loop.jl:47-51, time_integration.jl:205-212)."""
    ssh = h - O.resting_thickness_sum(m)
    return float(np.sum(ssh * ssh))


def run_forward(m, u, h, dt, nsteps):
    traj = [(u, h)]
    for _ in range(nsteps):
        traj.append(rk4_step(m, *traj[-1], dt))
    return traj


def gradient_sum_ssh2(m, u0, h0, dt, nsteps):
    """(J, dJ/du0, dJ/dh0) for J = sum ssh_N^2 after `nsteps` RK4 steps: what the reference's
    `autodiff(Reverse, ocn_run_loop, ...)` leaves in d_Prog (test_Enzyme_end2end.jl:78-96)."""
    traj = run_forward(m, u0, h0, dt, nsteps)
    uN, hN = traj[-1]
    lam_u = np.zeros(m["nEdges"])
    lam_h = 2.0 * (hN - O.resting_thickness_sum(m))
    for n in range(nsteps - 1, -1, -1):
        lam_u, lam_h = rk4_step_vjp(m, traj[n][0], traj[n][1], dt, lam_u, lam_h)
    return objective_sum_ssh2(m, hN), lam_u, lam_h


def finite_difference(m, u0, h0, dt, nsteps, kind, k, eps=1e-8):
    """Central difference of J in one component, as test_Enzyme_end2end.jl:112-170 (relative step eps)."""
    up, um, hp, hm = u0.copy(), u0.copy(), h0.copy(), h0.copy()
    if kind == "h":
        hp[k] += abs(h0[k]) * eps
        hm[k] -= abs(h0[k]) * eps
        dist = hp[k] - hm[k]
    else:
        up[k] += abs(u0[k]) * eps
        um[k] -= abs(u0[k]) * eps
        dist = up[k] - um[k]
    jp = objective_sum_ssh2(m, run_forward(m, up, hp, dt, nsteps)[-1][1])
    jm = objective_sum_ssh2(m, run_forward(m, um, hm, dt, nsteps)[-1][1])
    return (jp - jm) / dist


# ---- multi-level states (nVertLevels = K; arrays of shape (K, n), project-defined semantics: moka_oracle.ssh_from_thickness) ----
def tendencies_vjp_levels(m, u, h, ku_bar, kh_bar):
    """(ubar, hbar) = (dF/d(u, h))^T (ku_bar, kh_bar) for the K-level tendencies (moka_oracle.tendencies_consistent on (K, n)
    arrays): Coriolis and thickness flux act level by level, the pressure gradient -g/dc (ssh2 - ssh1) with
    ssh = sum_k h_k - restingThicknessSum is ONE term shared by every level of the column -- so its adjoint, formed from
    the level sum of ku_bar, goes to every level of h alike."""
    K = u.shape[0]
    c1 = m["cellsOnEdge"][:, 0] - 1
    c2 = m["cellsOnEdge"][:, 1] - 1
    ubar, hbar = np.zeros_like(u), np.zeros_like(h)
    g_dc = O.GRAVITY * (1.0 / m["dcEdge"])
    for k in range(K):
        ub, hb = tendencies_vjp(m, u[k], h[k], ku_bar[k], kh_bar[k])
        gk = g_dc * ku_bar[k]                                            # take this level's own pressure adjoint out again ...
        np.add.at(hb, c2, gk)
        np.add.at(hb, c1, -gk)
        ubar[k], hbar[k] = ub, hb
    gsum = g_dc * ku_bar.sum(axis=0)                                     # ... and give every level the column's
    p = np.zeros(m["nCells"])
    np.add.at(p, c2, -gsum)
    np.add.at(p, c1, gsum)
    return ubar, hbar + p[None, :]


def rk4_step_vjp_levels(m, u, h, dt, lam_u, lam_h):
    a = [dt / 2.0, dt / 2.0, dt]
    b = [dt / 6.0, dt / 3.0, dt / 3.0, dt / 6.0]
    ys = rk4_stage_states(m, u, h, dt)                                   # (tendencies_consistent broadcasts over the level axis)
    out_u, out_h = lam_u.copy(), lam_h.copy()
    kbu, kbh = b[3] * lam_u, b[3] * lam_h
    for s in (3, 2, 1, 0):
        ybu, ybh = tendencies_vjp_levels(m, ys[s][0], ys[s][1], kbu, kbh)
        out_u, out_h = out_u + ybu, out_h + ybh
        if s > 0:
            kbu, kbh = b[s - 1] * lam_u + a[s - 1] * ybu, b[s - 1] * lam_h + a[s - 1] * ybh
    return out_u, out_h


def run_forward_levels(m, u, h, dt, nsteps):
    traj = [(u, h)]
    for _ in range(nsteps):
        prog = O.new_state(m, O.ssh_from_thickness(m, traj[-1][1]), traj[-1][0], traj[-1][1])
        O.timestep_rk4(m, prog, dt)
        traj.append((prog["normalVelocity"][-1], prog["layerThickness"][-1]))
    return traj


def gradient_sum_ssh2_levels(m, u0, h0, dt, nsteps):
    """(J, dJ/du0, dJ/dh0) of J = sum ssh_N^2, ssh = sum_k h_k - restingThicknessSum, for (K, n) arrays."""
    traj = run_forward_levels(m, u0, h0, dt, nsteps)
    sshN = O.ssh_from_thickness(m, traj[-1][1])
    lam_u = np.zeros_like(u0)
    lam_h = np.repeat((2.0 * sshN)[None, :], u0.shape[0], axis=0)
    for n in range(nsteps - 1, -1, -1):
        lam_u, lam_h = rk4_step_vjp_levels(m, traj[n][0], traj[n][1], dt, lam_u, lam_h)
    return float(np.sum(sshN * sshN)), lam_u, lam_h


# ---- ForwardEuler: the stepper the reference actually differentiates (test_Enzyme_end2end.jl:78-96) ---------------
# One step of moka_oracle.timestep_forward_euler maps (u, h, ssh, hE) -> (u', h', ssh', hE'), hE = Diag.layerThicknessEdge:
#   flux = u * hE                       (the LAGGED hEdge: diagnostic_compute! forms the flux before it refreshes hEdge,
#                                        DiagnosticVars.jl:112-116; zeros on the first step)
#   hE'  = interpolateCell2Edge(h)      (Operators.jl:201-222)
#   u'   = u + dt * (-(g/dc) (ssh[c2] - ssh[c1]) + sum_i w_i u[eoe_i] f[eoe_i])
#   h'   = h + dt * (1/area) sum_i flux[e_i] dv[e_i] sign_i
#   ssh' = h' - restingThicknessSum
# so ssh is an independent input of the FIRST step only, and Enzyme's d_Prog.ssh[end] is the gradient with respect to it.
def fe_step(m, u, h, ssh, hE, dt):
    flux = u * hE
    tu = O.compute_normal_velocity_tendency(m, ssh, u)
    th = O.compute_layer_thickness_tendency(m, flux)
    h_new = h + dt * th
    return u + dt * tu, h_new, h_new - O.resting_thickness_sum(m), O.interpolate_cell2edge(m, h)


def fe_step_vjp(m, u, hE, dt, lam_u, lam_h, lam_s, lam_e):
    """Adjoint of one ForwardEuler step in SCATTER form, statement by statement in reverse order.  (u, hE) are the step's
    inputs the Jacobian depends on; returns the adjoints of (u, h, ssh, hE)."""
    nC, nE = m["nCells"], m["nEdges"]
    c1 = m["cellsOnEdge"][:, 0] - 1
    c2 = m["cellsOnEdge"][:, 1] - 1
    out_u, out_h, out_s, out_e = lam_u.copy(), np.zeros(nC), np.zeros(nC), np.zeros(nE)
    # hE' = 0.5 * (h[c1] + h[c2])
    np.add.at(out_h, c1, 0.5 * lam_e)
    np.add.at(out_h, c2, 0.5 * lam_e)
    # ssh' = h' - H ; h' = h + dt * th
    mu = lam_h + lam_s
    out_h += mu
    # th[c] = invArea[c] * sum_i flux[e_i] * dv[e_i] * sign[i, c]
    inv_area = 1.0 / m["areaCell"]
    eoc, sgn, n = m["edgesOnCell"], m["edgeSignOnCell"], m["nEdgesOnCell"]
    flux_bar = np.zeros(nE)
    for i in range(m["maxEdges"]):
        act = i < n
        e = eoc[act, i] - 1
        np.add.at(flux_bar, e, dt * m["dvEdge"][e] * sgn[act, i] * inv_area[act] * mu[act])
    # flux = u * hE
    out_u += flux_bar * hE
    out_e += flux_bar * u
    # tu: Coriolis (horizontal_advection_and_coriolis.jl:70-72)
    eoe, w, ne, f = m["edgesOnEdge"], m["weightsOnEdge"], m["nEdgesOnEdge"], m["fEdge"]
    for i in range(eoe.shape[1]):
        act = (i < ne) & (eoe[:, i] != 0)
        j = eoe[act, i] - 1
        np.add.at(out_u, j, dt * w[act, i] * f[j] * lam_u[act])
    # tu: pressure gradient (pressure_gradient.jl:63) on the ssh ARRAY
    gk = dt * O.GRAVITY * (1.0 / m["dcEdge"]) * lam_u
    np.add.at(out_s, c2, -gk)
    np.add.at(out_s, c1, gk)
    return out_u, out_h, out_s, out_e


def run_forward_fe(m, ssh0, u0, h0, dt, nsteps, hE0=None):
    hE = np.zeros(m["nEdges"]) if hE0 is None else hE0                    # DiagnosticVars.jl:90-93
    traj = [(u0, h0, ssh0, hE)]
    for _ in range(nsteps):
        traj.append(fe_step(m, *traj[-1], dt))
    return traj


def gradient_sum_ssh2_fe(m, ssh0, u0, h0, dt, nsteps, hE0=None):
    """(J, dJ/du0, dJ/dh0, dJ/dssh0, dJ/dhE0) for J = sum ssh_N^2 after `nsteps` ForwardEuler steps: what
    `autodiff(Reverse, ocn_run_loop, ..., Duplicated(Prog, d_Prog), Duplicated(Diag, d_Diag), ...)` leaves in
    d_Prog.{normalVelocity, layerThickness, ssh}[end] and d_Diag.layerThicknessEdge (test_Enzyme_end2end.jl:78-96)."""
    traj = run_forward_fe(m, ssh0, u0, h0, dt, nsteps, hE0)
    sshN = traj[-1][2]
    lam = (np.zeros(m["nEdges"]), np.zeros(m["nCells"]), 2.0 * sshN, np.zeros(m["nEdges"]))
    for n in range(nsteps - 1, -1, -1):
        lam = fe_step_vjp(m, traj[n][0], traj[n][3], dt, *lam)
    return (float(np.sum(sshN * sshN)),) + tuple(lam)


def finite_difference_fe(m, ssh0, u0, h0, dt, nsteps, kind, k, eps=1e-8):
    """Central difference of J in one component of layerThickness ("h"), normalVelocity ("u") or ssh ("s"), with the
    relative step of test_Enzyme_end2end.jl:112-170."""
    x = {"h": h0, "u": u0, "s": ssh0}[kind]
    step = abs(x[k]) * eps if x[k] != 0.0 else eps
    out = []
    for sgn in (1.0, -1.0):
        a = {"h": h0.copy(), "u": u0.copy(), "s": ssh0.copy()}
        a[kind][k] += sgn * step
        sshN = run_forward_fe(m, a["s"], a["u"], a["h"], dt, nsteps)[-1][2]
        out.append((float(np.sum(sshN * sshN)), a[kind][k]))
    return (out[0][0] - out[1][0]) / (out[0][1] - out[1][1])
