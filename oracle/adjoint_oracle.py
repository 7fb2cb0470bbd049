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
