"""ORACLE (test infrastructure, not product code): numpy restatement of MOKA's forward hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.

Every function restates one reference kernel or driver routine and cites it (paths relative to
/root/reference).  Arithmetic follows the reference's Julia left-to-right evaluation order per
output element, slot loops run sequentially, so the numpy and the C
restatement (moka_oracle.c, built with -ffp-contract=off) agree bit for bit.

PARITY PINNING.  The reference cannot run here (no Julia).  What its own
tests pin for this path is only the six operator errors of test/ocn/test_Operators.jl:52-53,
72-73,90-91 -- reproduced by `tests/test_oracle_golden.py` on a self-generated 48x48 mesh.
Field values of the tendency kernels, the ForwardEuler step and the run loop are pinned by no
reference test ("parity unpinned" for those rows); the RungeKutta4 stepper of the reference
is dead code (time_integration.jl:61-148) and its semantics here are project-defined, see
`timestep_rk4`.

State layout: dict with ssh[2] (nCells), normalVelocity[2] (nEdges), layerThickness[2] (nCells)
(two time levels, index 0 = previous, -1 = new; PrognosticVars.jl:6-57, nVertLevels = 1 so the
leading singleton dimension of the reference arrays is dropped).
"""
from __future__ import annotations

import numpy as np

GRAVITY = 9.80616          # literal at pressure_gradient.jl:63


# --------------------------------------------------------------------------------------------
# operators (src/ocn/Operators.jl)
# --------------------------------------------------------------------------------------------
def gradient_on_edge(m, scalar_cell):
    """GradientOnEdge, Operators.jl:84-100: (s[c2] - s[c1]) / dcEdge."""
    c1 = m["cellsOnEdge"][:, 0] - 1
    c2 = m["cellsOnEdge"][:, 1] - 1
    return (scalar_cell[..., c2] - scalar_cell[..., c1]) / m["dcEdge"]


def divergence_on_cell(m, vec_edge):
    """DivergenceOnCell_P1 + _P2, Operators.jl:12-44. Returns (div, temp)."""
    temp = vec_edge * m["dvEdge"]                                        # :18
    div = np.zeros(vec_edge.shape[:-1] + (m["nCells"],))                  # :34
    eoc, sgn, n = m["edgesOnCell"], m["edgeSignOnCell"], m["nEdgesOnCell"]
    for i in range(m["maxEdges"]):
        act = i < n
        e = np.where(act, eoc[:, i] - 1, 0)
        div = np.where(act, div - temp[..., e] * sgn[:, i], div)         # :39
    return div / m["areaCell"], temp                                     # :42


def curl_on_vertex(m, vec_edge, curl_in=None):
    """CurlOnVertex, Operators.jl:122-149. Accumulates into curl_in (never zeroed, :135)."""
    curl = np.zeros(vec_edge.shape[:-1] + (m["nVertices"],)) if curl_in is None else curl_in.copy()
    inv = 1.0 / m["areaTriangle"]                                        # :137
    for j in range(m["vertexDegree"]):
        e = m["edgesOnVertex"][:, j] - 1
        curl = curl + m["dcEdge"][e] * inv * vec_edge[..., e] * m["edgeSignOnVertex"][:, j]   # :142-145
    return curl


def interpolate_cell2edge(m, cell_value):
    """interpolateCell2Edge, Operators.jl:201-222."""
    c1 = m["cellsOnEdge"][:, 0] - 1
    c2 = m["cellsOnEdge"][:, 1] - 1
    return 0.5 * (cell_value[..., c1] + cell_value[..., c2])


# --------------------------------------------------------------------------------------------
# tendencies (src/ocn/Tendencies)
# --------------------------------------------------------------------------------------------
def ssh_grad_on_edge(m, tend, ssh):
    """SSHGradOnEdge!, pressure_gradient.jl:45-65: tend -= 9.80616 * (1/dc) * (ssh[c2]-ssh[c1])."""
    c1 = m["cellsOnEdge"][:, 0] - 1
    c2 = m["cellsOnEdge"][:, 1] - 1
    inv = 1.0 / m["dcEdge"]
    return tend - GRAVITY * inv * (ssh[c2] - ssh[c1])


def coriolis_force_tendency(m, tend, u):
    """coriolis_force_tendency_kernel!, horizontal_advection_and_coriolis.jl:50-75."""
    eoe, w, n, f = m["edgesOnEdge"], m["weightsOnEdge"], m["nEdgesOnEdge"], m["fEdge"]
    out = tend.copy()
    for i in range(eoe.shape[1]):
        act = (i < n) & (eoe[:, i] != 0)                                 # :61, :67
        j = np.where(act, eoe[:, i] - 1, 0)
        out = np.where(act, out + w[:, i] * u[..., j] * f[j], out)       # :70-72
    return out


def thickness_flux_div_on_cell(m, tend, flux):
    """thicknessFluxDivOnCell!, horizontal_advection.jl:42-69."""
    inv_area = 1.0 / m["areaCell"]                                       # :55
    eoc, sgn, n = m["edgesOnCell"], m["edgeSignOnCell"], m["nEdgesOnCell"]
    out = tend.copy()
    for i in range(m["maxEdges"]):
        act = i < n
        e = np.where(act, eoc[:, i] - 1, 0)
        out = np.where(act, out + flux[..., e] * m["dvEdge"][e] * sgn[:, i] * inv_area, out)   # :64-65
    return out


def compute_normal_velocity_tendency(m, ssh, u):
    """computeNormalVelocityTendency!, normalVelocity.jl:21-53 (zero, gradient, Coriolis)."""
    t = np.zeros(u.shape[:-1] + (m["nEdges"],))
    t = ssh_grad_on_edge(m, t, ssh)
    return coriolis_force_tendency(m, t, u)


def compute_layer_thickness_tendency(m, flux):
    """computeLayerThicknessTendency!, layerThickness.jl:14-28."""
    return thickness_flux_div_on_cell(m, np.zeros(flux.shape[:-1] + (m["nCells"],)), flux)


# --------------------------------------------------------------------------------------------
# diagnostics (src/ocn/DiagnosticVars.jl:108-207)
# --------------------------------------------------------------------------------------------
def new_diag(m):
    """DiagnosticVars(config, mesh), DiagnosticVars.jl:75-99: zeros."""
    return {"layerThicknessEdge": np.zeros(m["nEdges"]), "thicknessFlux": np.zeros(m["nEdges"]),
            "velocityDivCell": np.zeros(m["nCells"]),
            "relativeVorticity": np.zeros(m["nVertices"])}


def diagnostic_compute(m, diag, u, h):
    """diagnostic_compute!, DiagnosticVars.jl:108-117, in the reference's order (Q1, Q2)."""
    diag["thicknessFlux"] = u * diag["layerThicknessEdge"]               # :141-173, stale hEdge
    div, temp = divergence_on_cell(m, u)                                 # :175-193
    diag["velocityDivCell"] = div
    diag["layerThicknessEdge"] = temp                                    # scratch alias :187-190
    if m["nVertices"]:
        diag["relativeVorticity"] = curl_on_vertex(m, u, diag["relativeVorticity"])   # :195-207
    diag["layerThicknessEdge"] = interpolate_cell2edge(m, h)             # :126-139


# --------------------------------------------------------------------------------------------
# time stepping (src/forward/time_integration.jl)
# --------------------------------------------------------------------------------------------
def resting_thickness_sum(m):
    """restingThicknessSum = sum(restingThickness; dims=1), VertMesh.jl:73."""
    return m["restingThickness"].sum(axis=1)


def ssh_from_thickness(m, h):
    """Update_ssh!, time_integration.jl:205-212: ssh = layerThickness[1, j] - restingThicknessSum[j] for the single layer the
    reference runs.  MULTI-LEVEL (project-defined; state arrays of shape (nVertLevels, n), level-major): the
    free surface is the top of the whole column, ssh = (h[0] + h[1] + ... in level order) - restingThicknessSum -- with one
    level the reference's expression, bit for bit.  Everything else in this file takes the level axis along by broadcasting:
    the reference's kernels already carry the `for k in 1:maxLevelEdgeTop` loops (pressure_gradient.jl:61-64,
    horizontal_advection_and_coriolis.jl:69-73, horizontal_advection.jl:60-66) with the SAME pressure gradient for every level."""
    if h.ndim == 1:
        return h - resting_thickness_sum(m)
    col = h[0].copy()
    for k in range(1, h.shape[0]):
        col = col + h[k]
    return col - resting_thickness_sum(m)


def new_state(m, ssh, u, h):
    """PrognosticVars(ssh, normalVelocity, layerThickness, 2), PrognosticVars.jl:28-56."""
    return {"ssh": [ssh.copy(), ssh.copy()], "normalVelocity": [u.copy(), u.copy()],
            "layerThickness": [h.copy(), h.copy()]}


def advance_time_levels(prog):
    """advanceTimeLevels!, time_integration.jl:10-59: field[1] <- field[2]."""
    for k in ("ssh", "normalVelocity", "layerThickness"):
        prog[k][0] = prog[k][-1].copy()


def timestep_forward_euler(m, prog, diag, dt):
    """ocn_timestep(::ForwardEuler), time_integration.jl:150-193 (the live path)."""
    advance_time_levels(prog)
    diagnostic_compute(m, diag, prog["normalVelocity"][-1], prog["layerThickness"][-1])
    tu = compute_normal_velocity_tendency(m, prog["ssh"][-1], prog["normalVelocity"][-1])
    th = compute_layer_thickness_tendency(m, diag["thicknessFlux"])
    prog["normalVelocity"][-1] = prog["normalVelocity"][-1] + dt * tu    # :196-202
    prog["layerThickness"][-1] = prog["layerThickness"][-1] + dt * th
    prog["ssh"][-1] = ssh_from_thickness(m, prog["layerThickness"][-1])       # :205-212
    return tu, th


def tendencies_consistent(m, u, h):
    """Tendencies at a (provisional) state with diagnostics of that same state.

    RK4 lines time_integration.jl:124-130 taken at face value: hEdge from the provisional h
    (interpolateCell2Edge), flux = u*hEdge (compute_thicknessFlux!), then the two tendency
    entry points -- without the ForwardEuler ordering artefact Q1.
    """
    ssh = ssh_from_thickness(m, h)                                       # :127
    flux = u * interpolate_cell2edge(m, h)
    return compute_normal_velocity_tendency(m, ssh, u), compute_layer_thickness_tendency(m, flux)


def timestep_rk4(m, prog, dt):
    """ocn_timestep(::RungeKutta4) as intended by time_integration.jl:61-148 (project-defined).

    a = [dt/2, dt/2, dt], b = [dt/6, dt/3, dt/3, dt/6] (:77-78); New <- state (:108-110);
    per stage: tendencies at the provisional state (:114-115); s<4: Provis = Curr + a_s*tend
    (:124-125); New += b_s*tend (:134-135); ssh = h - restingThicknessSum (:127,:136).
    """
    advance_time_levels(prog)
    a = [dt / 2.0, dt / 2.0, dt]
    b = [dt / 6.0, dt / 3.0, dt / 3.0, dt / 6.0]
    u_cur, h_cur = prog["normalVelocity"][0], prog["layerThickness"][0]
    u_pro, h_pro = u_cur.copy(), h_cur.copy()
    u_new, h_new = u_cur.copy(), h_cur.copy()
    for s in range(4):
        tu, th = tendencies_consistent(m, u_pro, h_pro)
        if s < 3:
            u_pro = u_cur + a[s] * tu
            h_pro = h_cur + a[s] * th
        u_new = u_new + b[s] * tu
        h_new = h_new + b[s] * th
    prog["normalVelocity"][-1] = u_new
    prog["layerThickness"][-1] = h_new
    prog["ssh"][-1] = ssh_from_thickness(m, h_new)


def sum_array(ssh):
    """sumArray, run_loop.jl:47-51: serial sum of ssh[j]^2 in index order."""
    s = 0.0
    for v in ssh.tolist():
        s = s + v * v
    return s


def reference_dt(m):
    """ocn_init_alarms, init.jl:118: dt = floor(2*(mean(dc)/1e3)*mean(dc)/200e3) seconds."""
    d = float(np.mean(m["dcEdge"]))
    return float(np.floor(2 * (d / 1e3) * d / 200e3))


# --------------------------------------------------------------------------------------------
# inertial gravity wave (src/inertialGravityWave.jl)
# --------------------------------------------------------------------------------------------
class InertialGravityWave:
    """Constants inertialGravityWave.jl:6-19; lx follows the mesh period (10 000 km in the
    reference's polaris case) so the same formulas serve every mesh size."""

    def __init__(self, m, lx_km=None):
        self.g, self.f0, self.npx, self.npy = GRAVITY, 1e-4, 2.0, 2.0
        self.eta0, self.bottom_depth = 1.0, 1000.0
        self.lx = m["x_period"] / 1e3 if lx_km is None else lx_km
        self.ly = np.sqrt(3.0) / 2.0 * self.lx
        self.kx = self.npx * 2.0 * np.pi / (self.lx * 1e3)
        self.ky = self.npy * 2.0 * np.pi / (self.ly * 1e3)
        self.omega = np.sqrt(self.f0 ** 2 + self.g * self.bottom_depth * (self.kx ** 2 + self.ky ** 2))
        self.m = m

    def exact_ssh(self, t):                                              # :38-45
        m = self.m
        return self.eta0 * np.cos(self.kx * m["xCell"] + self.ky * m["yCell"] - self.omega * t)

    def exact_norm_vel(self, t):                                         # :47-64
        m = self.m
        ph = self.kx * m["xEdge"] + self.ky * m["yEdge"] - self.omega * t
        c = self.g / (self.omega ** 2.0 - self.f0 ** 2.0)
        u = self.eta0 * (c * (self.omega * self.kx * np.cos(ph) - self.f0 * self.ky * np.sin(ph)))
        v = self.eta0 * (c * (self.omega * self.ky * np.cos(ph) + self.f0 * self.kx * np.sin(ph)))
        return u * np.cos(m["angleEdge"]) + v * np.sin(m["angleEdge"])

    def initial_state(self):
        ssh = self.exact_ssh(0.0)
        return ssh, self.exact_norm_vel(0.0), self.bottom_depth + ssh


# --------------------------------------------------------------------------------------------
# error norms and analytic fields of the operator test (test/utilities.jl)
# --------------------------------------------------------------------------------------------
def error_measures(numeric, analytic, area):
    """ErrorMeasures, test/utilities.jl:18-28. Returns (L_two, L_inf)."""
    d = analytic - numeric
    linf = np.max(np.abs(d)) / np.max(np.abs(analytic))
    ltwo = np.linalg.norm(d * area) / np.linalg.norm(analytic * area)
    return ltwo, linf


def planar_test_fields(m):
    """Analytic fields of test/utilities.jl:93-190 (PlanarTest)."""
    Lx = np.round(np.max(m["xCell"]))                                    # :71
    Ly = np.sqrt(3.0) / 2.0 * Lx                                         # :72
    xc, yc, xe, ye = m["xCell"], m["yCell"], m["xEdge"], m["yEdge"]
    nxn, nyn = np.cos(m["angleEdge"]), np.sin(m["angleEdge"])
    tp = 2.0 * np.pi
    out = {"h": np.sin(tp * xc / Lx) * np.sin(tp * yc / Ly)}
    Fx = np.sin(tp * xe / Lx) * np.cos(tp * ye / Ly)
    Fy = np.cos(tp * xe / Lx) * np.sin(tp * ye / Ly)
    out["F_edge"] = nxn * Fx + nyn * Fy
    dhdx = tp / Lx * np.cos(tp * xe / Lx) * np.sin(tp * ye / Ly)
    dhdy = tp / Ly * np.sin(tp * xe / Lx) * np.cos(tp * ye / Ly)
    out["grad_h_edge"] = nxn * dhdx + nyn * dhdy
    out["div_F"] = tp * (1.0 / Lx + 1.0 / Ly) * np.cos(tp * xc / Lx) * np.cos(tp * yc / Ly)
    if m["nVertices"]:
        xv, yv = m["xVertex"], m["yVertex"]
        out["curl_F"] = tp * (-1.0 / Lx + 1.0 / Ly) * np.sin(tp * xv / Lx) * np.sin(tp * yv / Ly)
    return out


def rel_l2(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))
