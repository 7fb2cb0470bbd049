"""GPU: multi-level states (nVertLevels = 10 like the reference's own operator test,
test/ocn/test_Operators.jl:18).  The reference's kernels carry the level loops (pressure_gradient.jl:61-64,
horizontal_advection_and_coriolis.jl:69-73, horizontal_advection.jl:60-66); the oracle takes the level axis along by
broadcasting (oracle/moka_oracle.py: ssh_from_thickness), the library steps level-major columns with the static data read
once per column (fused::k_rk_stage_ml)."""
import numpy as np
import pytest

import moka_b200 as mb
import moka_oracle as O
import moka_oracle_c as OC
from conftest import hex_mesh, rel_l2

pytestmark = pytest.mark.gpu


def _column_case(nx, K, seed=0):
    """A K-level column over the inertia-gravity-wave state: level thicknesses that differ per level and sum to H + ssh,
    level velocities that differ per level (a sheared copy of the wave's velocity)."""
    m = dict(hex_mesh(nx))
    ssh, u, h = mb.inertialGravityWave(m).initial_state()
    rng = np.random.default_rng(seed)
    frac = rng.uniform(0.5, 1.5, K)
    frac /= frac.sum()
    rest = np.outer(np.full(m["nCells"], 1000.0), frac)                  # restingThickness (nCells, K)
    hk = rest + np.outer(ssh, frac)                                      # layerThickness   (nCells, K): columns sum to H + ssh
    uk = np.outer(u, 1.0 + 0.1 * np.arange(K))                           # normalVelocity   (nEdges, K)
    m["restingThickness"], m["nVertLevels"] = rest, K
    return m, ssh, uk, hk


def _oracle_state(m, ssh, uk, hk):
    # oracle arrays are level-major (K, n); restingThicknessSum comes from m["restingThickness"].sum(axis=1)
    return O.new_state(m, ssh, np.ascontiguousarray(uk.T), np.ascontiguousarray(hk.T))


@pytest.mark.parametrize("K", [1, 3, 10])
def test_multilevel_rk4_fused_and_unfused_are_the_oracle(backend, K):
    m, ssh, uk, hk = _column_case(32, K)
    dt = mb.cfl_dt(m["dc"])
    mesh = mb.Mesh(m, backend)
    prog = _oracle_state(m, ssh, uk, hk)
    for _ in range(5):
        O.timestep_rk4(m, prog, dt)
    for fused in (True, False):
        p = mb.PrognosticVars(ssh, uk if K > 1 else uk[:, 0], hk if K > 1 else hk[:, 0], 2, mesh)
        mb.ocn_timestep(dt, p, None, None, None, mb.RungeKutta4, nsteps=3, fused=fused)
        mb.ocn_timestep(dt, p, None, None, None, mb.RungeKutta4, nsteps=2, fused=fused)     # odd + even: both time-level parities
        gu, gh = np.asarray(p.normalVelocity).reshape(m["nEdges"], K), np.asarray(p.layerThickness).reshape(m["nCells"], K)
        assert np.array_equal(gu.T, prog["normalVelocity"][-1]) and np.array_equal(gh.T, prog["layerThickness"][-1]), (K, fused)
        assert np.array_equal(p.ssh, prog["ssh"][-1]), (K, fused)
        assert np.array_equal(np.asarray(p.normalVelocity_prev).reshape(m["nEdges"], K).T, prog["normalVelocity"][0])
        mass = mb.reduce_sum(p, "mass")
        assert abs(mass - float(np.sum(m["areaCell"] * hk.sum(axis=1)))) <= 1e-13 * mass
        assert abs(mb.reduce_sum(p, "ssh2") - float(np.sum(prog["ssh"][-1] ** 2))) <= 1e-12 * float(np.sum(prog["ssh"][-1] ** 2))


def test_multilevel_numpy_oracle_with_one_level_is_the_c_oracle():
    """Pins the level-axis generalisation of the numpy oracle: with one level it is the C restatement bit for bit."""
    m, ssh, uk, hk = _column_case(24, 1)
    dt = mb.cfl_dt(m["dc"])
    prog = _oracle_state(m, ssh, uk, hk)
    om = OC.OracleModel(m, ssh, uk[:, 0], hk[:, 0])
    for _ in range(4):
        O.timestep_rk4(m, prog, dt)
    om.run_loop(dt, 4, "RungeKutta4")
    assert np.array_equal(prog["normalVelocity"][-1][0], om.normalVelocity[1]) and np.array_equal(prog["ssh"][-1], om.ssh[1])


def test_multilevel_forward_euler_and_entry_points(backend):
    """The reference's live stepper and the src/ocn entry points on a 10-level state, level by level in the reference's
    operation order (bit for bit the oracle with the level axis)."""
    K = 10
    m, ssh, uk, hk = _column_case(24, K, seed=3)
    dt = 100.0
    mesh = mb.Mesh(m, backend)
    p = mb.PrognosticVars(ssh, uk, hk, 2, mesh)
    diag, tend = mb.DiagnosticVars(p), mb.TendencyVars(p)
    prog = _oracle_state(m, ssh, uk, hk)
    od = {"layerThicknessEdge": np.zeros((K, m["nEdges"])), "thicknessFlux": np.zeros((K, m["nEdges"])),
          "velocityDivCell": np.zeros((K, m["nCells"])), "relativeVorticity": np.zeros((K, m["nVertices"]))}
    mb.ocn_timestep(dt, p, diag, tend, None, mb.ForwardEuler, nsteps=4)
    for _ in range(4):
        tu, th = O.timestep_forward_euler(m, prog, od, dt)
    lv = lambda a, n: np.asarray(a).reshape(n, K).T                      # (n, K) -> level-major
    assert np.array_equal(lv(p.normalVelocity, m["nEdges"]), prog["normalVelocity"][-1])
    assert np.array_equal(lv(p.layerThickness, m["nCells"]), prog["layerThickness"][-1]) and np.array_equal(p.ssh, prog["ssh"][-1])
    assert np.array_equal(lv(tend.tendNormalVelocity, m["nEdges"]), tu) and np.array_equal(lv(tend.tendLayerThickness, m["nCells"]), th)
    assert np.array_equal(lv(diag.thicknessFlux, m["nEdges"]), od["thicknessFlux"])
    assert np.array_equal(lv(diag.relativeVorticity, m["nVertices"]), od["relativeVorticity"])


def test_identical_layers_move_like_one_layer(backend):
    """Known answer: K layers with resting thickness H / K, the same velocity and h_k = (H + ssh) / K are one layer cut into
    slices -- every layer keeps the single-layer velocity and the column keeps the single-layer thickness."""
    K = 8
    m = dict(hex_mesh(32))
    ssh, u, h = mb.inertialGravityWave(m).initial_state()
    dt = mb.cfl_dt(m["dc"])
    one = mb.PrognosticVars(ssh, u, h, 2, mb.Mesh(m, backend))
    mb.ocn_timestep(dt, one, None, None, None, mb.RungeKutta4, nsteps=20)
    mk = dict(m)
    mk["restingThickness"], mk["nVertLevels"] = np.full((m["nCells"], K), 1000.0 / K), K
    many = mb.PrognosticVars(ssh, np.outer(u, np.ones(K)), np.outer(h, np.full(K, 1.0 / K)), 2, mb.Mesh(mk, backend))
    mb.ocn_timestep(dt, many, None, None, None, mb.RungeKutta4, nsteps=20)
    uk, hk = np.asarray(many.normalVelocity), np.asarray(many.layerThickness)
    assert all(rel_l2(uk[:, k], one.normalVelocity) <= 1e-12 for k in range(K))
    assert rel_l2(hk.sum(axis=1), one.layerThickness) <= 1e-12 and rel_l2(many.ssh, one.ssh) <= 1e-9
    # (ssh = column - 1000 cancels ten digits: 1e-9 on ssh is 1e-12 on the column)


def test_multilevel_states_refuse_what_is_single_level_only(backend):
    m, ssh, uk, hk = _column_case(16, 2)
    p = mb.PrognosticVars(ssh, uk, hk, 2, mb.Mesh(m, backend))
    from moka_b200 import _lib as L
    L.check(L.lib().mokab_tape_begin(p.dev.handle, 1))                 # the multi-level reverse mode is RungeKutta4:
    with pytest.raises(mb.MokaError, match="multi-level reverse mode is RungeKutta4"):
        mb.ocn_timestep(mb.cfl_dt(m["dc"]), p, None, None, None, mb.ForwardEuler)          # ... ForwardEuler steps cannot be recorded
    with pytest.raises(mb.MokaError, match="single-level"):
        L.check(L.lib().mokab_adjoint_forward_euler(p.dev.handle))
    with pytest.raises(mb.MokaError, match="Float64"):
        mb.PrognosticVars(ssh.astype(np.float32), uk.astype(np.float32), hk.astype(np.float32), 2, mb.Mesh(m, backend))


@pytest.mark.parametrize("K,nx", [(3, 24), (10, 16), (1, 16)])
def test_multilevel_reverse_mode_matches_the_level_axis_adjoint_oracle(backend, K, nx):
    """`autodiff(Reverse, ocn_run_loop, ...)` of J = sum ssh^2 on a K-level state (RungeKutta4): the forward recompute is the
    column kernel, every adjoint stage the single-level gather kernel per level with the pressure term taken from the level
    sum of kbar_u (csrc/moka_b200.cu: adjoint_step_ml) -- against oracle/adjoint_oracle.py: gradient_sum_ssh2_levels, which is
    pinned by finite differences of the multi-level oracle (tests/test_adjoint_oracle.py)."""
    import adjoint_oracle as AO
    m, ssh, uk, hk = _column_case(nx, K, seed=5)
    dt = mb.cfl_dt(m["dc"])
    nsteps = 4
    mesh = mb.Mesh(m, backend)
    prog = mb.PrognosticVars(ssh, uk if K > 1 else uk[:, 0], hk if K > 1 else hk[:, 0], 2, mesh)
    d_prog = mb.ocn_init_shadows(prog)
    J = mb.autodiff_reverse_run_loop(dt, prog, d_prog, None, None, None, mb.RungeKutta4, nsteps)
    Jo, gu, gh = AO.gradient_sum_ssh2_levels(m, np.ascontiguousarray(uk.T), np.ascontiguousarray(hk.T), dt, nsteps)
    assert abs(J - Jo) <= 1e-12 * Jo
    du = np.asarray(d_prog.normalVelocity).reshape(m["nEdges"], K)
    dh = np.asarray(d_prog.layerThickness).reshape(m["nCells"], K)
    assert rel_l2(du.T, gu) <= 1e-12 and rel_l2(dh.T, gh) <= 1e-12
    # the forward run under the tape is the ordinary run, bit for bit
    ref = mb.PrognosticVars(ssh, uk if K > 1 else uk[:, 0], hk if K > 1 else hk[:, 0], 2, mesh)
    mb.ocn_timestep(dt, ref, None, None, None, mb.RungeKutta4, nsteps=nsteps)
    assert np.array_equal(prog.normalVelocity, ref.normalVelocity) and np.array_equal(prog.ssh, ref.ssh)
