"""GPU: the hand-written reverse mode of the fused RK4 path (csrc/kernels_adjoint.cuh) against the adjoint
oracle (scatter form, oracle/adjoint_oracle.py) and against the reference's own acceptance criterion for
gradients -- central finite differences of J = sum ssh^2 (test/enzyme/test_Enzyme_end2end.jl:112-180)."""
import numpy as np
import pytest

import adjoint_oracle as A
import moka_b200 as mb
import moka_oracle_c as OC
from conftest import hex_mesh, rel_l2

pytestmark = pytest.mark.gpu
TOL64, TOL32 = 1e-12, 1e-3


def _case(nx, kelvin):
    if kelvin:
        m = mb.channel_hex(nx, nx, 1.0e7 / nx)
        ssh, u, h = mb.kelvinWave(m).initial_state()
        mo = OC.apply_boundary_mask(m)
        OC.sign_index_fields(mo)
    else:
        m = mo = hex_mesh(nx)
        ssh, u, h = mb.inertialGravityWave(m).initial_state()
    return m, mo, ssh, u, h, mb.cfl_dt(m["dc"])


@pytest.mark.parametrize("kelvin", [False, True])
@pytest.mark.parametrize("renumber", [True, False])
def test_gradient_of_sum_ssh2_matches_adjoint_oracle(backend, kelvin, renumber):
    m, mo, ssh, u, h, dt = _case(24, kelvin)
    nsteps = 6
    mesh = mb.Mesh(m, backend, renumber=renumber)
    prog = mb.PrognosticVars(ssh, u, h, 2, mesh)
    d_prog = mb.ocn_init_shadows(prog)
    J = mb.autodiff_reverse_run_loop(dt, prog, d_prog, None, None, None, mb.RungeKutta4, nsteps)
    Jo, gu, gh = A.gradient_sum_ssh2(mo, u, h, dt, nsteps)
    assert abs(J - Jo) <= 1e-12 * Jo
    du, dh = d_prog.normalVelocity, d_prog.layerThickness
    assert rel_l2(du, gu) <= TOL64 and rel_l2(dh, gh) <= TOL64
    assert np.all(d_prog.ssh == 0)                # folded into d_layerThickness
    # the forward state is untouched by the reverse sweep
    prog2 = mb.PrognosticVars(ssh, u, h, 2, mesh)
    mb.ocn_timestep(dt, prog2, None, None, None, mb.RungeKutta4, nsteps=nsteps)
    assert np.array_equal(prog.layerThickness, prog2.layerThickness) and np.array_equal(prog.normalVelocity, prog2.normalVelocity)


def test_gradient_matches_central_finite_differences_on_the_gpu(backend):
    """The reference's check (test_Enzyme_end2end.jl:112-180): perturb one layerThickness / normalVelocity entry by a
    relative eps, difference J of two forward runs -- both runs on the GPU here."""
    m, mo, ssh, u, h, dt = _case(16, False)
    nsteps = 5
    mesh = mb.Mesh(m, backend)
    prog = mb.PrognosticVars(ssh, u, h, 2, mesh)
    d_prog = mb.ocn_init_shadows(prog)
    mb.autodiff_reverse_run_loop(dt, prog, d_prog, None, None, None, mb.RungeKutta4, nsteps)
    du, dh = d_prog.normalVelocity, d_prog.layerThickness

    def J_of(u_, h_):
        p = mb.PrognosticVars(h_ - 1000.0, u_, h_, 2, mesh)
        return mb.ocn_run_loop(dt, p, None, None, None, mb.RungeKutta4, nsteps, sum_ssh2=True)

    for k in (4, 100):                                                   # the reference checks index 5 (1-based)
        hp, hm = h.copy(), h.copy()
        hp[k] += abs(h[k]) * 1e-7
        hm[k] -= abs(h[k]) * 1e-7
        fd = (J_of(u, hp) - J_of(u, hm)) / (hp[k] - hm[k])
        assert abs(dh[k] - fd) < 1e-4                                    # atol of test_Enzyme_end2end.jl:176
        up, um = u.copy(), u.copy()
        up[k] += abs(u[k]) * 1e-4
        um[k] -= abs(u[k]) * 1e-4
        fd = (J_of(up, h) - J_of(um, h)) / (up[k] - um[k])
        assert abs(du[k] - fd) < 1e-2                                    # atol of :177
        assert abs(du[k] - fd) < 1e-5 * abs(du[k]) + 1e-4


def test_user_seed_and_float32(backend):
    """A caller-provided adjoint of the final state (seed=None) and the Float32 path."""
    m, mo, ssh, u, h, dt = _case(16, False)
    rng = np.random.default_rng(5)
    lu, lh = rng.standard_normal(m["nEdges"]), rng.standard_normal(m["nCells"])
    want_u, want_h = lu.copy(), lh.copy()
    traj = A.run_forward(mo, u, h, dt, 3)
    for n in (2, 1, 0):
        want_u, want_h = A.rk4_step_vjp(mo, traj[n][0], traj[n][1], dt, want_u, want_h)
    for dtype, tol in ((np.float64, TOL64), (np.float32, TOL32)):
        mesh = mb.Mesh(m, backend)
        prog = mb.PrognosticVars(ssh.astype(dtype), u.astype(dtype), h.astype(dtype), 2, mesh)
        d_prog = mb.ocn_init_shadows(prog)
        d_prog.normalVelocity, d_prog.layerThickness = lu.astype(dtype), lh.astype(dtype)
        mb.autodiff_reverse_run_loop(dt, prog, d_prog, None, None, None, mb.RungeKutta4, 3, seed=None)
        assert rel_l2(d_prog.normalVelocity, want_u) <= tol and rel_l2(d_prog.layerThickness, want_h) <= tol


def test_tape_overflow_and_stepper_errors(backend):
    m, mo, ssh, u, h, dt = _case(8, False)
    mesh = mb.Mesh(m, backend)
    prog = mb.PrognosticVars(ssh, u, h, 2, mesh)
    d_prog = mb.ocn_init_shadows(prog)
    with pytest.raises(mb.MokaError, match="unknown stepper"):
        mb.autodiff_reverse_run_loop(dt, prog, d_prog, None, None, None, object, 2)
    from moka_b200 import _lib as L
    L.check(L.lib().mokab_tape_begin(prog.dev.handle, 2))
    with pytest.raises(mb.MokaError, match="tape is full"):
        mb.ocn_timestep(dt, prog, None, None, None, mb.RungeKutta4, nsteps=3)
    # a tape holds steps of one stepper only, and each reverse sweep refuses the other's tape
    L.check(L.lib().mokab_tape_begin(prog.dev.handle, 4))
    mb.ocn_timestep(dt, prog, None, None, None, mb.RungeKutta4, nsteps=1)
    with pytest.raises(mb.MokaError, match="already holds RungeKutta4"):
        mb.ocn_timestep(dt, prog, None, None, None, mb.ForwardEuler)
    with pytest.raises(mb.MokaError, match="holds RungeKutta4"):
        L.check(L.lib().mokab_adjoint_forward_euler(prog.dev.handle))
    L.check(L.lib().mokab_tape_begin(prog.dev.handle, 4))
    mb.ocn_timestep(dt, prog, None, None, None, mb.ForwardEuler)
    with pytest.raises(mb.MokaError, match="already holds ForwardEuler"):
        mb.ocn_timestep(dt, prog, None, None, None, mb.RungeKutta4, nsteps=1)
    with pytest.raises(mb.MokaError, match="holds ForwardEuler"):
        L.check(L.lib().mokab_adjoint_rk4(prog.dev.handle))
    with pytest.raises(mb.MokaError, match="tape is full"):
        for _ in range(4):
            mb.ocn_timestep(dt, prog, None, None, None, mb.ForwardEuler)


def test_operator_adjoints_like_test_Enzyme_Operators(backend):
    """test/enzyme/test_Enzyme_Operators.jl: reverse mode of GradientOnEdge! (seed d_gradNum[1] = 1, read d_Scalar[1],
    :47-64) and of DivergenceOnCell! (seed d_divNum[1] = 1, read d_VecEdge[2], :130-152) against central finite
    differences with a relative step of 1e-8 (atol 1e-6, :121 and :221), on the 48x48 planar test mesh and fields."""
    import moka_oracle as O
    m = hex_mesh(48, 48, 1000.0)
    mesh = mb.Mesh(m, backend)
    f = O.planar_test_fields(m)
    eps = 1e-8
    # gradient: input index kBegin = 1, output index kEnd = 1 (1-based in the reference)
    seed = np.zeros(m["nEdges"]); seed[0] = 1.0
    rev = mb.GradientOnEdge_vjp(seed, mesh)[0]
    sp, sm = f["h"].copy(), f["h"].copy()
    sp[0] += abs(sp[0]) * eps
    sm[0] -= abs(sm[0]) * eps
    fd = (mb.GradientOnEdge(None, sp, mesh)[0] - mb.GradientOnEdge(None, sm, mesh)[0]) / (sp[0] - sm[0])
    assert abs(rev - fd) < 1e-6
    # divergence: input index kBegin = 2, output index kEnd = 1
    seed = np.zeros(m["nCells"]); seed[0] = 1.0
    rev = mb.DivergenceOnCell_vjp(seed, mesh)[1]
    vp, vm = f["F_edge"].copy(), f["F_edge"].copy()
    vp[1] += abs(vp[1]) * eps
    vm[1] -= abs(vm[1]) * eps
    fd = (mb.DivergenceOnCell(None, vp, None, mesh)[0] - mb.DivergenceOnCell(None, vm, None, mesh)[0]) / (vp[1] - vm[1])
    assert abs(rev - fd) < 1e-6
    # and the full transposes: <A x, y> == <x, A^T y> for random x, y
    rng = np.random.default_rng(7)
    x, y = rng.standard_normal(m["nCells"]), rng.standard_normal(m["nEdges"])
    lhs, rhs = mb.GradientOnEdge(None, x, mesh) @ y, x @ mb.GradientOnEdge_vjp(y, mesh)
    assert abs(lhs - rhs) <= 1e-12 * abs(lhs)
    lhs, rhs = mb.DivergenceOnCell(None, y, None, mesh) @ x, y @ mb.DivergenceOnCell_vjp(x, mesh)
    assert abs(lhs - rhs) <= 1e-12 * abs(lhs)


def test_committed_adjoint_fixture(backend):
    """tests/golden/igw16_adjoint.npz (made by tests/golden/make_golden_adjoint.py from the adjoint oracle)."""
    import json
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "igw16_adjoint.npz"))
    meta = json.loads(str(g["meta"]))
    m = hex_mesh(16)
    ssh, u, h = mb.inertialGravityWave(m).initial_state()
    mesh = mb.Mesh(m, backend)
    prog = mb.PrognosticVars(ssh, u, h, 2, mesh)
    d_prog = mb.ocn_init_shadows(prog)
    J = mb.autodiff_reverse_run_loop(meta["dt"], prog, d_prog, None, None, None, mb.RungeKutta4, meta["nsteps"])
    assert abs(J - float(g["J"])) <= 1e-12 * J
    assert rel_l2(d_prog.normalVelocity, g["d_normalVelocity"]) <= TOL64 and rel_l2(d_prog.layerThickness, g["d_layerThickness"]) <= TOL64
    k = meta["fd_index"]
    assert abs(d_prog.layerThickness[k] - float(g["fd_layerThickness"])) < 1e-4          # test_Enzyme_end2end.jl:176
    assert abs(d_prog.normalVelocity[k] - float(g["fd_normalVelocity"])) < 1e-2          # :177


# ---- ForwardEuler: the stepper the reference differentiates (test_Enzyme_end2end.jl:78-96) ---------------------------
@pytest.mark.parametrize("kelvin,renumber,nx", [(False, True, 16), (False, False, 12), (True, True, 16)])
def test_forward_euler_gradient_matches_adjoint_oracle(backend, kelvin, renumber, nx):
    m, mo, ssh, u, h, dt = _case(nx, kelvin)
    mesh = mb.Mesh(m, backend, renumber=renumber)
    prog = mb.PrognosticVars(ssh, u, h, 2, mesh)
    d_prog = mb.ocn_init_shadows(prog)
    J = mb.autodiff_reverse_run_loop(dt, prog, d_prog, None, None, None, mb.ForwardEuler, 7)
    Jo, gu, gh, gs, _ = A.gradient_sum_ssh2_fe(mo, ssh, u, h, dt, 7)
    assert abs(J - Jo) <= 1e-12 * abs(Jo)
    assert rel_l2(d_prog.normalVelocity, gu) <= TOL64
    assert rel_l2(d_prog.layerThickness, gh) <= TOL64
    assert rel_l2(d_prog.ssh, gs) <= TOL64
    # the forward run under the tape is the ordinary ForwardEuler run, bit for bit
    om = OC.OracleModel(mo, ssh, u, h)
    om.run_loop(dt, 7, "ForwardEuler")
    assert np.array_equal(prog.normalVelocity, om.normalVelocity[1]) and np.array_equal(prog.ssh, om.ssh[1])


def test_forward_euler_gradient_like_test_Enzyme_end2end(backend):
    """The reference's own acceptance test (test_Enzyme_end2end.jl:112-180): AD value at one cell / edge against central
    finite differences of the forward run, atol 1e-4 (layerThickness) and 1e-2 (normalVelocity) -- its CUDA result is NaN."""
    m, mo, ssh, u, h, dt = _case(16, False)
    mesh = mb.Mesh(m, backend)
    nsteps, k = 6, 4                                                     # the reference checks index 5 (1-based)

    def J_of(u0, h0):
        p = mb.PrognosticVars(ssh, u0, h0, 2, mesh)
        mb.ocn_run_loop(dt, p, None, None, None, mb.ForwardEuler, nsteps)
        return mb.reduce_sum(p, "ssh2")

    prog = mb.PrognosticVars(ssh, u, h, 2, mesh)
    d_prog = mb.ocn_init_shadows(prog)
    mb.autodiff_reverse_run_loop(dt, prog, d_prog, None, None, None, mb.ForwardEuler, nsteps)
    gh, gu = d_prog.layerThickness, d_prog.normalVelocity
    assert np.all(np.isfinite(gh)) and np.all(np.isfinite(gu))
    hp, hm = h.copy(), h.copy()
    hp[k] += abs(h[k]) * 1e-7
    hm[k] -= abs(h[k]) * 1e-7
    fd_h = (J_of(u, hp) - J_of(u, hm)) / (hp[k] - hm[k])
    up, um = u.copy(), u.copy()
    up[k] += abs(u[k]) * 1e-4
    um[k] -= abs(u[k]) * 1e-4
    fd_u = (J_of(up, h) - J_of(um, h)) / (up[k] - um[k])
    assert abs(gh[k] - fd_h) < 1e-4 and abs(gu[k] - fd_u) < 1e-2
    assert abs(gh[k] - fd_h) < 1e-5 * abs(gh[k]) + 1e-7


def test_forward_euler_adjoint_on_runtime_width_rows(backend):
    """Padded rows with MOKAB_MESH_KEEP_WIDTHS: the unfused ForwardEuler steps record the tape and the run-time-width
    adjoint kernel reverses them."""
    from test_gpu_parity import _padded
    m, mo, ssh, u, h, dt = _case(12, False)
    mesh = mb.Mesh(_padded(m), backend, keep_widths=True)
    prog = mb.PrognosticVars(ssh, u, h, 2, mesh)
    d_prog = mb.ocn_init_shadows(prog)
    mb.autodiff_reverse_run_loop(dt, prog, d_prog, None, None, None, mb.ForwardEuler, 5)
    _, gu, gh, gs, _ = A.gradient_sum_ssh2_fe(mo, ssh, u, h, dt, 5)
    assert rel_l2(d_prog.normalVelocity, gu) <= TOL64 and rel_l2(d_prog.layerThickness, gh) <= TOL64
    assert rel_l2(d_prog.ssh, gs) <= TOL64


def test_committed_forward_euler_adjoint_fixture(backend):
    """tests/golden/igw16_adjoint_fe.npz (made by tests/golden/make_golden_adjoint.py from the adjoint oracle)."""
    import json
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "igw16_adjoint_fe.npz"))
    meta = json.loads(str(g["meta"]))
    m = hex_mesh(16)
    ssh, u, h = mb.inertialGravityWave(m).initial_state()
    prog = mb.PrognosticVars(ssh, u, h, 2, mb.Mesh(m, backend))
    d_prog = mb.ocn_init_shadows(prog)
    J = mb.autodiff_reverse_run_loop(meta["dt"], prog, d_prog, None, None, None, mb.ForwardEuler, meta["nsteps"])
    assert abs(J - float(g["J"])) <= 1e-12 * J
    assert rel_l2(d_prog.normalVelocity, g["d_normalVelocity"]) <= TOL64 and rel_l2(d_prog.layerThickness, g["d_layerThickness"]) <= TOL64
    assert rel_l2(d_prog.ssh, g["d_ssh"]) <= TOL64
    k = meta["fd_index"]
    assert abs(d_prog.layerThickness[k] - float(g["fd_layerThickness"])) < 1e-4          # test_Enzyme_end2end.jl:176
    assert abs(d_prog.normalVelocity[k] - float(g["fd_normalVelocity"])) < 1e-2          # :177


def test_reverse_sweep_of_an_empty_tape_returns_the_seed(backend):
    """nsteps = 0: the gradient of J = sum ssh^2 at the initial state itself.  RungeKutta4 defines ssh = h - H, so dJ/dh =
    2 (h - H); for ForwardEuler ssh is an input of its own: dJ/dssh = 2 ssh, dJ/dh = 0.  dJ/du = 0 either way."""
    m, mo, ssh, u, h, dt = _case(12, False)
    ssh = ssh + 0.25                      # ssh != h - H: the two definitions must give different answers
    mesh = mb.Mesh(m, backend)
    prog = mb.PrognosticVars(ssh, u, h, 2, mesh)
    d = mb.ocn_init_shadows(prog)
    mb.autodiff_reverse_run_loop(dt, prog, d, None, None, None, mb.RungeKutta4, 0)
    _, gu, gh = A.gradient_sum_ssh2(mo, u, h, dt, 0)
    assert np.array_equal(d.normalVelocity, gu) and not gu.any()
    assert rel_l2(d.layerThickness, gh) <= TOL64
    prog = mb.PrognosticVars(ssh, u, h, 2, mesh)
    d = mb.ocn_init_shadows(prog)
    J = mb.autodiff_reverse_run_loop(dt, prog, d, None, None, None, mb.ForwardEuler, 0)
    Jo, gu, gh, gs, _ = A.gradient_sum_ssh2_fe(mo, ssh, u, h, dt, 0)
    assert abs(J - Jo) <= 1e-12 * Jo
    assert not np.asarray(d.normalVelocity).any() and not np.asarray(d.layerThickness).any() and not gh.any()
    assert np.array_equal(d.ssh, gs) and np.array_equal(gs, 2.0 * ssh)
