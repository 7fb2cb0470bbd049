"""GPU: the domain-decomposed path.  (1) P ranks emulated in ONE process on one GPU (messages moved
with device copies instead of NCCL) -- checks owned/halo renumbering, interior/boundary block parts,
halo pack/unpack and the staged RK4 entry points against the single-domain oracle; (2) the real
one-process-per-GPU NCCL path under torchrun when the box has >= 2 GPUs."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

import moka_b200 as mb
import moka_oracle_c as OC
from conftest import DevBuf, device_synchronize, hex_mesh, rel_l2
from moka_b200 import _lib as L
from moka_b200 import multi_gpu, partition

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _Rank:
    def __init__(self, backend, loc, state, nparts, dtype=np.float64):
        self.loc = loc
        self.mesh = mb.Mesh(loc, backend)
        self.sidx, self.scnt, self.ridx, self.rcnt = partition.flat_halo(loc, nparts)
        self.mesh.halo_setup(self.sidx, self.ridx)
        ssh, u, h = multi_gpu.local_state(loc, *state)
        self.prog = mb.PrognosticVars(ssh.astype(dtype), u.astype(dtype), h.astype(dtype), 2, self.mesh)
        self.send = DevBuf(len(self.sidx), dtype)
        self.recv = DevBuf(len(self.ridx), dtype)
        self.h = self.prog.dev.handle


def _setup_p2p(ranks, nparts):
    """The set-up of the direct-store halo exchange (DecomposedModel._setup_p2p) for ranks emulated in one process."""
    lib = L.lib()
    recv_dev = []
    for r in ranks:
        a = np.zeros(max(1, len(r.ridx)), np.int32)
        L.check(lib.mokab_halo_recv_device_indices(r.mesh.handle, a.ctypes.data_as(L._I32P)))
        recv_dev.append(a)
    size = C.c_int64()
    L.check(lib.mokab_p2p_blob_size(C.byref(size)))
    blobs = []
    for r in ranks:
        b = C.create_string_buffer(size.value)
        L.check(lib.mokab_p2p_export(r.h, r.loc["rank"], b))
        blobs.append(b.raw)
    blobs = b"".join(blobs)
    for r in ranks:
        me = r.loc["rank"]
        receivers = senders = [q for q in range(nparts) if r.scnt[q] > 0 or r.rcnt[q] > 0]   # symmetric (kernels_p2p.cuh)
        dst = []
        for q in receivers:                      # rank q's halo segment filled by me, in q's own (device) numbering
            o = sum(ranks[q].rcnt[:me])
            dst.append(recv_dev[q][o:o + ranks[q].rcnt[me]])
        dst = np.ascontiguousarray(np.concatenate(dst + [np.zeros(0, np.int32)]), np.int32)
        rr, sr = np.asarray(receivers, np.int32), np.asarray(senders, np.int32)
        cnt = np.asarray([r.scnt[q] for q in receivers], np.int64)
        L.check(lib.mokab_p2p_setup(r.h, me, nparts, blobs, len(receivers), rr.ctypes.data_as(L._I32P),
                                    cnt.ctypes.data_as(C.POINTER(C.c_int64)), dst.ctypes.data_as(L._I32P), len(senders),
                                    sr.ctypes.data_as(L._I32P)))


def _run_emulated(backend, m, state, nparts, dt, nsteps, split_parts=True, dtype=np.float64, halo="nccl"):
    locs = partition.decompose(m, nparts)
    ranks = [_Rank(backend, loc, state, nparts, dtype) for loc in locs]
    lib = L.lib()
    if halo != "nccl":
        _setup_p2p(ranks, nparts)
    for _ in range(nsteps if halo == "p2p_fused" else 0):
        for s in (1, 2, 3, 4):                   # the boundary launch waits, computes, stores into the neighbours and ticks them
            for r in ranks:
                L.check(lib.mokab_rk4_stage(r.h, dt, s, L.PART_BOUNDARY_PUSH, None))
                L.check(lib.mokab_rk4_stage(r.h, dt, s, L.PART_INTERIOR, None))
        for r in ranks:
            L.check(lib.mokab_rk4_finish_step(r.h))
    if halo == "p2p_fused":
        for r in ranks:
            L.check(lib.mokab_halo_wait_arrivals(r.h, None))
    for _ in range(nsteps if halo == "p2p" else 0):
        for s in (1, 2, 3, 4):
            for r in ranks:                      # every rank's stores (and arrival ticks) are enqueued before anybody waits
                L.check(lib.mokab_rk4_stage(r.h, dt, s, L.PART_BOUNDARY if split_parts else L.PART_ALL, None))
                L.check(lib.mokab_halo_push(r.h, s, None))
                if split_parts:
                    L.check(lib.mokab_rk4_stage(r.h, dt, s, L.PART_INTERIOR, None))
            for r in ranks:
                L.check(lib.mokab_halo_wait(r.h, None))
        for r in ranks:
            L.check(lib.mokab_rk4_finish_step(r.h))
    for _ in range(nsteps if halo == "nccl" else 0):
        for s in (1, 2, 3, 4):
            for r in ranks:
                if split_parts:
                    L.check(lib.mokab_rk4_stage(r.h, dt, s, L.PART_BOUNDARY, None))
                    L.check(lib.mokab_halo_pack(r.h, s, C.c_void_p(r.send.data_ptr()), None))
                    L.check(lib.mokab_rk4_stage(r.h, dt, s, L.PART_INTERIOR, None))
                else:
                    L.check(lib.mokab_rk4_stage(r.h, dt, s, L.PART_ALL, None))
                    L.check(lib.mokab_halo_pack(r.h, s, C.c_void_p(r.send.data_ptr()), None))
            backend.synchronize()
            for r in ranks:                      # the "all-to-all": recv segment from q <- q's send segment to r
                ro = 0
                for q, cr in enumerate(r.rcnt):
                    if cr:
                        so = sum(ranks[q].scnt[:r.loc["rank"]])
                        r.recv.copy_from(ro, ranks[q].send, so, cr)
                    ro += cr
            device_synchronize()
            for r in ranks:
                L.check(lib.mokab_halo_unpack(r.h, s, C.c_void_p(r.recv.data_ptr()), None))
        for r in ranks:
            L.check(lib.mokab_rk4_finish_step(r.h))
    gu, gh, gs = np.full(m["nEdges"], np.nan), np.full(m["nCells"], np.nan), np.full(m["nCells"], np.nan)
    for r in ranks:
        if halo != "nccl":
            err = C.c_int(1)
            L.check(lib.mokab_p2p_error(r.h, C.byref(err)))
            assert err.value == 0, "a halo wait timed out"
        L.check(lib.mokab_refresh_ssh(r.h, None))
        no, ne = r.loc["nCellsOwned"], r.loc["nEdgesOwned"]
        gu[r.loc["edgesGlobal"][:ne]] = r.prog.normalVelocity[:ne]
        gh[r.loc["cellsGlobal"][:no]] = r.prog.layerThickness[:no]
        gs[r.loc["cellsGlobal"][:no]] = r.prog.ssh[:no]
    return gu, gh, gs, ranks


@pytest.mark.parametrize("nx,nparts,split", [(32, 2, True), (48, 4, True), (64, 8, True), (64, 3, False)])
def test_emulated_ranks_match_oracle(backend, nx, nparts, split):
    m = hex_mesh(nx)
    state = mb.inertialGravityWave(m).initial_state()
    dt = mb.cfl_dt(m["dc"])
    gu, gh, gs, ranks = _run_emulated(backend, m, state, nparts, dt, 12, split_parts=split)
    om = OC.OracleModel(m, *state)
    om.run_loop(dt, 12, "RungeKutta4")
    assert rel_l2(gu, om.normalVelocity[1]) <= 1e-12
    assert rel_l2(gh, om.layerThickness[1]) <= 1e-12
    assert rel_l2(gs, om.ssh[1]) <= 1e-12
    # and identical to the single-domain GPU run (same kernel, same per-entity arithmetic)
    mesh = mb.Mesh(m, backend)
    prog = mb.PrognosticVars(*state, 2, mesh)
    mb.ocn_run_loop(dt, prog, None, None, None, mb.RungeKutta4, 12)
    assert np.array_equal(gu, prog.normalVelocity) and np.array_equal(gh, prog.layerThickness)
    for r in ranks:
        ni, nb = r.mesh.block_counts()
        assert ni + nb == r.mesh.derived_blocks()[0] and nb >= 1        # every block of owned cells is in exactly one part


@pytest.mark.parametrize("nx,nparts,split,dtype", [(32, 2, True, np.float64), (96, 8, True, np.float64), (64, 3, False, np.float64),
                                                   (48, 4, True, np.float32)])
def test_direct_store_halo_exchange_matches_the_packed_one(backend, nx, nparts, split, dtype):
    """csrc/kernels_p2p.cuh with the ranks emulated in one process (peer pointers = plain addresses): the same bits as the
    pack / copy / unpack path, which the test above ties to the oracle."""
    m = hex_mesh(nx, with_dual=False)
    state = mb.inertialGravityWave(m).initial_state()
    dt = mb.cfl_dt(m["dc"])
    gu0, gh0, gs0, _ = _run_emulated(backend, m, state, nparts, dt, 7, split_parts=split, dtype=dtype)
    gu1, gh1, gs1, _ = _run_emulated(backend, m, state, nparts, dt, 7, split_parts=split, dtype=dtype, halo="p2p")
    assert np.array_equal(gu0, gu1) and np.array_equal(gh0, gh1) and np.array_equal(gs0, gs1)
    if split:                                    # and with the exchange folded into the boundary launch
        gu2, gh2, gs2, _ = _run_emulated(backend, m, state, nparts, dt, 7, dtype=dtype, halo="p2p_fused")
        assert np.array_equal(gu0, gu2) and np.array_equal(gh0, gh2) and np.array_equal(gs0, gs2)
    if dtype == np.float64:
        om = OC.OracleModel(m, *state)
        om.run_loop(dt, 7, "RungeKutta4")
        assert np.array_equal(gu1, om.normalVelocity[1]) and np.array_equal(gh1, om.layerThickness[1])


def test_block_parts_at_scale(backend):
    """512x512 over 4 ranks: most blocks are interior; decomposed == single-domain result."""
    m = hex_mesh(512, with_dual=False)
    state = mb.inertialGravityWave(m).initial_state()
    dt = mb.cfl_dt(m["dc"])
    gu, gh, gs, ranks = _run_emulated(backend, m, state, 4, dt, 4)
    mesh = mb.Mesh(m, backend)
    prog = mb.PrognosticVars(*state, 2, mesh)
    mb.ocn_run_loop(dt, prog, None, None, None, mb.RungeKutta4, 4)
    assert np.array_equal(gu, prog.normalVelocity) and np.array_equal(gh, prog.layerThickness)
    for r in ranks:
        ni, nb = r.mesh.block_counts()
        assert ni > 3 * nb, (ni, nb)
    mass = sum(mb.reduce_sum(r.prog, "mass") for r in ranks)
    assert abs(mass - mb.reduce_sum(prog, "mass")) <= 1e-13 * mass


def test_decomposed_state_refuses_single_domain_stepper(backend):
    m = hex_mesh(16)
    loc = partition.decompose(m, 2)[0]
    mesh = mb.Mesh(loc, backend)
    ssh, u, h = multi_gpu.local_state(loc, *mb.inertialGravityWave(m).initial_state())
    prog = mb.PrognosticVars(ssh, u, h, 2, mesh)
    with pytest.raises(mb.MokaError, match="halo"):
        mb.ocn_timestep(1.0, prog, None, None, None, mb.RungeKutta4)
    with pytest.raises(mb.MokaError, match="halo"):              # nor would ForwardEuler see its neighbours' values
        mb.ocn_timestep(1.0, prog, None, None, None, mb.ForwardEuler)
    with pytest.raises(mb.MokaError, match="halo_setup"):
        L.check(L.lib().mokab_halo_pack(prog.dev.handle, 1, None, None))


def test_nccl_two_ranks_torchrun():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    out = subprocess.run(["timeout", "300", sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29611",
                          os.path.join(ROOT, "tests", "multi_gpu_check.py")], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "MULTI_GPU_CHECK_OK" in out.stdout


def _run_emulated_fe(backend, m, state, nparts, dt, nsteps, split_parts=True):
    """ForwardEuler over ranks emulated in one process: mokab_forward_euler_stage per part, the two exchanges of a step
    (stage 4: (h, u)[new]; stage 5: (ssh, layerThicknessEdge)[new]) as device copies, then the finish call."""
    locs = partition.decompose(m, nparts)
    ranks = [_Rank(backend, loc, state, nparts) for loc in locs]
    lib = L.lib()
    for _ in range(nsteps):
        for r in ranks:
            if split_parts:
                L.check(lib.mokab_forward_euler_stage(r.h, dt, L.PART_BOUNDARY, None))
                L.check(lib.mokab_forward_euler_stage(r.h, dt, L.PART_INTERIOR, None))
            else:
                L.check(lib.mokab_forward_euler_stage(r.h, dt, L.PART_ALL, None))
        for s in (4, 5):
            for r in ranks:
                L.check(lib.mokab_halo_pack(r.h, s, C.c_void_p(r.send.data_ptr()), None))
            backend.synchronize()
            for r in ranks:
                ro = 0
                for q, cr in enumerate(r.rcnt):
                    if cr:
                        so = sum(ranks[q].scnt[:r.loc["rank"]])
                        r.recv.copy_from(ro, ranks[q].send, so, cr)
                    ro += cr
            device_synchronize()
            for r in ranks:
                L.check(lib.mokab_halo_unpack(r.h, s, C.c_void_p(r.recv.data_ptr()), None))
        for r in ranks:
            L.check(lib.mokab_forward_euler_finish_step(r.h))
    gu, gh, gs = np.full(m["nEdges"], np.nan), np.full(m["nCells"], np.nan), np.full(m["nCells"], np.nan)
    for r in ranks:
        no, ne = r.loc["nCellsOwned"], r.loc["nEdgesOwned"]
        gu[r.loc["edgesGlobal"][:ne]] = r.prog.normalVelocity[:ne]
        gh[r.loc["cellsGlobal"][:no]] = r.prog.layerThickness[:no]
        gs[r.loc["cellsGlobal"][:no]] = r.prog.ssh[:no]
    return gu, gh, gs


@pytest.mark.parametrize("nx,nparts,split,kelvin", [(32, 2, True, False), (48, 8, True, False), (40, 3, False, False), (32, 4, True, True)])
def test_staged_forward_euler_on_emulated_ranks_is_the_reference_sequence(backend, nx, nparts, split, kelvin):
    """The reference's live stepper on a decomposed mesh: bit for bit the oracle's ForwardEuler (first-step quirk and lagged
    layerThicknessEdge included), which is what the single-domain fused step gives."""
    if kelvin:
        m = mb.channel_hex(nx, nx, 1.0e7 / nx)
        state = mb.kelvinWave(m).initial_state()
        mo = OC.apply_boundary_mask(m)
    else:
        m = mo = hex_mesh(nx)
        state = mb.inertialGravityWave(m).initial_state()
    OC.sign_index_fields(mo)
    dt, nsteps = mb.cfl_dt(m["dc"]), 7
    om = OC.OracleModel(mo, *state)
    om.run_loop(dt, nsteps, "ForwardEuler")
    md = {k: v for k, v in m.items() if k not in ("edgesOnVertex", "cellsOnVertex", "verticesOnEdge", "kiteAreasOnVertex",
                                                  "areaTriangle", "verticesOnCell", "edgeSignOnVertex")}
    md["nVertices"] = 0
    gu, gh, gs = _run_emulated_fe(backend, md, state, nparts, dt, nsteps, split_parts=split)
    assert np.array_equal(gu, om.normalVelocity[1])
    assert np.array_equal(gh, om.layerThickness[1])
    assert np.array_equal(gs, om.ssh[1])


@pytest.mark.parametrize("parts", [(L.PART_ALL,), (L.PART_BOUNDARY, L.PART_INTERIOR)])
def test_staged_forward_euler_on_an_undecomposed_mesh_interoperates_with_the_whole_mesh_entry_point(backend, parts):
    """No halo at all: five staged steps are five ocn_timestep(ForwardEuler) steps, and three more through the whole-mesh entry
    point continue the same trajectory (the lagged layerThicknessEdge is handed over).  Meshes with vertex arrays are refused
    (the staged step does not advance relativeVorticity), and so is a finish call without a stage."""
    m = hex_mesh(20)
    ssh, u, h = mb.inertialGravityWave(m).initial_state()
    dt = mb.cfl_dt(m["dc"])
    lib = L.lib()
    ref5 = mb.PrognosticVars(ssh, u, h, 2, mb.Mesh(m, backend))
    mb.ocn_timestep(dt, ref5, None, None, None, mb.ForwardEuler, nsteps=5)
    ref8 = mb.PrognosticVars(ssh, u, h, 2, mb.Mesh(m, backend))
    mb.ocn_timestep(dt, ref8, None, None, None, mb.ForwardEuler, nsteps=8)
    with_vertices = mb.PrognosticVars(ssh, u, h, 2, mb.Mesh(m, backend))
    with pytest.raises(mb.MokaError, match="carry no vertices"):
        L.check(lib.mokab_forward_euler_stage(with_vertices.dev.handle, dt, L.PART_ALL, None))
    md = {k: v for k, v in m.items() if k not in ("edgesOnVertex", "cellsOnVertex", "verticesOnEdge", "kiteAreasOnVertex",
                                                  "areaTriangle", "verticesOnCell", "edgeSignOnVertex")}
    md["nVertices"] = 0
    prog = mb.PrognosticVars(ssh, u, h, 2, mb.Mesh(md, backend))
    with pytest.raises(mb.MokaError, match="no staged ForwardEuler step has run"):
        L.check(lib.mokab_forward_euler_finish_step(prog.dev.handle))
    for _ in range(5):
        for part in parts:
            L.check(lib.mokab_forward_euler_stage(prog.dev.handle, dt, part, None))
        L.check(lib.mokab_forward_euler_finish_step(prog.dev.handle))
    assert np.array_equal(prog.normalVelocity, ref5.normalVelocity) and np.array_equal(prog.layerThickness, ref5.layerThickness)
    assert np.array_equal(prog.ssh, ref5.ssh)
    mb.ocn_timestep(dt, prog, None, None, None, mb.ForwardEuler, nsteps=3)
    assert np.array_equal(prog.normalVelocity, ref8.normalVelocity) and np.array_equal(prog.layerThickness, ref8.layerThickness)


class _SoloRuntime:
    """Control plane of a one-rank "job": nothing to hand to anybody."""

    def rank_and_size(self):
        return 0, 1

    def broadcast_bytes(self, blob, src=0):
        return blob


@pytest.mark.parametrize("halo,graph,stepper", [("nccl", True, "RungeKutta4"), ("nccl", False, "RungeKutta4"), ("p2p_fused", True, "RungeKutta4"),
                                                ("nccl", True, "ForwardEuler"), ("p2p_ll", True, "RungeKutta4")])
def test_in_library_decomposed_entry_points_with_one_rank(backend, halo, graph, stepper):
    """mokab_comm_init / mokab_decomp_setup / mokab_timestep_*_decomposed (csrc/decomposed.cuh) on a communicator of ONE rank:
    the real NCCL initialisation, the halo stream fork / join, the captured 1- and 2-step graphs for both time-level parities
    -- everything but a neighbour -- on a one-GPU box; the multi-rank schedule runs under torchrun (tests/multi_gpu_check.py)
    and with emulated ranks on the simulated runtime (tests/sim/check_decomposed.py).  Same bits as the single-domain call."""
    m = hex_mesh(48, with_dual=False)
    state = mb.inertialGravityWave(m).initial_state()
    dt = mb.cfl_dt(m["dc"])
    loc = partition.decompose(m, 1)[0]
    model = multi_gpu.DecomposedModel(loc, multi_gpu.local_state(loc, *state), backend, 0, graph=graph, runtime=_SoloRuntime(), halo=halo)
    st = getattr(mb, stepper)
    for n in (3, 4, 1, 2):                       # odd counts move the time-level parity between the calls
        model.step(dt, n, stepper=st)
    model.finish()
    gs, gu, gh = model.gather(m["nCells"], m["nEdges"])
    mass = model.reduce("mass")
    status = model.graph_status
    model.close()
    om = OC.OracleModel(m, *state)
    om.run_loop(dt, 10, stepper)
    assert np.array_equal(gu, om.normalVelocity[1]) and np.array_equal(gh, om.layerThickness[1]) and np.array_equal(gs, om.ssh[1])
    assert abs(mass - float(np.sum(m["areaCell"] * om.layerThickness[1]))) <= 1e-13 * mass
    if graph and stepper == "RungeKutta4":
        assert status.startswith("validated"), status


def test_in_library_decomposed_graphs_on_the_very_first_call(backend):
    """A C-ABI caller's first call is mokab_timestep_rk4_decomposed with graphs on (no host-launched warm-up, no validation
    pass): everything the stage launches build lazily -- the interleaved weight copy of the default kernel variant included --
    must be in place before the capture starts."""
    m = hex_mesh(48, with_dual=False)
    state = mb.inertialGravityWave(m).initial_state()
    dt = mb.cfl_dt(m["dc"])
    loc = partition.decompose(m, 1)[0]
    model = multi_gpu.DecomposedModel(loc, multi_gpu.local_state(loc, *state), backend, 0, graph=True, runtime=_SoloRuntime())
    model._advance(dt, 5, False)                 # straight to the library, past DecomposedModel's validation
    model.finish()
    gs, gu, gh = model.gather(m["nCells"], m["nEdges"])
    model.close()
    om = OC.OracleModel(m, *state)
    om.run_loop(dt, 5, "RungeKutta4")
    assert np.array_equal(gu, om.normalVelocity[1]) and np.array_equal(gs, om.ssh[1])


@pytest.mark.parametrize("stepper,graph", [("RungeKutta4", True), ("RungeKutta4", False), ("ForwardEuler", True)])
def test_in_library_decomposed_reverse_mode_with_one_rank(backend, stepper, graph):
    """DecomposedModel.reverse_run_loop -- the tape recorded by mokab_timestep_*_decomposed, then mokab_adjoint_* with the halo
    copies of the reverse sweep (csrc/moka_b200.cu: halo_exchange_arrays) -- on a communicator of one rank: J and dJ/d(initial
    state) against the scatter-form adjoint oracle.  With neighbours: tests/sim/check_decomposed.py (2 - 8 emulated ranks, every
    halo path) and tests/multi_gpu_check.py under torchrun."""
    import adjoint_oracle as AO
    m = hex_mesh(32, with_dual=False)
    ssh, u, h = mb.inertialGravityWave(m).initial_state()
    dt = mb.cfl_dt(m["dc"])
    loc = partition.decompose(m, 1)[0]
    model = multi_gpu.DecomposedModel(loc, multi_gpu.local_state(loc, ssh, u, h), backend, 0, graph=graph, runtime=_SoloRuntime())
    fe = stepper == "ForwardEuler"
    J = model.reverse_run_loop(dt, 5, stepper=getattr(mb, stepper))
    model.finish()
    gu, gh = (np.array(a) for a in model.gradient())
    gs = np.array(model.gradient_ssh())
    model.close()
    inv_e, inv_c = np.argsort(loc["edgesGlobal"]), np.argsort(loc["cellsGlobal"])
    gu, gh, gs = gu[inv_e], gh[inv_c], gs[inv_c]
    if fe:
        Jo, ou, oh, os_, _ = AO.gradient_sum_ssh2_fe(m, ssh, u, h, dt, 5)
        assert rel_l2(gs, os_) <= 1e-12
    else:
        Jo, ou, oh = AO.gradient_sum_ssh2(m, u, h, dt, 5)
    assert abs(J - Jo) <= 1e-12 * Jo
    assert rel_l2(gu, ou) <= 1e-12 and rel_l2(gh, oh) <= 1e-12


@pytest.mark.parametrize("K,graph", [(3, True), (10, False)])
def test_in_library_decomposed_multilevel_with_one_rank(backend, K, graph):
    """Multi-level states through mokab_timestep_rk4_decomposed (fused::k_rk_stage_ml over the owned blocks + one K + 1 plane halo
    message per stage, csrc/moka_b200.cu: halo_exchange_levels) on a communicator of one rank: the numpy oracle with a level
    axis, bit for bit.  With neighbours: tests/sim/check_decomposed.py (3 - 8 emulated ranks)."""
    import moka_oracle as O
    m = dict(hex_mesh(32, with_dual=False))
    ssh, u, h = mb.inertialGravityWave(m).initial_state()
    frac = np.random.default_rng(K).uniform(0.5, 1.5, K)
    frac /= frac.sum()
    rest = np.outer(np.full(m["nCells"], 1000.0), frac)
    hk, uk = rest + np.outer(ssh, frac), np.outer(u, 1.0 + 0.1 * np.arange(K))
    m["restingThickness"], m["nVertLevels"] = rest, K
    dt = mb.cfl_dt(m["dc"])
    loc = partition.decompose(m, 1)[0]
    model = multi_gpu.DecomposedModel(loc, multi_gpu.local_state(loc, ssh, uk, hk), backend, 0, graph=graph, runtime=_SoloRuntime())
    for n in (3, 2):
        model.step(dt, n)
    model.finish()
    gu, gh, gs = (np.array(model.owned(f)) for f in ("normalVelocity", "layerThickness", "ssh"))
    status = model.graph_status
    model.close()
    inv_e, inv_c = np.argsort(loc["edgesGlobal"]), np.argsort(loc["cellsGlobal"])
    prog = O.new_state(m, ssh, np.ascontiguousarray(uk.T), np.ascontiguousarray(hk.T))
    for _ in range(5):
        O.timestep_rk4(m, prog, dt)
    assert np.array_equal(gu[inv_e].T, prog["normalVelocity"][-1]) and np.array_equal(gh[inv_c].T, prog["layerThickness"][-1])
    assert np.array_equal(gs[inv_c], prog["ssh"][-1])
    if graph:
        assert status.startswith("validated"), status
