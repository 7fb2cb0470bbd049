"""torchrun entry: one process per GPU, the library's own decomposed stepping (mokab_comm_init / mokab_decomp_setup /
mokab_timestep_*_decomposed: NCCL send/recv or direct peer stores inside libmoka_b200.so), parity of the gathered result
against the single-domain CPU oracle (rel-L2 <= 1e-12) for every halo path and schedule.  torch.distributed (gloo) is the
control plane only."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path[:0] = [os.path.join(ROOT, "mpas-ocean.jl_b200"), os.path.join(ROOT, "oracle")]
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import moka_b200 as mb  # noqa: E402
import moka_oracle_c as OC  # noqa: E402
from moka_b200 import multi_gpu, partition  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    nx, nsteps = int(os.environ.get("MOKAB_CHECK_NX", "96")), 21
    m = mb.periodic_hex(nx, nx, 1.0e7 / nx, with_dual=False)
    state = mb.inertialGravityWave(m).initial_state()
    dt = mb.cfl_dt(m["dc"])
    loc = partition.decompose(m, world)[rank]
    backend = mb.B200(local)
    rt = multi_gpu.TorchRuntime(local, device="cpu")
    comm = multi_gpu.Communicator(backend, rt)          # one NCCL communicator for all the models below
    errs = []
    # (overlap, graph, halo, stepper): the packed NCCL exchange in its three schedules, the direct-store exchange as push / wait
    # kernels and folded into the boundary launch, ForwardEuler (the reference driver's stepper) host-launched and as graphs
    cases = [(True, False, "nccl", "RungeKutta4"), (False, False, "nccl", "RungeKutta4"), (True, True, "nccl", "RungeKutta4"),
             (True, False, "p2p", "RungeKutta4"), (True, True, "p2p", "RungeKutta4"), (True, False, "p2p_fused", "RungeKutta4"),
             (True, True, "p2p_fused", "RungeKutta4"), (True, False, "nccl", "ForwardEuler"), (False, False, "nccl", "ForwardEuler"),
             (True, True, "nccl", "ForwardEuler"),
             (True, False, "p2p_ll", "RungeKutta4"), (True, True, "p2p_ll", "RungeKutta4"), (False, False, "p2p_ll", "RungeKutta4"),
             (True, True, "p2p_ll", "ForwardEuler")]   # flag-in-data packets
    if os.environ.get("MOKAB_CHECK_HALO"):
        cases = [c for c in cases if c[2] in os.environ["MOKAB_CHECK_HALO"].split(",")]
    for overlap, graph, halo, stepper in cases:
        model = multi_gpu.DecomposedModel(loc, multi_gpu.local_state(loc, *state), backend, local, overlap=overlap, graph=graph, halo=halo,
                                          runtime=rt, comm=comm)
        for n in (8, 1, 10, 2):                      # 21 steps; the odd call moves the time-level parity between graph replays
            model.step(dt, n, stepper=getattr(mb, stepper))
        model.finish()
        gs, gu, gh = multi_gpu.gather_owned(model, m["nCells"], m["nEdges"])
        mass = model.reduce("mass")
        status = model.graph_status
        model.close()
        del model
        if rank == 0:
            OC.sign_index_fields(m)
            om = OC.OracleModel(m, *state)
            om.run_loop(dt, nsteps, stepper)
            rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))
            e = (rel(gs, om.ssh[1]), rel(gu, om.normalVelocity[1]), rel(gh, om.layerThickness[1]))
            m0 = float(np.sum(m["areaCell"] * om.layerThickness[1]))
            errs.append(((overlap, graph, halo, stepper), e, abs(mass - m0) / m0, status))
    # reverse mode on the decomposed mesh: J = sum ssh^2 after 6 steps and dJ/d(initial state) against the adjoint oracle
    for halo, graph, stepper in (("nccl", True, "RungeKutta4"), ("p2p", False, "RungeKutta4"), ("nccl", True, "ForwardEuler"), ("p2p_ll", True, "RungeKutta4")):
        if os.environ.get("MOKAB_CHECK_HALO") and halo not in os.environ["MOKAB_CHECK_HALO"].split(","):
            continue
        model = multi_gpu.DecomposedModel(loc, multi_gpu.local_state(loc, *state), backend, local, graph=graph, halo=halo, runtime=rt, comm=comm)
        J = model.reverse_run_loop(dt, 6, stepper=getattr(mb, stepper))
        model.finish()
        gu_o, gh_o = model.gradient()
        no, ne = loc["nCellsOwned"], loc["nEdgesOwned"]
        gu, gh = np.zeros(m["nEdges"]), np.zeros(m["nCells"])
        gu[loc["edgesGlobal"][:ne]], gh[loc["cellsGlobal"][:no]] = gu_o, gh_o
        gu, gh = comm.allreduce(gu), comm.allreduce(gh)        # (owned parts are disjoint: the sum assembles the global arrays)
        model.close()
        del model
        if rank == 0:
            import adjoint_oracle as AO
            OC.sign_index_fields(m)
            if stepper == "ForwardEuler":
                Jo, ou, oh = AO.gradient_sum_ssh2_fe(m, state[0], state[1], state[2], dt, 6)[:3]
            else:
                Jo, ou, oh = AO.gradient_sum_ssh2(m, state[1], state[2], dt, 6)
            rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))
            errs.append((("reverse mode", graph, halo, stepper), (abs(J - Jo) / Jo, rel(gu, ou), rel(gh, oh)), 0.0, "n/a"))
    # multi-level states: forward (bit for bit the numpy oracle with a level axis) and the reverse mode (level-axis adjoint oracle)
    # (opt-in until its first hardware run: MOKAB_CHECK_LEVELS=1; with emulated ranks: tests/sim/check_decomposed.py run_levels / run_adjoint_levels)
    if os.environ.get("MOKAB_CHECK_LEVELS") and not os.environ.get("MOKAB_CHECK_HALO"):
        K = 3
        mk = dict(m)
        frac = np.random.default_rng(K).uniform(0.5, 1.5, K)
        frac /= frac.sum()
        rest = np.outer(state[2] - state[0], frac)
        hk, uk = rest + np.outer(state[0], frac), np.outer(state[1], 1.0 + 0.1 * np.arange(K))
        mk["restingThickness"], mk["nVertLevels"] = rest, K
        lock = partition.decompose(mk, world)[rank]
        model = multi_gpu.DecomposedModel(lock, multi_gpu.local_state(lock, state[0], uk, hk), backend, local, graph=True, runtime=rt, comm=comm)
        J = model.reverse_run_loop(dt, 4)
        model.finish()
        no, ne = lock["nCellsOwned"], lock["nEdgesOwned"]
        fu, gu_o, gh_o = np.array(model.owned("normalVelocity")), *model.gradient()
        allu, gu, gh = (np.zeros((m["nEdges"], K)), np.zeros((m["nEdges"], K)), np.zeros((m["nCells"], K)))
        allu[lock["edgesGlobal"][:ne]], gu[lock["edgesGlobal"][:ne]], gh[lock["cellsGlobal"][:no]] = fu, gu_o, gh_o
        allu, gu, gh = (comm.allreduce(a.ravel()).reshape(a.shape) for a in (allu, gu, gh))
        model.close()
        del model
        if rank == 0:
            import adjoint_oracle as AO
            OC.sign_index_fields(mk)
            Jo, ou, oh = AO.gradient_sum_ssh2_levels(mk, np.ascontiguousarray(uk.T), np.ascontiguousarray(hk.T), dt, 4)
            fwd = AO.run_forward_levels(mk, np.ascontiguousarray(uk.T), np.ascontiguousarray(hk.T), dt, 4)[-1][0]
            rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))
            errs.append((("multi-level reverse mode", True, "nccl", "RungeKutta4"), (abs(J - Jo) / Jo, rel(gu.T, ou), rel(gh.T, oh)), 0.0 if np.array_equal(allu.T, fwd) else 1.0, "n/a"))
    if rank == 0:
        print(errs)
        ok = all(max(e) <= 1e-12 and dm <= 1e-13 for _, e, dm, _ in errs)
        print("MULTI_GPU_CHECK_OK" if ok else "MULTI_GPU_CHECK_FAILED")
    sys.stdout.flush()
    comm.destroy()
    dist.barrier()
    torch.cuda.synchronize()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
