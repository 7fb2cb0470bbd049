"""One process per GPU: the host side of domain-decomposed runs.

The reference has no multi-device path; this follows BASELINE.json's north_star: owned / halo cell and
edge layers (partition.py), one message per neighbour and RK stage over NVLink -- NCCL send/recv or direct stores into the
neighbours' memory -- overlapped with the interior blocks of the same stage.  The exchange, the two streams, the events and
the captured step graphs live INSIDE libmoka_b200.so (csrc/comm.cuh, csrc/decomposed.cuh: mokab_comm_init,
mokab_decomp_setup, mokab_timestep_*_decomposed, mokab_reduce_decomposed); this module builds the local meshes, brings the
ranks together (a 128-byte id from rank 0 to the others) and forwards calls.  torch.distributed appears only as that control
plane (and as the transport of the CPU-only gloo tests, `HaloExchanger`).
"""
from __future__ import annotations

import ctypes as C
import json
import os
import time

import numpy as np

from . import _lib as L
from . import api, partition


class HaloExchanger:
    """Packed all-to-all of halo messages over torch.distributed -- the transport of the CPU-only gloo tests
    (tests/test_partition_cpu.py: numpy oracle as the per-rank compute).  The GPU path does not use it: its exchange is
    comm::all_to_all inside the library."""

    def __init__(self, send_counts, recv_counts, dtype, device, group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group = torch, dist, group
        self.send_counts, self.recv_counts = list(send_counts), list(recv_counts)
        self.send = torch.zeros(max(1, sum(send_counts)), dtype=dtype, device=device)
        self.recv = torch.zeros(max(1, sum(recv_counts)), dtype=dtype, device=device)
        self.rank = dist.get_rank(group)
        self.use_a2a = device != "cpu" and str(device) != "cpu"

    def exchange(self) -> None:
        ns, nr = sum(self.send_counts), sum(self.recv_counts)
        if self.use_a2a:
            self.dist.all_to_all_single(self.recv[:nr], self.send[:ns], self.recv_counts, self.send_counts, group=self.group)
            return
        ops, so, ro = [], 0, 0
        for q, (cs, cr) in enumerate(zip(self.send_counts, self.recv_counts)):
            if cr:
                ops.append(self.dist.P2POp(self.dist.irecv, self.recv[ro:ro + cr], q, group=self.group))
            if cs:
                ops.append(self.dist.P2POp(self.dist.isend, self.send[so:so + cs], q, group=self.group))
            so, ro = so + cs, ro + cr
        if ops:
            for w in self.dist.batch_isend_irecv(ops):
                w.wait()


class TorchRuntime:
    """The CONTROL plane of a torchrun job: who am I, and a way to hand rank 0's communicator id to the others (any
    torch.distributed backend; gloo is enough).  The data plane -- halo messages, reductions, set-up exchanges -- is
    NCCL inside libmoka_b200.so (csrc/comm.cuh).  tests/sim passes an object of the same shape for emulated ranks."""

    def __init__(self, device_index: int = 0, group=None, device=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group = torch, dist, group
        self.dev = torch.device("cuda", device_index) if device is None else torch.device(device)   # "cpu": gloo tests of the host exchanges

    def rank_and_size(self):
        return self.dist.get_rank(self.group), self.dist.get_world_size(self.group)

    def broadcast_bytes(self, blob, src: int = 0) -> bytes:
        box = [blob]
        self.dist.broadcast_object_list(box, src=src, group=self.group)
        return box[0]

    # host-side exchanges of the gloo tests (tests/test_partition_cpu.py); the GPU path does these inside the library
    def exchanger(self, send_counts, recv_counts, npdtype):
        tdt = self.torch.float64 if np.dtype(npdtype) == np.float64 else self.torch.float32
        return HaloExchanger(send_counts, recv_counts, tdt, self.dev, self.group)


class StoreRuntime:
    """The same control plane without a process group: a torch.distributed TCPStore carries the 128-byte id (torchrun's
    MASTER_ADDR / MASTER_PORT + 1 by default).  What the Julia shim does with MPI.jl or a shared file."""

    def __init__(self, rank: int, world: int, host: str | None = None, port: int | None = None):
        import torch.distributed as dist
        self.rank, self.world = rank, world
        host = host or os.environ.get("MASTER_ADDR", "127.0.0.1")
        port = port or int(os.environ.get("MASTER_PORT", "29500")) + 1
        self.store = dist.TCPStore(host, port, world, is_master=(rank == 0))
        self._n = 0

    def rank_and_size(self):
        return self.rank, self.world

    def broadcast_bytes(self, blob, src: int = 0) -> bytes:
        key = f"mokab_bcast_{self._n}"
        self._n += 1
        if self.rank == src:
            self.store.set(key, bytes(blob))
        return bytes(self.store.get(key))


class Communicator:
    """mokab_comm: the NCCL communicator of the ranks, inside the library.  Rank 0 draws the unique id, `runtime` hands it to
    the others, every rank joins (collective)."""

    def __init__(self, backend: api.B200, runtime):
        lib = L.lib()
        self.backend, self.rt = backend, runtime
        self.rank, self.nranks = runtime.rank_and_size()
        blob = None
        if self.rank == 0:
            buf = C.create_string_buffer(L.COMM_ID_BYTES)
            L.check(lib.mokab_comm_get_unique_id(buf))
            blob = buf.raw
        blob = runtime.broadcast_bytes(blob, 0)
        h = C.c_void_p()
        L.check(lib.mokab_comm_init(backend.handle, blob, self.rank, self.nranks, C.byref(h)))
        self.handle = h

    def barrier(self) -> None:
        L.check(L.lib().mokab_comm_barrier(self.handle))

    def allreduce(self, values, op: str = "sum") -> np.ndarray:
        a = np.ascontiguousarray(np.atleast_1d(values), np.float64).copy()
        L.check(L.lib().mokab_comm_allreduce_f64(self.handle, a.ctypes.data_as(L._F64P), a.size, {"sum": 0, "max": 1, "min": 2}[op]))
        return a

    def allgather(self, mine: np.ndarray) -> np.ndarray:
        """Every rank's array (same shape and dtype on every rank), stacked rank-major."""
        a = np.ascontiguousarray(mine)
        out = np.empty((self.nranks,) + a.shape, a.dtype)
        L.check(L.lib().mokab_comm_allgather_bytes(self.handle, a.ctypes.data_as(C.c_void_p), a.nbytes, out.ctypes.data_as(C.c_void_p)))
        return out

    def destroy(self) -> None:
        if self.handle is not None:
            L.check(L.lib().mokab_comm_destroy(self.handle))
            self.handle = None


HALO_MODES = {"nccl": L.HALO_NCCL, "p2p": L.HALO_P2P, "p2p_fused": L.HALO_P2P_FUSED, "p2p_ll": L.HALO_P2P_LL}


class DecomposedModel:
    """This rank's share of the mesh on its GPU, stepped by the library's own decomposed entry points
    (mokab_decomp_setup / mokab_timestep_*_decomposed, csrc/decomposed.cuh): the halo exchange per RK stage -- NCCL
    send/recv, or direct stores into the neighbours' memory -- runs on a high-priority stream next to the interior blocks,
    and one / two consecutive steps are replayed as captured CUDA graphs.  This class only builds the local mesh and state,
    brings the ranks together and forwards calls; no stream, event or collective is issued from Python.

    `graph=True` graphs are VALIDATED on first use: replaying them must reproduce, bit for bit on every rank, what the
    host-launched schedule computes from the same state, else the model keeps launching from the host (`graph_status`)."""

    def __init__(self, loc: dict, state, backend: api.B200, device_index: int = 0, dtype=np.float64, group=None, overlap=True,
                 graph=False, runtime=None, halo="nccl", comm: Communicator | None = None):
        if halo not in HALO_MODES:
            raise api.MokaError("DecomposedModel: halo must be 'nccl' (packed NCCL send/recv), 'p2p' (direct peer stores, push and wait "
                                "kernels), 'p2p_fused' (direct peer stores from inside the boundary launch) or 'p2p_ll' (flag-in-data "
                                "packets into the peers' receive areas)")
        self.halo_mode = halo
        self.rt = runtime if runtime is not None else TorchRuntime(device_index, group)
        self.loc, self.backend, self.overlap, self.use_graph = loc, backend, overlap, graph
        self.nparts = loc["nparts"]
        self._own_comm = comm is None
        self.comm = comm if comm is not None else Communicator(backend, self.rt)
        if self.comm.nranks != self.nparts:
            raise api.MokaError("DecomposedModel: the mesh was decomposed for a different number of ranks")
        self.mesh = api.Mesh(loc, backend)
        sidx, scnt, ridx, rcnt = partition.flat_halo(loc, self.nparts)
        self.mesh.halo_setup(sidx, ridx)
        self.send_counts, self.recv_counts = [int(c) for c in scnt], [int(c) for c in rcnt]
        ssh, u, h = state
        self.prog = api.PrognosticVars(np.asarray(ssh, dtype), np.asarray(u, dtype), np.asarray(h, dtype), 2, self.mesh)
        self.handle = self.prog.dev.handle
        sc, rc = np.asarray(scnt, np.int64), np.asarray(rcnt, np.int64)
        i64p = C.POINTER(C.c_int64)
        L.check(L.lib().mokab_decomp_setup(self.handle, self.comm.handle, sc.ctypes.data_as(i64p), rc.ctypes.data_as(i64p),
                                           HALO_MODES[halo], self._flags(graph)))
        self._validated, self.graph_status = False, "not used"
        self._graph_dt = None
        self._fe, self._stepped = False, False                   # the model steps with ForwardEuler; it has stepped
        self._closed = False

    def _flags(self, graph: bool) -> int:
        return (0 if self.overlap else L.DECOMP_NO_OVERLAP) | (0 if graph else L.DECOMP_NO_GRAPH)

    def _set_graph(self, on: bool) -> None:
        L.check(L.lib().mokab_decomp_set_flags(self.handle, self._flags(on)))

    def _advance(self, dt: float, nsteps: int, fe: bool) -> None:
        fn = L.lib().mokab_timestep_forward_euler_decomposed if fe else L.lib().mokab_timestep_rk4_decomposed
        L.check(fn(self.handle, float(dt), int(nsteps)))

    def _snapshot(self):
        return [self.prog.dev.get(f) for f in (L.SSH, L.NORMAL_VELOCITY, L.LAYER_THICKNESS)]

    def _restore(self, snap) -> None:
        for f, a in zip((L.LAYER_THICKNESS, L.NORMAL_VELOCITY, L.SSH), (snap[2], snap[1], snap[0])):
            self.prog.dev.set(f, a)

    def validate_graph(self, dt: float, nsteps: int = 7) -> bool:
        """Accept the captured graphs only if replaying them (one 2-step graph per time-level parity and the 1-step graph:
        `nsteps` odd) reproduces what the host-launched schedule computes from the same state; the state is restored."""
        snap = self._snapshot()
        self._set_graph(False)
        self._advance(dt, nsteps, False)
        self.finish()
        want = self._snapshot()
        self._restore(snap)
        self._set_graph(True)
        self._advance(dt, nsteps, False)
        self.finish()
        got = self._snapshot()
        self._restore(snap)
        ok = all(np.array_equal(a, b) for a, b in zip(want, got))
        flag = int(self.comm.allreduce(1.0 if ok else 0.0, "min")[0])
        self._validated, self._graph_dt = True, dt
        if flag == 0:
            self.use_graph = False
            self._set_graph(False)
            self.graph_status = "rejected: graph replay did not reproduce the host-launched schedule; launching from the host"
            return False
        self.graph_status = f"validated against the host-launched schedule over {nsteps} steps"
        return True

    def step(self, dt: float, nsteps: int = 1, stepper=None) -> None:
        """`nsteps` steps of `stepper` (api.RungeKutta4, the default, or api.ForwardEuler); asynchronous."""
        if stepper not in (None, api.RungeKutta4, api.ForwardEuler):
            raise api.MokaError("DecomposedModel.step: unknown stepper")
        fe = stepper is api.ForwardEuler
        if self._stepped and fe != self._fe:                     # (ForwardEuler carries a lagged layerThicknessEdge between its steps)
            raise api.MokaError("DecomposedModel.step: one stepper per model -- build another DecomposedModel to change it")
        self._stepped = self._stepped or nsteps > 0
        self._fe = fe
        if self.use_graph and not fe and nsteps > 0 and (not self._validated or self._graph_dt != dt):
            self.validate_graph(dt)
        self._advance(dt, nsteps, fe)

    def reverse_run_loop(self, dt: float, nsteps: int, seed: str | None = "ssh2", stepper=None) -> float:
        """`autodiff(Reverse, ocn_run_loop, ...)` (test_Enzyme_end2end.jl:78-96) on the decomposed mesh: `nsteps` steps of
        `stepper` (api.RungeKutta4, the default, or api.ForwardEuler, the stepper the reference differentiates) recording
        every rank's part of the trajectory, then the reverse sweep -- the hand-written gather-form adjoints with plain halo
        copies of everything a reversed stage / step consumes (recomputed stage states, adjoint variables), over the lists
        of the forward exchange.  Returns J = sum ssh^2 over the whole mesh (`seed=None`: the caller has set the adjoint of
        the final state with `prog.dev.set(L.D_*, ...)`); `gradient()` then holds dJ/d(initial state)."""
        if stepper not in (None, api.RungeKutta4, api.ForwardEuler):
            raise api.MokaError("DecomposedModel.reverse_run_loop: unknown stepper")
        fe = stepper is api.ForwardEuler
        if self._stepped and fe != self._fe:
            raise api.MokaError("DecomposedModel.reverse_run_loop: one stepper per model -- build another DecomposedModel to change it")
        if self.use_graph and not fe and nsteps > 0 and (not self._validated or self._graph_dt != dt):
            self.validate_graph(dt)                               # (before the tape starts: validation steps the model)
        lib = L.lib()
        L.check(lib.mokab_tape_begin(self.handle, int(nsteps)))
        self._stepped = self._stepped or nsteps > 0
        self._fe = fe
        self._advance(dt, nsteps, fe)
        J = float("nan")
        if seed is not None:
            if seed != "ssh2":
                raise api.MokaError("DecomposedModel.reverse_run_loop: unknown seed")
            J = self.reduce("ssh2")
            L.check(lib.mokab_adjoint_seed(self.handle, L.SUM_SSH2))
        L.check((lib.mokab_adjoint_forward_euler if fe else lib.mokab_adjoint_rk4)(self.handle))
        return J

    def gradient(self):
        """(d_normalVelocity on the owned edges, d_layerThickness on the owned cells) after `reverse_run_loop`."""
        gu = np.asarray(self.prog.dev.get(L.D_NORMAL_VELOCITY), np.float64)
        gh = np.asarray(self.prog.dev.get(L.D_LAYER_THICKNESS), np.float64)
        if gu.size != self.loc["nEdges"]:                         # multi-level states: (local entities, nVertLevels)
            gu, gh = gu.reshape(self.loc["nEdges"], -1), gh.reshape(self.loc["nCells"], -1)
        return gu[:self.loc["nEdgesOwned"]], gh[:self.loc["nCellsOwned"]]

    def gradient_ssh(self):
        """d_ssh on the owned cells: ForwardEuler reads the initial ssh array as an input of its own (first step's pressure gradient)."""
        return np.asarray(self.prog.dev.get(L.D_SSH), np.float64)[:self.loc["nCellsOwned"]]

    def refresh_ssh(self) -> None:
        """ssh = layerThickness - restingThicknessSum (the decomposed RK4 entry point leaves it refreshed already)."""

    def synchronize(self) -> None:
        L.check(L.lib().mokab_decomp_synchronize(self.handle))

    def finish(self) -> None:
        self.synchronize()

    def close(self) -> None:
        """Collective: unmap the peers (direct-store paths), drop the graphs and the halo stream, leave the communicator."""
        if self._closed:
            return
        self._closed = True
        L.check(L.lib().mokab_decomp_close(self.handle))
        if self._own_comm:
            self.comm.destroy()

    def owned(self, field: str) -> np.ndarray:
        """The rank's owned part of a field; multi-level fields come back as (owned entities, nVertLevels)."""
        a = np.asarray(getattr(self.prog, field))
        cells = field in ("ssh", "layerThickness")
        n, nloc = (self.loc["nCellsOwned"], self.loc["nCells"]) if cells else (self.loc["nEdgesOwned"], self.loc["nEdges"])
        if a.size != nloc:
            a = a.reshape(nloc, -1)
        return a[:n]

    def reduce(self, which: str) -> float:
        out = C.c_double()
        L.check(L.lib().mokab_reduce_decomposed(self.handle, {"ssh2": L.SUM_SSH2, "mass": L.SUM_MASS, "energy": L.SUM_ENERGY}[which], C.byref(out)))
        return out.value

    def gather(self, nC: int, nE: int, previous: bool = False):
        """The global (ssh, normalVelocity, layerThickness) on every rank, assembled from the owned parts (`previous`: the
        time level one step back, which is what the reference's write_netcdf writes, PrognosticVars.jl:108-113)."""
        loc, out = self.loc, []
        names = (("ssh_prev", "normalVelocity_prev", "layerThickness_prev") if previous else ("ssh", "normalVelocity", "layerThickness"))
        for name, n, ids, no in ((names[0], nC, loc["cellsGlobal"], loc["nCellsOwned"]), (names[1], nE, loc["edgesGlobal"], loc["nEdgesOwned"]),
                                 (names[2], nC, loc["cellsGlobal"], loc["nCellsOwned"])):
            nmax = int(self.comm.allreduce(float(no), "max")[0])
            vals, gid = np.zeros(nmax, np.float64), np.full(nmax, -1, np.int64)
            vals[:no], gid[:no] = np.asarray(getattr(self.prog, name), np.float64)[:no], ids[:no]
            allv, alli = self.comm.allgather(vals), self.comm.allgather(gid)
            g = np.zeros(n, np.float64)
            keep = alli >= 0
            g[alli[keep]] = allv[keep]
            out.append(g)
        return out


def local_state(loc: dict, ssh, u, h):
    return ssh[loc["cellsGlobal"]], u[loc["edgesGlobal"]], h[loc["cellsGlobal"]]


def gather_owned(model: DecomposedModel, nC: int, nE: int):
    """Assemble the global (ssh, u, h) on every rank from the owned parts (test/diagnostic helper)."""
    return model.gather(nC, nE)


# ---- bench leg for torchrun (N > 1) -------------------------------------------------------------------------
def _share_locals(args, rank, world, nx, dtype, keep_global=False):
    """Every rank generates the (deterministic) global mesh, derives the same partition from it and keeps its own part --
    no rank waits for another, nothing goes through files.  With `keep_global` rank 0 also returns (mesh, state) for the
    parity check of the bench line."""
    import torch.distributed as dist
    t0 = time.time()
    from . import planar_hex
    if args.workload.startswith("kelvin"):
        m = planar_hex.channel_hex(nx, nx, 1.0e7 / nx)
        ssh, u, h = api.kelvinWave(m).initial_state()
    elif args.workload.startswith("voronoi"):
        from . import planar_voronoi
        m = planar_voronoi.periodic_voronoi(nx, nx, 1.0e7 / nx, jitter=0.25, seed=2, allow_obtuse=True, with_dual=False)
        ssh, u, h = api.inertialGravityWave(m).initial_state()
    elif args.workload.startswith("sphere"):
        from . import spherical_voronoi
        m = spherical_voronoi.spherical_voronoi(nx * nx, with_dual=False)
        ssh, u, h = spherical_voronoi.geostrophic_zonal_flow(m)
        u = u + 0.1 * np.random.default_rng(0).standard_normal(m["nEdges"])
        m["bench_dt"] = 0.25 * float(m["dcEdge"].min()) / float(np.sqrt(api.GRAVITY * 1000.0))
    else:
        m = planar_hex.periodic_hex(nx, nx, 1.0e7 / nx, with_dual=False)
        ssh, u, h = api.inertialGravityWave(m).initial_state()
    z = m.get("zCell")
    part = partition.rcb_partition(m["xCell"], m["yCell"], world, z if z is not None and np.ptp(z) > 0 else None)
    loc = partition.build_local_mesh(m, part, rank)
    sets = [None] * world                                        # what every rank holds as halo copies (sorted global ids: small)
    dist.all_gather_object(sets, (loc["cellsGlobal"][loc["nCellsOwned"]:], loc["edgesGlobal"][loc["nEdgesOwned"]:]))
    loc = partition.decompose_one(m, world, rank, part=part, loc=loc, halo_sets=sets)
    if "bench_dt" in m:
        loc["bench_dt"] = m["bench_dt"]
    state = local_state(loc, ssh, u, h)
    glob = (m, (ssh, u, h)) if (keep_global and rank == 0) else None
    del m
    dist.barrier()
    if keep_global:
        return loc, state, time.time() - t0, glob
    return loc, state, time.time() - t0


def _collect_on_rank0(model, rank, world, nC, nE, tag):
    """The owned (ssh, normalVelocity) of every rank assembled into global arrays on rank 0 (through /dev/shm: a one-off
    check, not a data path)."""
    import torch.distributed as dist
    loc = model.loc
    no, ne = loc["nCellsOwned"], loc["nEdgesOwned"]
    np.savez(f"{tag}_par_{rank}.npz", ssh=np.asarray(model.prog.ssh, np.float64)[:no], u=np.asarray(model.prog.normalVelocity, np.float64)[:ne],
             cells=loc["cellsGlobal"][:no], edges=loc["edgesGlobal"][:ne])
    dist.barrier()
    out = None
    if rank == 0:
        gs, gu = np.full(nC, np.nan), np.full(nE, np.nan)
        for r in range(world):
            z = np.load(f"{tag}_par_{r}.npz")
            gs[z["cells"]], gu[z["edges"]] = z["ssh"], z["u"]
            os.remove(f"{tag}_par_{r}.npz")
        out = (gs, gu)
    dist.barrier()
    return out


def bind_to_gpu_numa_node(device_index: int):
    """Pin this process (and with it the page-locked buffers it allocates from now on) to the CPUs of the NUMA node the GPU
    hangs off: with one process per GPU on a two-socket box, host<->device copies otherwise cross the socket interconnect and
    all ranks' pinned memory can end up on one node.  Best effort -- returns the node id, or None when sysfs says nothing."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(device_index).pci_bus_id
        dom = torch.cuda.get_device_properties(device_index).pci_domain_id
        dev = torch.cuda.get_device_properties(device_index).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def bench_main(args, rank, world, local):
    import torch
    import torch.distributed as dist
    from bench import (WORKLOADS, ClockSampler, algo_bytes_per_cell_step, algo_bytes_per_cell_step_general, measured_peak_gbs,
                       parity_against_oracle, workload_label)
    torch.cuda.set_device(local)
    numa = None if os.environ.get("MOKAB_NO_NUMA_BIND") else bind_to_gpu_numa_node(local)
    # torch.distributed is the CONTROL plane only (gloo: sharing the decomposition through /dev/shm, handing out the
    # communicator id); every byte of halo data and every reduction moves through NCCL inside libmoka_b200.so
    dist.init_process_group("gloo")
    nx = WORKLOADS[args.workload]
    npdt = np.float64 if args.dtype == "f64" else np.float32
    loc, state, t_setup, glob = _share_locals(args, rank, world, nx, npdt, keep_global=True)
    nC_glob = loc["nCellsGlobal"] if "nCellsGlobal" in loc else nx * nx
    sphere = args.workload.startswith("sphere")
    voronoi = args.workload.startswith("voronoi") or sphere      # unstructured: byte accounting from the actual rows
    dt = float(loc["bench_dt"]) if sphere else (0.25 if voronoi else 1.0) * api.cfl_dt(1.0e7 / nx)
    backend = api.B200(local)
    rt = TorchRuntime(local, device="cpu")
    comm = Communicator(backend, rt)
    model = DecomposedModel(loc, state, backend, local, dtype=npdt, overlap=not getattr(args, "no_overlap", False),
                            graph=not getattr(args, "no_graph", False), runtime=rt, halo=getattr(args, "halo", "nccl"), comm=comm)
    K, W = args.steps, max(args.warmup, 3)
    model.step(dt, W)                                             # (validates the captured graphs on first use)
    model.finish()
    # how many batches of K steps make the timed region >= 0.5 s (one scheduler hiccup must not be 5 % of the number)
    backend.timer_start()
    model.step(dt, K)
    probe_ms = float(comm.allreduce(backend.timer_stop(), "max")[0])
    reps = int(max(1, np.ceil(500.0 / max(probe_ms, 1e-3))))
    comm.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    comm.barrier()
    torch.cuda.synchronize()
    l0 = backend.launch_count()
    backend.timer_start()                                        # CUDA events on the compute stream; every step call ends with the halo stream joined into it
    for _ in range(reps):
        model.step(dt, K)
    ms_local = backend.timer_stop()
    model.synchronize()
    launches = backend.launch_count() - l0
    ms = float(comm.allreduce(ms_local, "max")[0])
    comm.barrier()
    torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    model.finish()
    steps_timed = K * reps

    # end to end with HOST buffers: every step uploads this rank's (normalVelocity, layerThickness | ssh) from pinned memory, takes
    # one step (a captured 1-step graph, exchange included) and reads back the new (ssh, normalVelocity) -- layerThickness = ssh +
    # restingThicknessSum is redundant with ssh and stays on the device; pipelined through the API's copy streams
    nCl, nEl = loc["nCells"], loc["nEdges"]
    hin = [(backend.pinned(nEl, npdt), backend.pinned(nCl, npdt)) for _ in range(2)]
    hout = [(backend.pinned(nCl, npdt), backend.pinned(nEl, npdt)) for _ in range(2)]
    f32 = npdt == np.float32
    for hu, hh in hin:
        hu[:], hh[:] = np.asarray(state[1], npdt), np.asarray(state[0] if f32 else state[2], npdt)
    Ke = max(3, min(K, 20))

    def e2e_steps(n):
        for i in range(n):
            if f32:
                model.prog.upload_async(normalVelocity=hin[i & 1][0], ssh=hin[i & 1][1])
            else:
                model.prog.upload_async(normalVelocity=hin[i & 1][0], layerThickness=hin[i & 1][1])
            model.step(dt, 1)
            model.prog.download_async(ssh=hout[i & 1][0], normalVelocity=hout[i & 1][1])
        model.prog.synchronize()
        model.synchronize()
    e2e_steps(2)
    comm.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e2e_steps(Ke)
    torch.cuda.synchronize()
    e2e_s = float(comm.allreduce((time.perf_counter() - t0) / Ke, "max")[0])
    # the two halves of that leg on their own (explain it: which of them the leg's time follows)
    def copy_only(n):
        for i in range(n):
            if f32:
                model.prog.upload_async(normalVelocity=hin[i & 1][0], ssh=hin[i & 1][1])
            else:
                model.prog.upload_async(normalVelocity=hin[i & 1][0], layerThickness=hin[i & 1][1])
            model.prog.download_async(ssh=hout[i & 1][0], normalVelocity=hout[i & 1][1])
        model.prog.synchronize()
        model.synchronize()
    copy_only(1)
    comm.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    copy_only(Ke)
    torch.cuda.synchronize()
    copy_s = float(comm.allreduce((time.perf_counter() - t0) / Ke, "max")[0])
    comm.barrier()
    t0 = time.perf_counter()
    for _ in range(Ke):
        model.step(dt, 1)
    model.synchronize()
    torch.cuda.synchronize()
    step1_s = float(comm.allreduce((time.perf_counter() - t0) / Ke, "max")[0])
    cnt = comm.allreduce([nCl + nEl, nCl + nEl, loc["nCellsOwned"], launches], "sum")
    blocks = model.mesh.block_counts()
    halo_bytes = (sum(model.send_counts) + sum(model.recv_counts)) * np.dtype(npdt).itemsize

    # parity: two steps from the initial state, gathered on rank 0, against the C oracle on the undecomposed mesh
    parity = None
    if not getattr(args, "no_parity", False):
        model._restore((state[0], state[1], state[2]))
        model.step(dt, 2)
        model.finish()
        got = _collect_on_rank0(model, rank, world, glob[0]["nCells"] if rank == 0 else 0, glob[0]["nEdges"] if rank == 0 else 0,
                                f"/dev/shm/mokab_{os.environ.get('MASTER_PORT', '0')}_{nx}_{world}")
        if rank == 0:
            parity = parity_against_oracle(glob[0], glob[1], dt, 2, got[0], got[1], args.dtype, args.workload)
    if rank == 0:
        item = np.dtype(npdt).itemsize
        peak, peak_src = measured_peak_gbs()
        value = nC_glob * steps_timed / (ms * 1e-3)
        nblk, nder = model.mesh.derived_blocks()
        if voronoi:
            nco, neo = loc["nCellsOwned"], loc["nEdgesOwned"]
            owned = {"nCells": nco, "nEdges": neo, "nEdgesOnCell": loc["nEdgesOnCell"][:nco], "nEdgesOnEdge": loc["nEdgesOnEdge"][:neo]}
            per_cell_step = algo_bytes_per_cell_step_general(owned, args.dtype, nder / max(nblk, 1))
        else:
            per_cell_step = algo_bytes_per_cell_step(args.dtype, nder / max(nblk, 1))
        algo_per_launch = per_cell_step / 4.0 * loc["nCellsOwned"]
        achieved = algo_per_launch / ((ms * 1e-3) / (4 * steps_timed)) / 1e9
        halo_txt = {"nccl": "NCCL send/recv inside libmoka_b200.so", "p2p": "direct peer stores + arrival counters (push / wait kernels)",
                    "p2p_fused": "direct peer stores from inside the boundary launch",
                    "p2p_ll": "flag-in-data packets into the peers' receive areas (no fence, no counter)"}[model.halo_mode]
        print(json.dumps({
            "metric": "RK4 cell-steps/sec", "value": value, "unit": "cell-steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / steps_timed, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": args.dtype,
            "data": "synthetic",
            "config": {"workload": workload_label(args.workload, nx), "name": args.workload,
                       "detail": f"{nC_glob} cells, {'Float64' if args.dtype == 'f64' else 'Float32'} RK4, dt={dt:.4g}s, recursive-coordinate-bisection "
                                 f"into {world} parts, 1 halo layer, {halo_txt} per stage "
                                 f"{'overlapped with interior blocks' if model.overlap else '(no overlap)'}"
                                 f"{', 1- / 2-step CUDA graphs incl. the exchange (' + model.graph_status + ')' if model.use_graph else ', CUDA graph ' + model.graph_status}",
                       "timed_steps": steps_timed, "repeats_of_steps": reps,
                       "l2": "inputs larger than L2 (no flush)", "setup_s": round(t_setup, 1), "rank0_numa_node": numa,
                       "rank0_blocks_interior_boundary": list(blocks), "rank0_halo_bytes_per_stage": int(halo_bytes),
                       "rank0_blocks_rebuilding_edgesOnEdge": [int(nder), int(nblk)],
                       "variant": {"lib": os.environ.get("MOKAB_LIB", "libmoka_b200.so"),
                                   "stage_tma": L.get_option("stage_tma"), "stage_prefetch": L.get_option("stage_prefetch"),
                                   "stage_prefetch_distance": L.get_option("stage_prefetch_distance"),
                                   "stage_auto": L.get_option("stage_auto"), "stage_pdl": L.get_option("stage_pdl")}},
            "clocks": clocks,
            "e2e": {"value": nC_glob / e2e_s, "unit": "cell-steps/s", "h2d_bytes_per_step": int(cnt[0] * item),
                    "d2h_bytes_per_step": int(cnt[1] * item), "ms_per_step": e2e_s * 1e3, "steps": Ke,
                    "pipelined": True, "copies_only_ms_per_step": copy_s * 1e3, "one_step_calls_only_ms_per_step": step1_s * 1e3,
                    "returns": "ssh + normalVelocity of the new state (layerThickness = ssh + restingThicknessSum stays on the device)"},
            "gpu_launches": int(cnt[3]),
            "parity": parity,
            "roofline": {"bound": "hbm", "kernel": "k_rk_stage", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                         "note": "per GPU: algorithmic bytes of rank 0's owned cells per stage / (max-over-ranks step time / 4)"},
        }))
    model.close()
    comm.destroy()
    dist.barrier()
    torch.cuda.synchronize()
    dist.destroy_process_group()
