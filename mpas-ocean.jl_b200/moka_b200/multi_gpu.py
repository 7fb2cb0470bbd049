"""One process per GPU: domain-decomposed RK4 with the halo exchange overlapped with interior compute.

The reference has no multi-device path; this follows BASELINE.json's north_star:
owned/halo cell and edge layers (partition.py), one packed message per neighbour and RK stage
exchanged with NCCL over NVLink (`torch.distributed` is the plumbing), overlapped with the interior
blocks of the same stage:

    per stage s:   compute stream:  wait X[s-1] -> BOUNDARY blocks(s) -> record B[s] -> INTERIOR blocks(s)
                   comm stream:     wait B[s] -> pack(s) -> all_to_all -> unpack(s) -> record X[s]

BOUNDARY blocks are those whose stencils read a halo entity or that hold an entity a neighbour needs
(mokab_halo_setup), so the message of stage s leaves while the bulk of stage s is still computing and
is in place before stage s+1 touches the halo.  Reductions are per-rank partial sums over owned
entities + all_reduce.
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
    """Packed all-to-all of halo messages (works for NCCL/CUDA and gloo/CPU tensors)."""

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
    """What DecomposedModel needs from the device runtime and the process group: streams / events / graph capture
    (`cuda`: torch.cuda), the halo all-to-all and two scalar reductions (torch.distributed, NCCL).  The simulation tests
    (tests/sim) pass an object of the same shape backed by their host runtime, so the schedule below is exercised
    under adversarial stream interleavings without a GPU."""

    def __init__(self, device_index: int, group=None, device=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group = torch, dist, group
        self.cuda = torch.cuda
        self.dev = torch.device("cuda", device_index) if device is None else torch.device(device)   # "cpu": gloo tests of the host exchanges

    def stream(self, priority: int = 0):
        return self.torch.cuda.Stream(self.dev, priority=priority)

    def exchanger(self, send_counts, recv_counts, npdtype):
        tdt = self.torch.float64 if np.dtype(npdtype) == np.float64 else self.torch.float32
        return HaloExchanger(send_counts, recv_counts, tdt, self.dev, self.group)

    def all_reduce_min(self, value: int) -> int:
        t = self.torch.tensor([int(value)], dtype=self.torch.int32, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN, group=self.group)
        return int(t.item())

    def all_reduce_sum(self, value: float) -> float:
        t = self.torch.tensor([float(value)], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, group=self.group)
        return float(t.item())

    def rank_and_size(self):
        return self.dist.get_rank(self.group), self.dist.get_world_size(self.group)

    def sum_arrays(self, a: np.ndarray) -> np.ndarray:
        """Element-wise sum of a host array over the ranks (output gathering: every rank fills its owned entries of a zero array)."""
        t = self.torch.from_numpy(np.ascontiguousarray(a, np.float64)).to(self.dev)
        self.dist.all_reduce(t, group=self.group)
        return t.cpu().numpy()

    # set-up traffic of the direct-store halo exchange (once per model; tiny)
    def all_gather_bytes(self, blob: bytes) -> list:
        n = self.dist.get_world_size(self.group)
        mine = self.torch.frombuffer(bytearray(blob), dtype=self.torch.uint8).to(self.dev)
        out = [self.torch.empty_like(mine) for _ in range(n)]
        self.dist.all_gather(out, mine, group=self.group)
        return [bytes(t.cpu().numpy().tobytes()) for t in out]

    def all_to_all_int32(self, send: list, recv_counts: list) -> list:
        """send[q]: int32 array for rank q; returns the arrays received from every rank (recv_counts[q] elements)."""
        torch = self.torch
        s = torch.from_numpy(np.concatenate([np.asarray(a, np.int32) for a in send] + [np.zeros(0, np.int32)])).to(self.dev)
        r = torch.empty(int(sum(recv_counts)), dtype=torch.int32, device=self.dev)
        self.dist.all_to_all_single(r, s, [int(c) for c in recv_counts], [int(np.asarray(a).size) for a in send], group=self.group)
        r = r.cpu().numpy()
        offs = np.concatenate([[0], np.cumsum(recv_counts)]).astype(np.int64)
        return [r[offs[q]:offs[q + 1]].copy() for q in range(len(recv_counts))]


def plan_steps(nsteps: int, parity: int, graph_parity: int):
    """How `nsteps` steps are issued when a 2-step graph captured at time-level parity `graph_parity` exists and the
    state currently has parity `parity` (steps taken so far mod 2): (stream-launched steps first, graph replays,
    stream-launched steps after).  The graph's kernels have the time-level buffers of its capture parity baked in, so
    it may only be replayed from that parity."""
    pre = 1 if (parity & 1) != (graph_parity & 1) and nsteps >= 1 else 0
    rest = nsteps - pre
    return pre, rest // 2, rest % 2


class DecomposedModel:
    """This rank's share of the mesh on its GPU + the stage/exchange schedule.

    Two streams per rank: `compute` runs the INTERIOR blocks, the high-priority `halo` stream runs the
    BOUNDARY blocks, then pack -> all-to-all -> unpack.  Stage s on either stream needs stage s-1 of BOTH
    (events `ev_i`, `ev_b`); the unpack of stage s-1 precedes the boundary launch of stage s on the same
    stream.  Both parts of a stage therefore start together, the boundary blocks win the SMs first, and
    the message travels while the interior blocks run.  With `graph=True` two consecutive steps (one per
    time-level parity) are captured -- NCCL calls included -- into one CUDA graph and replayed.
    """

    def __init__(self, loc: dict, state, backend: api.B200, device_index: int, dtype=np.float64, group=None, overlap=True,
                 graph=False, runtime=None, halo="nccl"):
        if halo not in ("nccl", "p2p", "p2p_fused"):
            raise api.MokaError("DecomposedModel: halo must be 'nccl' (packed all-to-all), 'p2p' (direct peer stores, push and wait "
                                "kernels) or 'p2p_fused' (direct peer stores from inside the boundary launch)")
        self.halo_mode = halo
        self.rt = runtime if runtime is not None else TorchRuntime(device_index, group)
        self.cuda = self.rt.cuda
        self.loc, self.backend, self.overlap, self.use_graph = loc, backend, overlap, graph
        self.nparts = loc["nparts"]
        self.mesh = api.Mesh(loc, backend)
        sidx, scnt, ridx, rcnt = partition.flat_halo(loc, self.nparts)
        self.mesh.halo_setup(sidx, ridx)
        ssh, u, h = state
        self.prog = api.PrognosticVars(np.asarray(ssh, dtype), np.asarray(u, dtype), np.asarray(h, dtype), 2, self.mesh)
        self.dev = getattr(self.rt, "dev", None)
        self.ex = self.rt.exchanger(scnt, rcnt, dtype)
        self.compute = self.rt.stream()
        self.halo = self.rt.stream(priority=-1)
        self.comm = self.halo
        # the context's own work (state set/get permutes, reductions, ssh refresh) joins the compute stream, so the
        # pipelined upload/download of the API orders itself with the steps without host synchronisation
        backend.set_stream(self.compute.cuda_stream)
        self.handle = self.prog.dev.handle
        self._graph, self._graph_dt = None, None
        self._validated, self.graph_status = False, "not used"
        self._parity, self._graph_parity = 0, 0                  # steps taken so far mod 2; the same at graph capture
        self._fe, self._stepped = False, False                   # the model steps with ForwardEuler; it has stepped
        if halo != "nccl":
            self._setup_p2p(scnt, rcnt)

    def _setup_p2p(self, scnt, rcnt) -> None:
        """Direct-store halo exchange (csrc/kernels_p2p.cuh): tell every sender where its values live in this rank's
        arrays, exchange the addresses / IPC handles of the state arrays, and build the push tables."""
        lib, rank = L.lib(), int(self.loc["rank"])
        nrecv_total = int(sum(rcnt))
        mine = np.zeros(max(1, nrecv_total), np.int32)
        L.check(lib.mokab_halo_recv_device_indices(self.mesh.handle, mine.ctypes.data_as(L._I32P)))
        offs = np.concatenate([[0], np.cumsum(rcnt)]).astype(np.int64)
        # rank q fills my segment q: it needs those indices; I need, from every rank I send to, its segment for me
        got = self.rt.all_to_all_int32([mine[offs[q]:offs[q + 1]] for q in range(self.nparts)], list(scnt))
        size = C.c_int64()
        L.check(lib.mokab_p2p_blob_size(C.byref(size)))
        blob = C.create_string_buffer(size.value)
        L.check(lib.mokab_p2p_export(self.handle, rank, blob))
        blobs = b"".join(self.rt.all_gather_bytes(blob.raw))
        receivers = [q for q in range(self.nparts) if scnt[q] > 0]
        senders = [q for q in range(self.nparts) if rcnt[q] > 0]
        dst = np.ascontiguousarray(np.concatenate([got[q] for q in receivers] + [np.zeros(0, np.int32)]), np.int32)
        rr, sr = np.asarray(receivers, np.int32), np.asarray(senders, np.int32)
        cnt = np.asarray([scnt[q] for q in receivers], np.int64)
        L.check(lib.mokab_p2p_setup(self.handle, rank, self.nparts, blobs, len(receivers), rr.ctypes.data_as(L._I32P),
                                    cnt.ctypes.data_as(C.POINTER(C.c_int64)), dst.ctypes.data_as(L._I32P), len(senders),
                                    sr.ctypes.data_as(L._I32P)))
        self.rt.all_reduce_min(1)                                # nobody pushes before everybody is mapped

    def _stage(self, dt, s, part, stream):
        L.check(L.lib().mokab_rk4_stage(self.handle, float(dt), s, part, C.c_void_p(stream.cuda_stream)))

    def _exchange(self, s, stream):
        lib = L.lib()
        if self.halo_mode == "p2p":
            L.check(lib.mokab_halo_push(self.handle, s, C.c_void_p(stream.cuda_stream)))
            L.check(lib.mokab_halo_wait(self.handle, C.c_void_p(stream.cuda_stream)))
            return
        with self.cuda.stream(stream):
            L.check(lib.mokab_halo_pack(self.handle, s, C.c_void_p(self.ex.send.data_ptr()), C.c_void_p(stream.cuda_stream)))
            self.ex.exchange()
            L.check(lib.mokab_halo_unpack(self.handle, s, C.c_void_p(self.ex.recv.data_ptr()), C.c_void_p(stream.cuda_stream)))

    def _wait_arrivals(self, stream) -> None:
        L.check(L.lib().mokab_halo_wait_arrivals(self.handle, C.c_void_p(stream.cuda_stream)))

    def _enqueue_steps(self, dt: float, nsteps: int) -> None:
        """Enqueue `nsteps` RK4 steps; on entry and exit both streams are joined on `compute`."""
        cuda = self.cuda
        fused = self.halo_mode == "p2p_fused"                    # the boundary launch carries the exchange itself
        boundary = L.PART_BOUNDARY_PUSH if fused else L.PART_BOUNDARY
        if not self.overlap:
            for _ in range(nsteps):
                for s in (1, 2, 3, 4):
                    if fused:
                        self._stage(dt, s, boundary, self.compute)
                        self._stage(dt, s, L.PART_INTERIOR, self.compute)
                    else:
                        self._stage(dt, s, L.PART_ALL, self.compute)
                        self._exchange(s, self.compute)
                L.check(L.lib().mokab_rk4_finish_step(self.handle))
            if fused:
                self._wait_arrivals(self.compute)
            return
        self.halo.wait_stream(self.compute)                      # fork
        for _ in range(nsteps):
            for s in (1, 2, 3, 4):
                ev_i, ev_b = cuda.Event(), cuda.Event()
                self._stage(dt, s, boundary, self.halo)
                ev_b.record(self.halo)
                self._stage(dt, s, L.PART_INTERIOR, self.compute)
                ev_i.record(self.compute)
                if not fused:
                    self._exchange(s, self.halo)
                self.compute.wait_event(ev_b)                    # stage s+1 interior reads stage s boundary output
                self.halo.wait_event(ev_i)                       # stage s+1 boundary reads stage s interior output
            L.check(L.lib().mokab_rk4_finish_step(self.handle))
        if fused:
            self._wait_arrivals(self.halo)                       # the neighbours' last stores, before anything else touches the halo slots
        self.compute.wait_stream(self.halo)                      # join

    def _enqueue_fe_steps(self, dt: float, nsteps: int) -> None:
        """Enqueue `nsteps` ForwardEuler steps (the reference's live stepper, time_integration.jl:150-193): one launch per part,
        then the halo copies of everything the step wrote -- (h, u) and (ssh, layerThicknessEdge), two messages -- while the
        interior blocks run.  Stream-launched (one exchange pair per step instead of RK4's four; no captured graph)."""
        if self.halo_mode != "nccl":
            raise api.MokaError("DecomposedModel: ForwardEuler steps use the packed exchange (halo='nccl')")
        lib, cuda = L.lib(), self.cuda
        self._fe = True

        def stage(part, stream):
            L.check(lib.mokab_forward_euler_stage(self.handle, float(dt), part, C.c_void_p(stream.cuda_stream)))

        if not self.overlap:
            for _ in range(nsteps):
                stage(L.PART_ALL, self.compute)
                self._exchange(4, self.compute)
                self._exchange(5, self.compute)
                L.check(lib.mokab_forward_euler_finish_step(self.handle))
            return
        self.halo.wait_stream(self.compute)                      # fork
        for _ in range(nsteps):
            ev_i, ev_b = cuda.Event(), cuda.Event()
            stage(L.PART_BOUNDARY, self.halo)
            ev_b.record(self.halo)
            stage(L.PART_INTERIOR, self.compute)
            ev_i.record(self.compute)
            self._exchange(4, self.halo)
            self._exchange(5, self.halo)
            self.compute.wait_event(ev_b)                        # the next interior launch overwrites what this boundary launch read
            self.halo.wait_event(ev_i)                           # the next boundary launch reads (and overwrites the inputs of) this interior launch
            L.check(lib.mokab_forward_euler_finish_step(self.handle))
        self.compute.wait_stream(self.halo)                      # join

    def _build_graph(self, dt: float) -> None:
        """Capture two consecutive steps (one per time-level parity), NCCL calls included, into one CUDA graph."""
        cuda = self.cuda
        self.compute.synchronize()
        self.halo.synchronize()
        self._enqueue_steps(dt, 2)                               # warm up NCCL + lazy library state outside capture
        self.compute.synchronize()
        g = cuda.CUDAGraph()
        with cuda.graph(g, stream=self.compute):
            self._enqueue_steps(dt, 2)
        self._graph, self._graph_dt, self._graph_parity = g, dt, self._parity

    def _snapshot(self):
        return [self.prog.dev.get(f) for f in (L.SSH, L.NORMAL_VELOCITY, L.LAYER_THICKNESS)]

    def _restore(self, snap) -> None:
        for f, a in zip((L.SSH, L.NORMAL_VELOCITY, L.LAYER_THICKNESS), snap):
            self.prog.dev.set(f, a)

    def validate_graph(self, dt: float, nsteps: int = 8) -> bool:
        """Build the 2-step graph and accept it only if replaying it reproduces, bit for bit on every rank, what the
        stream-launched schedule computes from the same state (`nsteps` even; the state is restored afterwards).
        A captured schedule that also contains NCCL traffic is not something to trust unchecked: on rejection the
        model keeps launching from the host (`use_graph` False, reason in `graph_status`)."""
        snap = self._snapshot()
        self._enqueue_steps(dt, nsteps)
        self.finish()
        want = self._snapshot()
        self._restore(snap)
        self._build_graph(dt)                                    # advances the state by its two warm-up steps
        with self.cuda.stream(self.compute):
            for _ in range((nsteps - 2) // 2):
                self._graph.replay()
        self.finish()
        got = self._snapshot()
        self._restore(snap)
        ok = all(np.array_equal(a, b) for a, b in zip(want, got))
        flag = self.rt.all_reduce_min(1 if ok else 0)
        self.compute.synchronize()
        self._validated = True
        if flag == 0:
            self._graph, self._graph_dt, self.use_graph = None, None, False
            self.graph_status = "rejected: graph replay did not reproduce the stream-launched schedule; launching from the host"
            import gc
            gc.collect()
            return False
        self.graph_status = f"validated against the stream-launched schedule over {nsteps} steps"
        return True

    def _run(self, dt: float, nsteps: int) -> None:
        """Stream-launched steps (outside any capture), keeping track of the time-level parity."""
        self._enqueue_steps(dt, nsteps)
        self._parity ^= nsteps & 1

    def step(self, dt: float, nsteps: int = 1, stepper=None) -> None:
        """`nsteps` steps of `stepper` (api.RungeKutta4, the default, or api.ForwardEuler)."""
        if stepper not in (None, api.RungeKutta4, api.ForwardEuler):
            raise api.MokaError("DecomposedModel.step: unknown stepper")
        fe = stepper is api.ForwardEuler
        if self._stepped and fe != self._fe:                     # (ForwardEuler carries a lagged layerThicknessEdge between its steps)
            raise api.MokaError("DecomposedModel.step: one stepper per model -- build another DecomposedModel to change it")
        self._stepped = self._stepped or nsteps > 0
        if fe:
            if nsteps > 0:
                self._enqueue_fe_steps(dt, nsteps)
                self._parity ^= nsteps & 1
            return
        if self.use_graph and nsteps >= 2:
            if self._graph is None or self._graph_dt != dt:
                if not self.validate_graph(dt):
                    self._run(dt, nsteps)
                    return
            pre, replays, nsteps = plan_steps(nsteps, self._parity, self._graph_parity)
            if pre:
                self._run(dt, pre)
            with self.cuda.stream(self.compute):
                for _ in range(replays):
                    self._graph.replay()
        if nsteps:
            self._run(dt, nsteps)

    def refresh_ssh(self) -> None:
        """ssh = layerThickness - restingThicknessSum on the compute stream (asynchronous)."""
        L.check(L.lib().mokab_refresh_ssh(self.handle, C.c_void_p(self.compute.cuda_stream)))

    def finish(self) -> None:
        if not self._fe:                                         # (ForwardEuler carries ssh as a state of its own)
            self.refresh_ssh()
        self.compute.synchronize()
        self.halo.synchronize()
        if self.halo_mode != "nccl":
            err = C.c_int()
            L.check(L.lib().mokab_p2p_error(self.handle, C.byref(err)))
            if err.value:
                raise api.MokaError("DecomposedModel: a halo wait timed out (a peer died or the ranks' schedules diverged)")

    def close(self) -> None:
        """Drop the captured graph (it pins NCCL resources: the process group cannot be destroyed while
        it is alive) and drain both streams."""
        self.compute.synchronize()
        self.halo.synchronize()
        if self.halo_mode != "nccl":                             # nobody unmaps while a neighbour may still store, nobody frees while mapped
            self.rt.all_reduce_min(1)
            L.check(L.lib().mokab_p2p_close(self.handle))
            self.rt.all_reduce_min(1)
        self.backend.set_stream(None)                            # the context goes back to its own stream
        self._graph = None
        import gc
        gc.collect()
        self.cuda.synchronize()

    def owned(self, field: str) -> np.ndarray:
        a = getattr(self.prog, field)
        n = self.loc["nCellsOwned"] if field in ("ssh", "layerThickness") else self.loc["nEdgesOwned"]
        return a[:n]

    def reduce(self, which: str) -> float:
        return self.rt.all_reduce_sum(api.reduce_sum(self.prog, which))

    def gather(self, nC: int, nE: int, previous: bool = False):
        """The global (ssh, normalVelocity, layerThickness) on every rank, assembled from the owned parts (`previous`: the
        time level one step back, which is what the reference's write_netcdf writes, PrognosticVars.jl:108-113)."""
        loc, out = self.loc, []
        names = (("ssh_prev", "normalVelocity_prev", "layerThickness_prev") if previous else ("ssh", "normalVelocity", "layerThickness"))
        for name, n, ids, no in ((names[0], nC, loc["cellsGlobal"], loc["nCellsOwned"]), (names[1], nE, loc["edgesGlobal"], loc["nEdgesOwned"]),
                                 (names[2], nC, loc["cellsGlobal"], loc["nCellsOwned"])):
            g = np.zeros(n, np.float64)
            g[ids[:no]] = np.asarray(getattr(self.prog, name), np.float64)[:no]
            out.append(self.rt.sum_arrays(g))
        return out


def local_state(loc: dict, ssh, u, h):
    return ssh[loc["cellsGlobal"]], u[loc["edgesGlobal"]], h[loc["cellsGlobal"]]


def gather_owned(model: DecomposedModel, nC: int, nE: int):
    """Assemble the global (ssh, u, h) on every rank from the owned parts (test/diagnostic helper)."""
    import torch
    import torch.distributed as dist
    loc = model.loc
    out = []
    for field, n, ids, no in (("ssh", nC, loc["cellsGlobal"], loc["nCellsOwned"]),
                              ("normalVelocity", nE, loc["edgesGlobal"], loc["nEdgesOwned"]),
                              ("layerThickness", nC, loc["cellsGlobal"], loc["nCellsOwned"])):
        g = torch.zeros(n, dtype=torch.float64, device=model.dev)
        g[torch.as_tensor(ids[:no], device=model.dev)] = torch.as_tensor(np.asarray(model.owned(field), np.float64), device=model.dev)
        dist.all_reduce(g)
        out.append(g.cpu().numpy())
    return out


# ---- bench leg for torchrun (N > 1) -------------------------------------------------------------------------
def _share_locals(args, rank, world, nx, dtype):
    """Rank 0 builds the global mesh, decomposes it and hands every rank its local mesh through /dev/shm."""
    import torch.distributed as dist
    tag = f"/dev/shm/mokab_{os.environ.get('MASTER_PORT', '0')}_{nx}_{world}"
    t0 = time.time()
    if rank == 0:
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
        locs = partition.decompose(m, world)
        for r, loc in enumerate(locs):
            ls = local_state(loc, ssh, u, h)
            flat = {k: v for k, v in loc.items() if isinstance(v, np.ndarray)}
            meta = {k: v for k, v in loc.items() if not isinstance(v, (np.ndarray, dict))}
            if "bench_dt" in m:
                meta["bench_dt"] = m["bench_dt"]
            halo = loc["halo"]
            for q in halo["peers"]:
                flat[f"halo_send_{q}"], flat[f"halo_recv_{q}"] = halo["send"][q], halo["recv"][q]
            meta["peers"] = halo["peers"]
            flat["state_ssh"], flat["state_u"], flat["state_h"] = ls
            np.savez(f"{tag}_{r}.npz", meta=json.dumps(meta), **flat)
        del m, locs
    dist.barrier()
    z = np.load(f"{tag}_{rank}.npz")
    meta = json.loads(str(z["meta"]))
    loc = {k: z[k] for k in z.files if k != "meta" and not k.startswith(("halo_", "state_"))}
    loc.update({k: v for k, v in meta.items() if k != "peers"})
    loc["halo"] = {"peers": meta["peers"], "send": {q: z[f"halo_send_{q}"] for q in meta["peers"]},
                   "recv": {q: z[f"halo_recv_{q}"] for q in meta["peers"]}}
    state = (z["state_ssh"], z["state_u"], z["state_h"])
    dist.barrier()
    if rank == 0:
        for r in range(world):
            try:
                os.remove(f"{tag}_{r}.npz")
            except OSError:
                pass
    return loc, state, time.time() - t0


def bench_main(args, rank, world, local):
    import torch
    import torch.distributed as dist
    from bench import WORKLOADS, ClockSampler, algo_bytes_per_cell_step, algo_bytes_per_cell_step_general, measured_peak_gbs
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nx = WORKLOADS[args.workload]
    npdt = np.float64 if args.dtype == "f64" else np.float32
    loc, state, t_setup = _share_locals(args, rank, world, nx, npdt)
    nC_glob = nx * nx
    sphere = args.workload.startswith("sphere")
    voronoi = args.workload.startswith("voronoi") or sphere      # unstructured: byte accounting from the actual rows
    dt = float(loc["bench_dt"]) if sphere else (0.25 if voronoi else 1.0) * api.cfl_dt(1.0e7 / nx)
    backend = api.B200(local)
    model = DecomposedModel(loc, state, backend, local, dtype=npdt, overlap=not getattr(args, "no_overlap", False),
                            graph=not getattr(args, "no_graph", False), halo=getattr(args, "halo", "nccl"))
    K, W = args.steps, max(args.warmup, 3)
    model.step(dt, W)
    model.finish()
    dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(model.compute)
    model.step(dt, K)
    e1.record(model.compute)
    model.compute.synchronize()
    model.halo.synchronize()
    ms_local = e0.elapsed_time(e1)
    # this library's kernels per step: 4 stages x (boundary + interior + pack + unpack), or 4 x (all + pack + unpack);
    # graph replays do not pass through the host-side counter, so the count is by construction
    launches = K * (16 if model.overlap else 12)
    t = torch.tensor([ms_local], dtype=torch.float64, device=model.dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    dist.barrier()
    torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    model.finish()

    # end to end with HOST buffers: every step uploads this rank's (u, h) from pinned memory and reads back ssh;
    # pipelined through the API's copy streams (PCIe transfers of steps n+1 / n-1 overlap the kernels of step n)
    nCl, nEl = loc["nCells"], loc["nEdges"]
    hin = [(backend.pinned(nEl, npdt), backend.pinned(nCl, npdt)) for _ in range(2)]
    hout = [backend.pinned(nCl, npdt) for _ in range(2)]
    for hu, hh in hin:
        hu[:], hh[:] = np.asarray(state[1], npdt), np.asarray(state[2], npdt)
    Ke = max(3, min(K, 20))

    def e2e_steps(n):
        for i in range(n):
            model.prog.upload_async(normalVelocity=hin[i & 1][0], layerThickness=hin[i & 1][1])
            model.step(dt, 1)
            model.refresh_ssh()
            model.prog.download_async(ssh=hout[i & 1])
        model.prog.synchronize()
        model.halo.synchronize()
    e2e_steps(2)
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e2e_steps(Ke)
    torch.cuda.synchronize()
    te = torch.tensor([(time.perf_counter() - t0) / Ke], dtype=torch.float64, device=model.dev)
    dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())
    cnt = torch.tensor([nCl + nEl, nCl, loc["nCellsOwned"], launches], dtype=torch.float64, device=model.dev)
    dist.all_reduce(cnt)
    blocks = model.mesh.block_counts()
    halo_bytes = (sum(model.ex.send_counts) + sum(model.ex.recv_counts)) * np.dtype(npdt).itemsize
    if rank == 0:
        item = np.dtype(npdt).itemsize
        peak, peak_src = measured_peak_gbs()
        value = nC_glob * K / (ms * 1e-3)
        nblk, nder = model.mesh.derived_blocks()
        if voronoi:
            nco, neo = loc["nCellsOwned"], loc["nEdgesOwned"]
            owned = {"nCells": nco, "nEdges": neo, "nEdgesOnCell": loc["nEdgesOnCell"][:nco], "nEdgesOnEdge": loc["nEdgesOnEdge"][:neo]}
            per_cell_step = algo_bytes_per_cell_step_general(owned, args.dtype, nder / max(nblk, 1))
        else:
            per_cell_step = algo_bytes_per_cell_step(args.dtype, nder / max(nblk, 1))
        algo_per_launch = per_cell_step / 4.0 * loc["nCellsOwned"]
        achieved = algo_per_launch / ((ms * 1e-3) / (4 * K)) / 1e9
        print(json.dumps({
            "metric": "RK4 cell-steps/sec", "value": value, "unit": "cell-steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": args.dtype,
            "data": "synthetic",
            "config": {"workload": ("coastal Kelvin wave, %dx%d channel hex mesh with boundary-edge masks" % (nx, nx) if args.workload.startswith("kelvin")
                                    else "geostrophic zonal flow + noise, spherical Voronoi mesh, fEdge = 2 Omega sin(lat)" if sphere
                                    else f"inertial gravity wave, {nx}x{nx} periodic planar Voronoi mesh of a jittered lattice" if voronoi
                                    else f"inertial gravity wave, {nx}x{nx} periodic planar hex mesh") + f" ({nC_glob} cells), "
                                   f"{'Float64' if args.dtype == 'f64' else 'Float32'} RK4, dt={dt:.4g}s, recursive-coordinate-bisection "
                                   f"into {world} parts, 1 halo layer, "
                                   f"{ {'nccl': 'NCCL all-to-all', 'p2p': 'direct peer stores + arrival counters (push / wait kernels)', 'p2p_fused': 'direct peer stores from inside the boundary launch'}[model.halo_mode]} per stage "
                                   f"{'overlapped with interior blocks' if model.overlap else '(no overlap)'}"
                                   f"{', 2-step CUDA graph incl. the exchange (' + model.graph_status + ')' if model.use_graph else ', CUDA graph ' + model.graph_status}",
                       "name": args.workload, "l2": "inputs larger than L2 (no flush)", "setup_s": round(t_setup, 1),
                       "rank0_blocks_interior_boundary": list(blocks), "rank0_halo_bytes_per_stage": int(halo_bytes),
                       "rank0_blocks_rebuilding_edgesOnEdge": [int(nder), int(nblk)],
                       "variant": {"lib": os.environ.get("MOKAB_LIB", "libmoka_b200.so"),
                                   "stage_tma": int(os.environ.get("MOKAB_STAGE_TMA", "0") or 0)}},
            "clocks": clocks,
            "e2e": {"value": nC_glob / e2e_s, "unit": "cell-steps/s", "h2d_bytes_per_step": int(cnt[0].item() * item),
                    "d2h_bytes_per_step": int(cnt[1].item() * item), "ms_per_step": e2e_s * 1e3, "steps": Ke,
                    "pipelined": True},
            "gpu_launches": int(cnt[3].item()),
            "roofline": {"bound": "hbm", "kernel": "k_rk_stage", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                         "note": "per GPU: algorithmic bytes of rank 0's owned cells per stage / (max-over-ranks step time / 4)"},
        }))
    model.close()
    dist.barrier()
    torch.cuda.synchronize()
    dist.destroy_process_group()
