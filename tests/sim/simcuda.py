"""`torch.cuda`-shaped facade over the simulated runtime (tests/sim/sim_runtime.cpp) + the simulation controls.
TEST INFRASTRUCTURE.  Only the handful of names moka_b200.multi_gpu uses exist: Stream, Event, CUDAGraph, graph(),
stream(), synchronize(), current_stream(); and a device-buffer type with `data_ptr()` for the halo messages."""
from __future__ import annotations

import contextlib
import ctypes as C
import os
import sys
import threading

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import build as _build  # noqa: E402

FIFO, LAZY, OTHERS_FIRST, RANDOM = range(4)
POLICIES = {"fifo": FIFO, "lazy": LAZY, "others_first": OTHERS_FIRST, "random": RANDOM}

_rt = None


def runtime():
    """The loaded simulation library (built on first use)."""
    global _rt
    if _rt is None:
        L = C.CDLL(_build.build())
        vp, i64 = C.c_void_p, C.c_int64
        for name, args in {
            "cudaStreamCreateWithPriority": [C.POINTER(vp), C.c_uint, C.c_int], "cudaStreamDestroy": [vp],
            "cudaStreamSynchronize": [vp], "cudaStreamWaitEvent": [vp, vp, C.c_uint], "cudaEventCreate": [C.POINTER(vp)],
            "cudaEventRecord": [vp, vp], "cudaEventSynchronize": [vp], "cudaDeviceSynchronize": [],
            "cudaStreamBeginCapture": [vp, C.c_int], "cudaStreamEndCapture": [vp, C.POINTER(vp)],
            "cudaGraphInstantiate": [C.POINTER(vp), vp, C.c_ulonglong], "cudaGraphDestroy": [vp], "cudaGraphExecDestroy": [vp],
            "cudaGraphLaunch": [vp, vp], "cudaMalloc": [C.POINTER(vp), C.c_size_t], "cudaFree": [vp],
            "cudaMemcpyAsync": [vp, vp, C.c_size_t, C.c_int, vp], "cudaMemsetAsync": [vp, C.c_int, C.c_size_t, vp],
            "mokab_sim_all_to_all": [vp, C.c_int, vp, vp, vp, C.POINTER(i64), C.POINTER(i64), i64],
        }.items():
            fn = getattr(L, name)
            fn.argtypes, fn.restype = args, C.c_int
        L.cudaGetErrorString.argtypes, L.cudaGetErrorString.restype = [C.c_int], C.c_char_p
        L.mokab_sim_set_policy.argtypes, L.mokab_sim_set_policy.restype = [C.c_int, C.c_uint64], None
        L.mokab_sim_stats.argtypes, L.mokab_sim_stats.restype = [C.POINTER(i64)], None
        L.mokab_sim_set_thread_device.argtypes, L.mokab_sim_set_thread_device.restype = [C.c_int], None
        L.mokab_sim_set_deadlock_seconds.argtypes, L.mokab_sim_set_deadlock_seconds.restype = [C.c_double], None
        L.mokab_sim_comm_create.argtypes, L.mokab_sim_comm_create.restype = [C.c_int], vp
        L.mokab_sim_comm_destroy.argtypes, L.mokab_sim_comm_destroy.restype = [vp], None
        _rt = L
    return _rt


class SimCudaError(RuntimeError):
    pass


def _ck(rc: int) -> None:
    if rc != 0:
        raise SimCudaError(runtime().cudaGetErrorString(rc).decode())


def set_policy(policy, seed: int = 1) -> None:
    runtime().mokab_sim_set_policy(POLICIES[policy] if isinstance(policy, str) else int(policy), seed)


def stats() -> dict:
    a = (C.c_int64 * 6)()
    runtime().mokab_sim_stats(a)
    return dict(zip(("kernels", "ops", "graph_nodes", "collectives", "pending", "device_bytes"), [int(x) for x in a]))


# ---- torch.cuda look-alikes -------------------------------------------------------------------------------------
class Stream:
    def __init__(self, device=None, priority: int = 0):
        h = C.c_void_p()
        _ck(runtime().cudaStreamCreateWithPriority(C.byref(h), 1, priority))
        self.cuda_stream = h.value

    def synchronize(self) -> None:
        _ck(runtime().cudaStreamSynchronize(self.cuda_stream))

    def wait_event(self, ev: "Event") -> None:
        _ck(runtime().cudaStreamWaitEvent(self.cuda_stream, ev.handle, 0))

    def wait_stream(self, other: "Stream") -> None:
        ev = Event()
        ev.record(other)
        self.wait_event(ev)


_tls = threading.local()


def current_stream() -> Stream:
    return getattr(_tls, "stream", None) or _default_stream()


def _default_stream() -> Stream:
    if not hasattr(_tls, "default"):
        _tls.default = Stream()
    return _tls.default


@contextlib.contextmanager
def stream(s: Stream):
    prev = getattr(_tls, "stream", None)
    _tls.stream = s
    try:
        yield
    finally:
        _tls.stream = prev


class Event:
    def __init__(self, enable_timing: bool = False):
        h = C.c_void_p()
        _ck(runtime().cudaEventCreate(C.byref(h)))
        self.handle = h.value

    def record(self, s: Stream | None = None) -> None:
        _ck(runtime().cudaEventRecord(self.handle, (s or current_stream()).cuda_stream))

    def synchronize(self) -> None:
        _ck(runtime().cudaEventSynchronize(self.handle))


def synchronize() -> None:
    _ck(runtime().cudaDeviceSynchronize())


class CUDAGraph:
    def __init__(self):
        self.exec = None

    def replay(self) -> None:
        _ck(runtime().cudaGraphLaunch(self.exec, current_stream().cuda_stream))

    def __del__(self):
        if self.exec is not None and _rt is not None:
            _rt.cudaGraphExecDestroy(self.exec)
            self.exec = None


@contextlib.contextmanager
def graph(g: CUDAGraph, stream: Stream | None = None, capture_error_mode: str = "thread_local"):
    """Like torch.cuda.graph: capture what the body enqueues on `stream` (and on streams forked from it).  Emulated
    ranks are host THREADS of one process here, so the default error mode is thread_local (with real ranks each
    process has its own runtime and torch's "global" mode sees only that rank's threads)."""
    s = stream or current_stream()
    prev = getattr(_tls, "stream", None)
    _tls.stream = s
    _ck(runtime().cudaStreamBeginCapture(s.cuda_stream, {"global": 0, "thread_local": 1, "relaxed": 2}[capture_error_mode]))
    try:
        yield
    except BaseException:
        h = C.c_void_p()
        runtime().cudaStreamEndCapture(s.cuda_stream, C.byref(h))
        if h.value:
            runtime().cudaGraphDestroy(h.value)
        raise
    finally:
        _tls.stream = prev
    h = C.c_void_p()
    _ck(runtime().cudaStreamEndCapture(s.cuda_stream, C.byref(h)))
    ge = C.c_void_p()
    _ck(runtime().cudaGraphInstantiate(C.byref(ge), h.value, 0))
    _ck(runtime().cudaGraphDestroy(h.value))
    g.exec = ge.value


class DeviceBuffer:
    """A "device" allocation of the simulated runtime (poisoned like every cudaMalloc there) with torch's data_ptr()."""

    def __init__(self, n: int, dtype):
        self.dtype, self.n = np.dtype(dtype), int(n)
        h = C.c_void_p()
        _ck(runtime().cudaMalloc(C.byref(h), max(1, self.n) * self.dtype.itemsize))
        self.ptr = h.value

    def data_ptr(self) -> int:
        return self.ptr

    def numpy(self) -> np.ndarray:
        """A live view (only meaningful once the streams writing it have been synchronised)."""
        return np.ctypeslib.as_array(C.cast(self.ptr, C.POINTER(C.c_uint8)), shape=(self.n * self.dtype.itemsize,)).view(self.dtype)

    def __del__(self):
        if self.ptr and _rt is not None:
            _rt.cudaFree(self.ptr)
            self.ptr = None


class Comm:
    """The emulated ranks' communicator: in-stream all-to-all with NCCL's completion semantics (sim_runtime.cpp)."""

    def __init__(self, nranks: int):
        self.n = nranks
        self.handle = runtime().mokab_sim_comm_create(nranks)
        self._barrier = threading.Barrier(nranks)
        self._slots = [None] * nranks

    def all_to_all(self, rank: int, send: DeviceBuffer, recv: DeviceBuffer, send_counts, recv_counts) -> None:
        sc = (C.c_int64 * self.n)(*[int(x) for x in send_counts])
        rc = (C.c_int64 * self.n)(*[int(x) for x in recv_counts])
        _ck(runtime().mokab_sim_all_to_all(self.handle, rank, current_stream().cuda_stream, send.data_ptr(), recv.data_ptr(), sc, rc,
                                           send.dtype.itemsize))

    def exchange_host(self, rank: int, value) -> list:
        """Every rank thread contributes a value; all get the list (a host-side all-gather)."""
        self._slots[rank] = value
        self._barrier.wait()
        out = list(self._slots)
        self._barrier.wait()
        return out

    def all_reduce_host(self, rank: int, value, op=min):
        """Host-side reduction between the rank threads (stands in for a blocking all_reduce of a scalar)."""
        self._slots[rank] = value
        self._barrier.wait()
        out = self._slots[0]
        for v in self._slots[1:]:
            out = op(out, v)
        self._barrier.wait()
        return out


class SimExchanger:
    """multi_gpu.HaloExchanger's shape on the simulated runtime: packed send / receive buffers + the in-stream
    all-to-all of the emulated ranks' communicator."""

    def __init__(self, comm: Comm, rank: int, send_counts, recv_counts, npdtype):
        self.comm, self.rank, self.group = comm, rank, None
        self.send_counts, self.recv_counts = [int(x) for x in send_counts], [int(x) for x in recv_counts]
        self.send = DeviceBuffer(max(1, sum(self.send_counts)), npdtype)
        self.recv = DeviceBuffer(max(1, sum(self.recv_counts)), npdtype)

    def exchange(self) -> None:
        self.comm.all_to_all(self.rank, self.send, self.recv, self.send_counts, self.recv_counts)


class SimRuntime:
    """multi_gpu.TorchRuntime's shape for one emulated rank (a host thread of this process)."""

    def __init__(self, comm: Comm, rank: int):
        self.comm, self.rank = comm, rank
        self.cuda = sys.modules[__name__]
        self.dev = None

    def stream(self, priority: int = 0) -> Stream:
        return Stream(priority=priority)

    def exchanger(self, send_counts, recv_counts, npdtype) -> SimExchanger:
        return SimExchanger(self.comm, self.rank, send_counts, recv_counts, npdtype)

    def all_reduce_min(self, value: int) -> int:
        return int(self.comm.all_reduce_host(self.rank, int(value), min))

    def all_reduce_sum(self, value: float) -> float:
        return float(self.comm.all_reduce_host(self.rank, float(value), lambda a, b: a + b))

    def rank_and_size(self):
        return self.rank, self.comm.n

    def broadcast_bytes(self, blob, src: int = 0):
        return self.comm.exchange_host(self.rank, blob)[src]

    def sum_arrays(self, a):
        parts = self.comm.exchange_host(self.rank, np.array(a, np.float64))
        out = parts[0].copy()
        for p in parts[1:]:
            out = out + p
        return out

    def all_gather_bytes(self, blob: bytes) -> list:
        return self.comm.exchange_host(self.rank, blob)

    def all_to_all_int32(self, send: list, recv_counts: list) -> list:
        everybody = self.comm.exchange_host(self.rank, [np.array(a, np.int32) for a in send])
        out = [everybody[q][self.rank] for q in range(self.comm.n)]
        assert [a.size for a in out] == [int(c) for c in recv_counts], "all_to_all_int32: counts do not match"
        return out


def run_ranks(nranks: int, body):
    """Run body(rank, comm) on `nranks` host threads (the emulated processes); returns their results, re-raises the first
    exception."""
    comm = Comm(nranks)
    out, err = [None] * nranks, [None] * nranks

    def work(r):
        try:
            runtime().mokab_sim_set_thread_device(r + 1)     # every emulated rank drives its own device
            out[r] = body(r, comm)
        except BaseException as e:  # noqa: BLE001
            err[r] = e
            comm._barrier.abort()

    ts = [threading.Thread(target=work, args=(r,)) for r in range(nranks)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    for e in err:
        if e is not None:
            raise e
    return out
