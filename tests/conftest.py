import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "mpas-ocean.jl_b200"), os.path.join(ROOT, "oracle"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_collection_modifyitems(config, items):
    """Tests of code that has not run on hardware yet (written on the simulated runtime after a round's GPU budget was spent)
    go LAST, so that on a GPU box the verified suite has reported before anything untested gets its first run."""
    items.sort(key=lambda it: 1 if it.get_closest_marker("hw_pending") else 0)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "hw_pending: exercises device code whose first hardware run is still pending (sorted last)")
    if os.environ.get("MOKAB_SIM"):
        # MOKAB_SIM=1 python -m pytest tests -m gpu: the `gpu` tests run against the HOST build of the library's own
        # sources on the simulated CUDA runtime (tests/sim) -- a logic / ordering check for containers without a GPU,
        # never used on the GPU box (tests/test_sim.py drives it from the CPU suite)
        import ctypes
        sys.path.insert(0, os.path.join(ROOT, "tests", "sim"))
        import simcuda
        from moka_b200 import _lib
        simcuda.runtime()
        _lib.bind(ctypes.CDLL(simcuda._build.LIB))
        simcuda.set_policy(os.environ.get("MOKAB_SIM_POLICY", "fifo"), int(os.environ.get("MOKAB_SIM_SEED", "1")))


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


@pytest.fixture(scope="session")
def backend():
    import moka_b200 as mb
    return mb.B200(0)


_MESH_CACHE = {}


def hex_mesh(nx, ny=None, dc=None, with_dual=True):
    """Product generator mesh + sign fields (for the oracle), cached per session."""
    import moka_b200 as mb
    import moka_oracle_c as OC
    ny = ny or nx
    dc = dc or 1.0e7 / nx
    key = (nx, ny, dc, with_dual)
    if key not in _MESH_CACHE:
        m = mb.periodic_hex(nx, ny, dc, with_dual=with_dual)
        OC.sign_index_fields(m)
        _MESH_CACHE[key] = m
    return _MESH_CACHE[key]


SIM = bool(os.environ.get("MOKAB_SIM"))


class DevBuf:
    """A device buffer for tests that move halo messages themselves: a torch CUDA tensor on the GPU box, an allocation
    of the simulated runtime under MOKAB_SIM=1.  `copy_from` is a host-ordered copy (call it between synchronisations)."""

    def __init__(self, n, dtype=np.float64):
        if SIM:
            import simcuda
            self.buf, self.t = simcuda.DeviceBuffer(max(1, n), dtype), None
        else:
            import torch
            self.t = torch.zeros(max(1, n), dtype=torch.float64 if np.dtype(dtype) == np.float64 else torch.float32, device="cuda")

    def data_ptr(self):
        return self.buf.data_ptr() if SIM else self.t.data_ptr()

    def copy_from(self, dst_off, src, src_off, n):
        if SIM:
            self.buf.numpy()[dst_off:dst_off + n] = src.buf.numpy()[src_off:src_off + n]
        else:
            self.t[dst_off:dst_off + n] = src.t[src_off:src_off + n]


def device_synchronize():
    if SIM:
        import simcuda
        simcuda.synchronize()
    else:
        import torch
        torch.cuda.synchronize()
