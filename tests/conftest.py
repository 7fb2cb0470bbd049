import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "mpas-ocean.jl_b200"), os.path.join(ROOT, "oracle"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


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
