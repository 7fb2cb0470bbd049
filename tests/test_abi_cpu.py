"""CPU: the C-ABI library loads, exports every symbol include/moka_b200.h declares, and fails loudly
(no CPU fallback) when there is no CUDA device."""
import ctypes
import os
import re

import pytest

import moka_b200 as mb
from moka_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "moka_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mokab_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported():
    assert os.path.exists(_lib.LIB_PATH), "build libmoka_b200.so first (python __graft_entry__.py)"
    L = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 29
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/moka_b200.h but not exported"
    assert sorted(_lib.SYMBOLS) == names, "python binding and header disagree on the symbol list"


def test_version_and_error_string():
    L = _lib.lib()
    assert L.mokab_version() == 100
    assert isinstance(L.mokab_last_error(), bytes)


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(mb.MokaError, match="no CPU fallback"):
        mb.B200(0)


def test_host_argument_checks_mirror_reference():
    import numpy as np
    with pytest.raises(mb.MokaError, match="same eltype"):
        mb.check_eltype_args((np.zeros(3), np.zeros(3, np.float32)))
    with pytest.raises(mb.MokaError, match="same type"):
        mb.check_typeof_args((np.zeros(3), [0.0]))


def test_product_does_not_import_oracle():
    """The product package and bench's own arm must not route through oracle/."""
    pkg = os.path.join(ROOT, "mpas-ocean.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".jl")):
                src = open(os.path.join(dirpath, f)).read()
                assert "moka_oracle" not in src and "mesh_oracle" not in src, f
