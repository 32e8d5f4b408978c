"""CPU-side checks of the C-ABI boundary: the library builds/loads and exports exactly the header's
symbols; argument validation returns error codes (no compute needs a GPU here)."""
import ctypes
import os
import re

import pytest

from scade_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "scade_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(scade_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert sorted(_lib.exported_symbols()) == header_symbols()


def test_library_loads_and_exports_every_symbol():
    lib = _lib.load()
    for name in header_symbols():
        assert hasattr(lib, name), name
    assert lib.scade_version() == 100


def test_argument_validation_without_gpu():
    lib = _lib.load()
    bad = _lib.NetDesc(8, 255, 9, 0, 4)      # odd width
    assert lib.scade_mlp_workspace_bytes(ctypes.byref(bad), 1024, 0, 0) == 0
    good = _lib.NetDesc(8, 256, 9, 0, 4)
    assert lib.scade_mlp_workspace_bytes(ctypes.byref(good), 1024, 0, 0) > 1024 * 256 * 4
    assert lib.scade_mlp_packed_bytes(ctypes.byref(good)) in (92 * 16384 + 3328, 87 * 16384 + 3328)   # weight stages (pair / single form) + fp32 head table
    assert lib.scade_mlp_packed_bytes(ctypes.byref(_lib.NetDesc(4, 128, 9, 0, 4))) == 0    # fp32 path only
    st = lib.scade_raw2outputs(None, None, None, 3, None, 4, 8, None, None, None, None, None, None)
    assert st == 1 and b"raw2outputs" in lib.scade_last_error_string()
    st = lib.scade_sample_pdf(None, None, 4, 1, 8, None, 0, None, None, None)
    assert st == 1


def test_cpu_tensors_are_rejected():
    import torch
    from scade_b200 import functional as F_
    with pytest.raises(_lib.ScadeError):
        F_.raw2outputs(torch.zeros(2, 4, 4), torch.zeros(2, 4), torch.ones(2, 3))
