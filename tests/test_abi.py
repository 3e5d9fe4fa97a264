"""The C-ABI library loads on a CPU-only host and exports every symbol include/al3d.h declares."""
import ctypes
import os
import re

from helpers import ROOT, pkg  # noqa: F401
import importlib

_lib = importlib.import_module("3dal_pytorch_b200._lib")


def _declared():
    text = open(os.path.join(ROOT, "include", "al3d.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(al3d_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert _declared() == _lib.exported_symbols()


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    l = ctypes.CDLL(_lib.LIB_PATH)
    for name in _declared():
        assert hasattr(l, name), name
    assert _lib.lib().al3d_abi_version() == 3
    assert _lib.lib().al3d_last_error() is not None


def test_argument_errors_are_reported_not_thrown():
    l = _lib.lib()
    rc = l.al3d_linear_f32(None, 0, 0, 0, None, 0, None, None, 0, 0, 0, None, 0, None, None)
    assert rc != 0 and b"al3d_linear_f32" in l.al3d_last_error()


def test_models_refuse_cpu_tensors():
    import pytest
    import torch
    sm = importlib.import_module("3dal_pytorch_b200.static_model")
    m = sm.StaticModelOneBoxEst().eval()
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(torch.zeros(1, 3, 128), torch.zeros(1, 7), torch.zeros(1, 7))
