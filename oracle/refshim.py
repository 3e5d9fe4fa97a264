"""Container-only loader for the REAL reference (jacky121298/3DAL_PyTorch at /root/reference).

TEST INFRASTRUCTURE.  Used only (a) by tests/golden/make_golden.py to generate the committed
golden vectors and (b) by `-m "not gpu"` tests that pin oracle/ against the reference when
/root/reference exists.  Nothing on the GPU box imports this (the reference does not travel).

The reference does not import as-is on numpy>=1.24 / without its un-vendored deps
(SURVEY.md section 8c).  The shims below are harness-side only; reference files are untouched:
  1. np.float alias              (tools/static_model.py:557,590; tools/dynamic_model.py:485,527)
  2. Tensor.cuda -> identity     (tools/static_model.py:95,126,...: hard-coded .cuda() in forward)
  3. stub fpointnet_train.provider_fpointnet   (tools/utils.py:5)
  4. empty namespace shells for det3d, det3d.core, det3d.core.bbox so that only
     det3d/core/bbox/{geometry,box_np_ops}.py are executed (they need numpy + numba only)
"""
import importlib
import os
import sys
import types

REF_ROOT = os.environ.get("AL3D_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "tools"))


_loaded = {}


def load():
    """Returns (static_model, dynamic_model, box_np_ops, utils) reference modules (CPU)."""
    if _loaded:
        return _loaded["mods"]
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    import numpy as np
    import torch

    if not hasattr(np, "float"):
        np.float = float  # shim 1
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self  # shim 2 (CPU oracle only)

    # shim 3
    if "fpointnet_train" not in sys.modules:
        pkg = types.ModuleType("fpointnet_train")
        pkg.__path__ = []
        prov = types.ModuleType("fpointnet_train.provider_fpointnet")
        pkg.provider_fpointnet = prov
        sys.modules["fpointnet_train"] = pkg
        sys.modules["fpointnet_train.provider_fpointnet"] = prov

    # shim 4
    for name, rel in (("det3d", "det3d"), ("det3d.core", "det3d/core"), ("det3d.core.bbox", "det3d/core/bbox")):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = [os.path.join(REF_ROOT, rel)]
            sys.modules[name] = m
    geometry = importlib.import_module("det3d.core.bbox.geometry")
    box_np_ops = importlib.import_module("det3d.core.bbox.box_np_ops")
    sys.modules["det3d.core.bbox"].box_np_ops = box_np_ops
    sys.modules["det3d.core.bbox"].geometry = geometry

    tools = os.path.join(REF_ROOT, "tools")
    if tools not in sys.path:
        sys.path.insert(0, tools)
    # The repo's own drop-in modules are also called static_model / dynamic_model; make sure the
    # names resolve to the reference here and are cached under private aliases.
    for n in ("static_model", "dynamic_model", "utils"):
        sys.modules.pop(n, None)
    static_model = importlib.import_module("static_model")
    dynamic_model = importlib.import_module("dynamic_model")
    utils = importlib.import_module("utils")
    for n, m in (("static_model", static_model), ("dynamic_model", dynamic_model), ("utils", utils)):
        sys.modules["_al3d_ref_" + n] = m
        sys.modules.pop(n, None)
    sys.path.remove(tools)
    _loaded["mods"] = (static_model, dynamic_model, box_np_ops, utils)
    return _loaded["mods"]
