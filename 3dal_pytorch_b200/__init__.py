"""B200-native implementation of the 3DAL object-centric auto-labeling hot path.

Import with ``importlib.import_module("3dal_pytorch_b200")`` (the name starts with a digit), or put
this directory on ``sys.path`` and use the reference's own module names:
``from static_model import StaticModelOneBoxEst`` / ``from dynamic_model import DynamicModel``.
"""
from . import spec, synth  # noqa: F401

__all__ = ["spec", "synth"]
