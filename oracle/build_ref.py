"""TEST INFRASTRUCTURE.  Builds oracle/_ref/: pieces of the REAL reference that compile from their own source files.

  iou3d_cpu   det3d/ops/iou3d_nms/src/iou3d_cpu.cpp (rotated BEV IoU on the CPU; needs only the torch C++ headers that are
              in this image) + oracle/refbind_iou3d.cpp (our pybind declaration of its one entry point).

The sources are compiled where they lie under the reference tree (AL3D_REFERENCE_ROOT, default /root/reference); outputs go
to oracle/_ref/ only (git-ignored, but shipped to the GPU box with the snapshot).  tests/golden/make_golden_iou.py uses the
module to generate tests/golden/iou_bev_ref.npz, which is what pins oracle/trackops.iou3d where the tree is absent."""
import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF_ROOT = os.environ.get("AL3D_REFERENCE_ROOT", "/root/reference")
NAME = "al3d_ref_iou3d_cpu"


def load_iou3d_cpu(build=True):
    """-> the extension module, or None when it is not built and cannot be (no reference tree)."""
    so = os.path.join(OUT, NAME + ".so")
    if os.path.exists(so):
        import torch  # noqa: F401  (the module links against libtorch)
        spec = importlib.util.spec_from_file_location(NAME, so)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod
    src = os.path.join(REF_ROOT, "det3d", "ops", "iou3d_nms", "src", "iou3d_cpu.cpp")
    if not build or not os.path.exists(src):
        return None
    from torch.utils.cpp_extension import load
    os.makedirs(OUT, exist_ok=True)
    return load(name=NAME, sources=[src, os.path.join(HERE, "refbind_iou3d.cpp")], extra_include_paths=[os.path.dirname(src), "/usr/local/cuda/include"],
                extra_cflags=["-O2", "-w"], build_directory=OUT, verbose=False)


if __name__ == "__main__":
    m = load_iou3d_cpu()
    print("oracle/_ref:", "built " + NAME if m is not None else "reference tree not found, nothing built")
    sys.exit(0)
