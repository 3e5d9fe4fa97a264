"""Cycle timeline of split_tail_kernel (CTA 0: MMA issuer + first epilogue thread).

  python scripts/split_timeline.py build     # here: compile chain_split.cu with -DAL3D_SPLIT_TIMELINE -> libal3d_timeline.so
  python scripts/split_timeline.py run [tail|chain]   # on the GPU: one forward, per-phase cycle statistics of CTA 0
"""
import ctypes
import importlib
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
LIB = os.path.join(ROOT, "3dal_pytorch_b200", "libal3d_timeline.so")


def _lib_path(which):
    return LIB if which == "tail" else LIB.replace("timeline", "timeline_chain")


def build():
    import __graft_entry__ as g
    g.build()
    objdir = os.path.join(g.CSRC, "_obj")
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = [os.path.join(objdir, f) for f in sorted(os.listdir(objdir))
            if f.endswith(".o") and ".stress." not in f and ".timeline" not in f and f != "chain_split.o"]
    for which, define in (("tail", "-DAL3D_SPLIT_TIMELINE"), ("chain", "-DAL3D_CHAIN_TIMELINE")):
        obj = os.path.join(objdir, "chain_split.timeline_%s.o" % which)
        subprocess.check_call([nvcc] + g.NVCC_COMPILE + [define, "-I", os.path.join(ROOT, "include"), "-c",
                                                       os.path.join(g.CSRC, "chain_split.cu"), "-o", obj])
        subprocess.check_call([nvcc] + g.NVCC_LINK + ["-o", _lib_path(which)] + objs + [obj, "-lcuda"])
        print("built", _lib_path(which))


def run(which="tail"):
    os.environ["AL3D_LIB"] = _lib_path(which)
    import numpy as np
    import torch
    pkg = importlib.import_module("3dal_pytorch_b200")
    synth = importlib.import_module("3dal_pytorch_b200.synth")
    sm = importlib.import_module("3dal_pytorch_b200.static_model")
    _lib = importlib.import_module("3dal_pytorch_b200._lib")
    dev = torch.device("cuda:0")
    sd = synth.random_state_dict("static_one", seed=3)
    m = sm.StaticModelOneBoxEst().to(dev).eval()
    m.load_state_dict(sd)
    m.precision = "bf16x3"
    tracks = int(os.environ.get("TRACKS", "1184"))        # 148 CTAs x 32 tiles x 8 ... enough for steady state on CTA 0
    tr = synth.static_tracks(8, n=4096, seed=3)
    pts = torch.from_numpy(tr["pts_pm"]).transpose(2, 1).to(dev)
    pts = pts.repeat((tracks + 7) // 8, 1, 1)[:tracks].contiguous()
    ib = torch.from_numpy(tr["init_box"]).to(dev).repeat((tracks + 7) // 8, 1)[:tracks].contiguous()
    for _ in range(2):
        m(pts, ib, None)
    torch.cuda.synchronize()
    host = ctypes.c_void_p()
    assert _lib.lib().al3d_tc_status_word_host(ctypes.byref(host)) == 0
    words = np.ctypeslib.as_array(ctypes.cast(host, ctypes.POINTER(ctypes.c_uint32)), shape=(16384,))
    for name, off in (("issuer", 1024), ("epilogue", 8192)):
        a = words[off:off + 3000].reshape(-1, 2).copy()
        ids, clk = a[:, 0], a[:, 1].astype(np.int64)
        n = int((ids != 0).sum())
        ids, clk = ids[:n], clk[:n]
        first = 0x100 if name == "issuer" else 0x200
        starts = [i for i in range(n) if ids[i] == first]
        print("== %s %s: %d stamps, %d tiles / units" % (which, name, n, len(starts)))
        rows = []
        for a_, b_ in zip(starts[2:-1], starts[3:]):           # skip the first two tiles (pipeline fill)
            seg_ids = ids[a_:b_ + 1]
            seg = (clk[a_:b_ + 1] - clk[a_]) & 0xFFFFFFFF
            rows.append((tuple(int(x) for x in seg_ids), seg))
        if not rows:
            continue
        key = rows[0][0]
        segs = np.array([r[1] for r in rows if r[0] == key])
        med = np.median(segs, axis=0)
        print("tiles averaged: %d; tile period median %.0f cycles" % (len(segs), med[-1]))
        for i, k in enumerate(key):
            print("  id 0x%03x  t=%7.0f  dt=%6.0f" % (k, med[i], med[i] - (med[i - 1] if i else 0)))


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build()
    else:
        run(sys.argv[2] if len(sys.argv) > 2 else "tail")
