"""Development aid: clock64 timeline of CTA 0 of seg_pass2_kernel / seg_pass1_kernel (four tiles / tile pairs, starting at
tile AL3D_DEBUG_SKIP of the CTA; AL3D_TIMELINE_TRACKS tracks, default 1024).
Usage (GPU box): python scripts/pass2_timeline.py > gpurun_out/timeline.txt"""
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import __graft_entry__ as ge

ge.build()
eb = importlib.import_module("3dal_pytorch_b200.engine_bf16")
eng = importlib.import_module("3dal_pytorch_b200.engine")
lib = importlib.import_module("3dal_pytorch_b200._lib")
sm = importlib.import_module("3dal_pytorch_b200.static_model")
synth = importlib.import_module("3dal_pytorch_b200.synth")

dev = "cuda:0"
model = sm.StaticModelOneBoxEst().to(dev).eval()
model.load_state_dict(synth.random_state_dict("static_one", seed=1))
model.precision = "bf16"
data = synth.static_tracks_device(int(os.environ.get("AL3D_TIMELINE_TRACKS", "1024")), seed=0, device=dev)
pts = data["pts_pm"].transpose(2, 1)
for _ in range(2):
    model(pts, data["init_box"], None)
dbg = torch.zeros(2 * 3 * 4 * 64, dtype=torch.int64, device=dev)
lib.check(lib.lib().al3d_set_debug_buffer(dbg.data_ptr()), "set_debug_buffer")
model(pts, data["init_box"], None)
torch.cuda.synchronize()
lib.check(lib.lib().al3d_set_debug_buffer(None), "set_debug_buffer")
dd = dbg.cpu().view(2, 3, 4, 64)
for k, kern in enumerate(["seg_pass2_kernel", "seg_pass1_kernel"]):
    d = dd[k]
    if not bool((d > 0).any()):
        continue
    t0 = int(d[d > 0].min())
    print("==== " + kern)
    for role, name in enumerate(["mma", "epilogue", "epilogue(CTA 1)"]):
        for it in range(4):
            ts = [int(v) - t0 for v in d[role, it] if v > 0]
            if not ts:
                continue
            print(name, "item", it, "n=%d" % len(ts), "start=%d" % ts[0], "end=%d" % ts[-1])
            print("   deltas:", [b - a for a, b in zip(ts, ts[1:])])
