"""Where the wall time of the end-to-end sweep (BASELINE.json configs[3]) goes: every phase of sweep.StaticSweep.run
bracketed by a device synchronise.  Usage (GPU box): python scripts/sweep_phases.py [--frames 200]"""
import argparse, importlib, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
crop = importlib.import_module("3dal_pytorch_b200.crop")
synth = importlib.import_module("3dal_pytorch_b200.synth")
sweep = importlib.import_module("3dal_pytorch_b200.sweep")
trackprep = importlib.import_module("3dal_pytorch_b200.trackprep")
sm = importlib.import_module("3dal_pytorch_b200.static_model")

ap = argparse.ArgumentParser(); ap.add_argument("--frames", type=int, default=200); a = ap.parse_args()
dev = "cuda:0"
frames = synth.lidar_frames(a.frames, seed=3)
for f in frames: f["points"] = torch.from_numpy(f["points"]).to(dev)
torch.manual_seed(0)
model = sm.StaticModelOneBoxEst(n_classes=3, n_channel=3).to(dev).eval()
sw = sweep.StaticSweep(model)
for _ in range(3): sw.run(frames)
torch.cuda.synchronize()
T = {}
def tick(name, t0):
    torch.cuda.synchronize(); T[name] = T.get(name, 0.0) + (time.perf_counter() - t0); return time.perf_counter()
N = 10
for _ in range(N):
    t = time.perf_counter()
    F = len(frames); Tn = int(np.asarray(frames[0]["det_boxes"]).shape[0])
    all_w = crop.detector_to_waymo(np.concatenate([np.asarray(f["det_boxes"]).reshape(-1, 7) for f in frames], 0)).reshape(F, -1, 7)
    poses = np.stack([np.asarray(f["pose"], dtype=np.float64) for f in frames])
    t = tick("box conversion (numpy)", t)
    plan = crop.CropPlan([f["points"] for f in frames], all_w, poses, device=dev)
    t = tick("CropPlan construction", t)
    res = plan.run()
    t = tick("plan.run (crop kernels + size read-back)", t)
    seg_start, seg_len = sweep.track_segments(res["offsets"], res["box_off"], F, Tn)
    rows = sweep.resample_rows(seg_start, seg_len, sw.npoints, sw.policy)
    best = seg_len.argmax(1).cpu().numpy()
    t = tick("track segments + resample rows", t)
    inv_pose = np.linalg.inv(poses[best]); init_box = all_w[best, np.arange(Tn)].astype(np.float64)
    d_inv = torch.from_numpy(np.ascontiguousarray(inv_pose)).to(dev); d_init = torch.from_numpy(np.ascontiguousarray(init_box)).to(dev)
    t = tick("best frame, inverse poses, H2D", t)
    pts = trackprep.prep_points(res["xyz_global"].contiguous(), rows.contiguous(), d_inv, d_init, heading_col=6, c_out=3)
    t = tick("track prep kernel", t)
    boxes = sw.labeler.label_device(pts.transpose(2, 1), d_init.float())
    t = tick("model forward (seg + gather + box head)", t)
tot = sum(T.values())
for k, v in T.items(): print("%-45s %7.3f ms  %5.1f %%" % (k, 1e3 * v / N, 100 * v / tot))
print("%-45s %7.3f ms" % ("sum (every phase synchronised)", 1e3 * tot / N))
