"""End-to-end sweep (configs[3]) at small scale: CUDA pipeline vs the oracle chain, stage by stage."""
import importlib

import numpy as np
import pytest
import torch

from helpers import rel_err, synth
from oracle import crop as ocrop, models, trackprep as otp

pytestmark = pytest.mark.gpu
sweep = importlib.import_module("3dal_pytorch_b200.sweep")
crop = importlib.import_module("3dal_pytorch_b200.crop")
sm = importlib.import_module("3dal_pytorch_b200.static_model")


def test_static_sweep_matches_oracle_chain():
    dev = "cuda:0"
    F, T, NP = 5, 10, 1024
    frames = synth.lidar_frames(F, n_points=20000, n_boxes=T, seed=21)
    sd = synth.random_state_dict("static_one", seed=9)
    model = sm.StaticModelOneBoxEst().to(dev).eval()
    model.load_state_dict(sd)
    model.precision = "fp32"
    out = sweep.StaticSweep(model, npoints=NP, policy="strided").run(frames)
    # ---- oracle chain on the CPU
    boxes_w = [ocrop.detector_to_waymo(f["det_boxes"]) for f in frames]
    per_track = [[] for _ in range(T)]
    for f in range(F):
        _, xyz = ocrop.crop_frame(frames[f]["points"], boxes_w[f], frames[f]["pose"])
        for t in range(T):
            per_track[t].append(xyz[t])
    best = np.array([int(np.argmax([len(a) for a in per_track[t]])) for t in range(T)])
    assert np.array_equal(best, out["best"])
    ref_pts, ref_init = [], []
    for t in range(T):
        merged = np.vstack(per_track[t])
        n = len(merged)
        choice = (np.arange(NP) * n) // NP
        pose = np.linalg.inv(frames[best[t]]["pose"])
        bbox = boxes_w[best[t]][t].astype(np.float64)[None]
        p = (pose @ np.concatenate([merged.T, np.ones((1, n))], 0))[:3].T[choice]
        p = (otp.rotz(-bbox[0, -1]) @ (p - bbox[:, :3]).T).T
        ref_pts.append(p); ref_init.append(bbox[0])
    ref_pts = np.stack(ref_pts)
    got_pts = out["pts"].cpu().numpy()
    assert np.allclose(got_pts, ref_pts.astype(np.float32), rtol=0, atol=1e-4)       # crop indices + resample rows exact, f64 transform
    init = torch.from_numpy(np.stack(ref_init)).float()
    ref = models.static_one_forward(sd, torch.from_numpy(got_pts).transpose(2, 1), init, policy="strided")
    ref_box, _, _ = models.decode_box(ref["center"], ref["heading_scores"], ref["heading_residuals"], ref["size_scores"],
                                      ref["size_residuals"], init[:, 6])
    assert rel_err(out["boxes"].cpu(), torch.from_numpy(ref_box).float()) < 1e-4
