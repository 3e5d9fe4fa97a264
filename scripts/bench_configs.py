"""Secondary measurements for the other BASELINE.json configs (one JSON line each):
  crop      configs[3], crop stage: 200 Waymo-shaped frames x ~180k points x 200 boxes -> achieved HBM GB/s
  dynamic   configs[1]: dynamic model forward, 64 tracks x (5x1024 points + 101 boxes), plus a large batch
  static32  configs[0] on the GPU: static one- and two-box forward, 32 tracks x 4096 points (fp32 and bf16)
  sweep     configs[3] end to end: crop -> regroup -> track prep -> static model -> decoded boxes
Usage (GPU box): python scripts/bench_configs.py [crop|sweep|dynamic|static32|all] [--frames N]"""
import argparse
import importlib
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import __graft_entry__ as ge

ge.build()
synth = importlib.import_module("3dal_pytorch_b200.synth")
crop = importlib.import_module("3dal_pytorch_b200.crop")
sm = importlib.import_module("3dal_pytorch_b200.static_model")
dm = importlib.import_module("3dal_pytorch_b200.dynamic_model")
spec = importlib.import_module("3dal_pytorch_b200.spec")
graphs = importlib.import_module("3dal_pytorch_b200.graphs")
PEAKS = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))) if os.path.exists(
    os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops_sustained": 1400.0}
DEV = "cuda:0"


def timed(fn, warmup=3, iters=10):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def bench_crop(n_frames):
    frames = synth.lidar_frames(n_frames, seed=3)
    pts = [torch.from_numpy(f["points"]).to(DEV) for f in frames]
    boxes = [crop.detector_to_waymo(f["det_boxes"]) for f in frames]
    poses = [f["pose"] for f in frames]
    plan = crop.CropPlan(pts, boxes, poses, device=DEV)
    res = plan.run()
    total = int(res["offsets"][-1].item())
    assert int(res["overflow"].item()) == 0
    ms = timed(lambda: plan.run())
    # cumulative prefixes of the valid launch sequence (the scan rewrites its input in place, so stages cannot
    # be replayed on their own); per-stage time = difference of consecutive prefixes
    seqs = [("grid",), ("grid", "hits_pass"), ("grid", "hits_pass", "scan"), ("grid", "hits_pass", "scan", "fill")]
    cum = [timed(lambda q=q: [getattr(plan, k)() for k in q]) for q in seqs]
    stage_ms = {"grid": cum[0], "hits": cum[1] - cum[0], "scan": cum[2] - cum[1], "fill": cum[3] - cum[2]}
    n_points = int(plan.n_points)
    alg_bytes = plan.read_bytes + total * (4 + 12)                  # + 4 B index + 12 B xyz per inside point
    moved = alg_bytes + total * 24                                  # the f64 global xyz is extra output
    # CPU reference for the same job: oracle (C restatement of the numba loop), one thread, on a 2-frame sample
    from oracle import crop as ocrop
    t0 = time.perf_counter()
    for f in range(min(2, n_frames)):
        ocrop.crop_frame(frames[f]["points"], boxes[f], poses[f])
    cpu_s_per_frame = (time.perf_counter() - t0) / min(2, n_frames)
    print(json.dumps({"bench": "crop", "frames": n_frames, "points": n_points, "boxes": int(plan.TB), "inside": total,
                      "ms": ms, "stage_ms": stage_ms, "frames_per_s": n_frames / (ms * 1e-3), "algorithmic_bytes": alg_bytes,
                      "achieved_gbs": alg_bytes / (ms * 1e-3) / 1e9, "achieved_gbs_incl_f64_output": moved / (ms * 1e-3) / 1e9,
                      "hbm_peak_gbs": PEAKS["hbm_gbs"], "frac_of_measured_hbm": alg_bytes / (ms * 1e-3) / 1e9 / PEAKS["hbm_gbs"],
                      "cpu_oracle_s_per_frame_1thread": cpu_s_per_frame,
                      "speedup_vs_cpu_oracle": cpu_s_per_frame * n_frames / (ms * 1e-3)}))


def bench_sweep(n_frames):
    """configs[3]: crop -> regroup -> track prep -> seg -> gather -> box head -> decode, inputs resident on the device."""
    sweep = importlib.import_module("3dal_pytorch_b200.sweep")
    frames = synth.lidar_frames(n_frames, seed=4)
    for f in frames:
        f["points"] = torch.from_numpy(f["points"]).to(DEV)
    sd = synth.random_state_dict("static_one", seed=synth.REFERENCE_SEED)
    m = sm.StaticModelOneBoxEst().to(DEV).eval()
    m.load_state_dict(sd)
    m.precision = "bf16x3"
    sw = sweep.StaticSweep(m)
    out = sw.run(frames)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    iters = 3
    for _ in range(iters):
        out = sw.run(frames)
    torch.cuda.synchronize()
    ms = 1e3 * (time.perf_counter() - t0) / iters
    print(json.dumps({"bench": "sweep", "frames": n_frames, "tracks": int(out["boxes"].shape[0]),
                      "points_per_frame": int(frames[0]["points"].shape[0]), "ms_wall_incl_host_plane_setup": ms,
                      "frames_per_s": n_frames / (ms * 1e-3), "boxes_per_s": n_frames * 200 / (ms * 1e-3)}))


def _calibrated(kind, cls, pts, aux):
    sd = synth.random_state_dict(kind, seed=synth.REFERENCE_SEED)
    m = cls().to(DEV).eval()
    m.load_state_dict(sd)
    m.precision = "bf16x3"
    with torch.no_grad():
        lg = m(pts[:64], aux[:64], aux[:64] if kind != "dynamic" else None)["logits"]
    synth.calibrate_seg_margin(sd, lg, fg_fraction=0.125 if kind != "dynamic" else 0.5)
    m.load_state_dict(sd)
    return m


def bench_dynamic():
    for bs in (64, 4096):
        tr = synth.dynamic_tracks(min(bs, 256), seed=2)
        rep = -(-bs // tr["pts_pm"].shape[0])
        pts = torch.from_numpy(np.tile(tr["pts_pm"], (rep, 1, 1))[:bs]).to(DEV).transpose(2, 1)
        box = torch.from_numpy(np.tile(tr["box_sm"], (rep, 1, 1))[:bs]).to(DEV).transpose(2, 1)
        m = _calibrated("dynamic", dm.DynamicModel, pts, box)
        for prec in ("mixed", "bf16x3", "bf16", "fp32") if bs == 64 else ("mixed", "bf16x3", "bf16"):
            m.precision = prec
            ms = timed(lambda: m(pts, box, None), iters=5)
            fl = spec.flops_per_object("dynamic", 5120)
            row = {"bench": "dynamic", "tracks": bs, "precision": prec, "ms": ms, "objects_per_s": bs / (ms * 1e-3),
                   "model_tflops": bs * fl / (ms * 1e-3) / 1e12}
            if prec == "bf16x3" and bs == 64:
                g = graphs.GraphedForward(m, pts, box, None)
                row["ms_cuda_graph"] = timed(lambda: g(pts, box, None), iters=10)
            print(json.dumps(row))


def bench_static32():
    tr = synth.static_tracks(32, seed=1)
    pts = torch.from_numpy(tr["pts_pm"]).to(DEV).transpose(2, 1)
    ib, gt = torch.from_numpy(tr["init_box"]).to(DEV), torch.from_numpy(tr["bbox_gt"]).to(DEV)
    for kind, cls in (("static_one", sm.StaticModelOneBoxEst), ("static_two", sm.StaticModelTwoBoxEst)):
        m = _calibrated(kind, cls, pts, ib)
        for prec in ("mixed", "bf16x3", "bf16", "fp32"):
            m.precision = prec
            ms = timed(lambda: m(pts, ib, gt), iters=10)
            row = {"bench": "static32", "model": kind, "tracks": 32, "precision": prec, "ms": ms, "objects_per_s": 32 / (ms * 1e-3)}
            if prec == "bf16x3":
                g = graphs.GraphedForward(m, pts, ib, gt)
                row["ms_cuda_graph"] = timed(lambda: g(pts, ib, gt), iters=10)
            print(json.dumps(row))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("what", nargs="?", default="all")
    ap.add_argument("--frames", type=int, default=200)
    a = ap.parse_args()
    if a.what in ("crop", "all"):
        bench_crop(a.frames)
    if a.what in ("sweep", "all"):
        bench_sweep(a.frames)
    if a.what in ("dynamic", "all"):
        bench_dynamic()
    if a.what in ("static32", "all"):
        bench_static32()
