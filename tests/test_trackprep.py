"""Track preparation: oracle vs the reference Dataset classes (CPU, when mounted) and the CUDA
kernels vs the oracle (GPU)."""
import importlib
import os
import pickle

import numpy as np
import pytest
import torch

from oracle import refshim, trackprep as otp

tp = importlib.import_module("3dal_pytorch_b200.trackprep")


def _pose(rng):
    yaw = rng.uniform(-np.pi, np.pi)
    P = np.eye(4)
    P[:2, :2] = [[np.cos(yaw), -np.sin(yaw)], [np.sin(yaw), np.cos(yaw)]]
    P[:3, 3] = [rng.uniform(-300, 300), rng.uniform(-300, 300), rng.uniform(-3, 3)]
    return P


def _static_case(rng, bs, npoints):
    tracks = []
    for _ in range(bs):
        n = int(rng.integers(1, 3000))
        P = _pose(rng)
        ctr = P[:3, 3] + rng.normal(0, 20, 3)
        pts = ctr + rng.normal(0, 2, (n, 3))
        box = np.concatenate([ctr + rng.normal(0, 0.3, 3), rng.uniform(1, 8, 3), [rng.uniform(-3, 3)]])
        tracks.append((pts, P, box))
    return tracks


@pytest.mark.skipif(not refshim.available(), reason="reference tree not mounted")
def test_oracle_static_item_equals_reference_dataset(tmp_path):
    sm, _, _, _ = refshim.load()
    rng = np.random.default_rng(0)
    pts, P, box = _static_case(rng, 1, 4096)[0]
    gt = np.concatenate([box[:3] * 0 + 1.0, [4.5, 1.9, 1.6], [0.0, 0.0], [0.3]]).astype(np.float32)   # (9,) like waymo_decoder
    anno = {"veh_to_global": P.reshape(-1), "objects": [{"name": "obj0", "box": gt}]}
    path = tmp_path / "anno.pkl"
    pickle.dump(anno, open(path, "wb"))
    half = len(pts) // 2
    track = {"t0": {"bbox": [box[None], box[None] + 0.1], "point": [pts[:half], pts[half:]], "score": [0.9, 0.5],
                    "token": ["tokA", "tokB"], "match": ["obj0"]}}
    infos = {"tokA": {"anno_path": str(path)}, "tokB": {"anno_path": str(path)}}
    ds = sm.STATICTRACK(track, infos)
    np.random.seed(11)
    item = ds[0]
    np.random.seed(11)
    choice = np.random.choice(len(pts), 4096, replace=True)
    p, bbox = otp.static_item(pts, choice, P, box)
    assert np.array_equal(item[3].numpy(), p) and np.array_equal(item[1].numpy(), bbox)


@pytest.mark.gpu
def test_static_prep_matches_oracle():
    rng = np.random.default_rng(1)
    bs, npts = 9, 4096
    tracks = _static_case(rng, bs, npts)
    counts = [len(t[0]) for t in tracks]
    offsets = np.concatenate([[0], np.cumsum(counts)])[:-1]
    np.random.seed(5)
    choice = tp.resample_choice_static(counts, offsets, npts, "numpy_legacy")
    np.random.seed(5)
    ref_pts, ref_box = [], []
    for (pts, P, box) in tracks:
        ch = np.random.choice(len(pts), npts, replace=True)
        p, bb = otp.static_item(pts, ch, P, box)
        ref_pts.append(p); ref_box.append(bb[0])
    dev = "cuda:0"
    src = torch.from_numpy(np.concatenate([t[0] for t in tracks])).to(dev)
    inv_pose = torch.from_numpy(np.stack([np.linalg.inv(t[1]) for t in tracks])).to(dev)
    init_box = torch.from_numpy(np.stack(ref_box)).to(dev)
    out = tp.prep_points(src, torch.from_numpy(choice).to(dev), inv_pose, init_box)
    ref = np.stack(ref_pts)
    got = out.cpu().numpy()
    assert got.shape == (bs, npts, 3) and got.dtype == np.float32
    # float64 arithmetic in another association order, then one rounding to f32: a few f32 ulps at most
    assert np.allclose(got, ref.astype(np.float32), rtol=0, atol=1e-5 * np.abs(ref).max())
    strided = tp.resample_choice_static(counts, offsets, npts, "strided")
    assert strided.min() >= 0 and np.all(np.diff(strided, axis=1) >= 0)


@pytest.mark.gpu
def test_dynamic_prep_matches_oracle():
    rng = np.random.default_rng(2)
    bs, npts = 6, 1024
    dev = "cuda:0"
    all_pts, fc, fo, boxes, poses, ref_p, ref_b, ref_i = [], [], [], [], [], [], [], []
    row = 0
    np.random.seed(9)
    for i in range(bs):
        P = _pose(rng)
        ctr = P[:3, 3] + rng.normal(0, 20, 3)
        frames, counts, offs = [], [], []
        for j in range(5):
            n = 0 if rng.random() < 0.25 else int(rng.integers(1, 900))
            pts = ctr + rng.normal(0, 2, (n, 3))
            frames.append(pts if n else None); counts.append(n); offs.append(row)
            all_pts.append(pts); row += n
        b = np.zeros((101, 8))
        lo, hi = int(rng.integers(0, 40)), int(rng.integers(60, 102))
        b[lo:hi, :3] = ctr + np.cumsum(rng.normal(0, 0.1, (hi - lo, 3)), 0)
        b[lo:hi, 3:6] = rng.uniform(1, 5, 3)
        b[lo:hi, 6] = rng.uniform(-3, 3) + np.cumsum(rng.normal(0, 0.01, hi - lo))
        b[:, 7] = 0.1 * (np.arange(101) - 50)
        boxes.append(b); poses.append(P); fc.append(counts); fo.append(offs)
    choice = tp.resample_choice_dynamic(fc, fo, npts, "numpy_legacy")
    np.random.seed(9)
    row = 0
    for i in range(bs):
        fr, chs = [], []
        for j in range(5):
            n = fc[i][j]
            pts = all_pts[i * 5 + j]
            fr.append(pts if n else None)
            chs.append(np.random.choice(n, npts, replace=True) if n else None)
        p, bb, ib = otp.dynamic_item(fr, chs, boxes[i], poses[i])
        ref_p.append(p); ref_b.append(bb); ref_i.append(ib)
    src = torch.from_numpy(np.concatenate([a.reshape(-1, 3) for a in all_pts])).to(dev)
    inv_pose = torch.from_numpy(np.stack([np.linalg.inv(P) for P in poses])).to(dev)
    box_out, init_box = tp.prep_boxseq(torch.from_numpy(np.stack(boxes)).to(dev), inv_pose)
    assert np.allclose(init_box.cpu().numpy(), np.stack(ref_i), rtol=1e-12, atol=1e-9)
    rb = np.stack(ref_b)
    assert np.allclose(box_out.cpu().numpy(), rb.astype(np.float32), rtol=0, atol=1e-5 * np.abs(rb).max())
    pts_out = tp.prep_points(src, torch.from_numpy(choice).to(dev), inv_pose, init_box, heading_col=6, c_out=4,
                             time_block=npts, time_center=2)
    rp = np.stack(ref_p)
    assert np.allclose(pts_out.cpu().numpy(), rp.astype(np.float32), rtol=0, atol=1e-5 * np.abs(rp[..., :3]).max())
    assert np.array_equal(pts_out.cpu().numpy()[..., 3], rp[..., 3].astype(np.float32))      # time channel exact
