"""Track-level glue (regroup, motion-state features, labels, write-back): oracle vs the reference code (CPU, when the
reference tree is mounted) and CUDA vs oracle (GPU)."""
import importlib

import numpy as np
import pytest
import torch

from helpers import synth
from oracle import codecs, crop as ocrop, refshim, trackops as otrack

tops = importlib.import_module("3dal_pytorch_b200.trackops")
DEV = "cuda:0"


def _observations(rng, n_frames=40, n_ids=60):
    ids, frames = [], []
    pool = rng.integers(10 ** 12, 10 ** 15, n_ids)
    for f in range(n_frames):
        present = rng.permutation(n_ids)[: rng.integers(0, n_ids // 2 + 2)]
        for k in present:
            ids.append(pool[k]); frames.append(f)
    return np.asarray(ids, dtype=np.int64), np.asarray(frames, dtype=np.int32)


@pytest.mark.skipif(not refshim.available(), reason="reference tree not mounted")
def test_oracle_codecs_and_transform_match_reference_sources():
    """angle2class / size2class / transform_box of the oracle against the reference's own functions."""
    rs, _, _, ut = refshim.load()
    rng = np.random.default_rng(0)
    ds = rs.STATICTRACK({}, {})
    for _ in range(50):
        box = rng.normal(0, 10, (1, 7))
        pose = np.eye(4); a = rng.uniform(-3, 3)
        pose[:2, :2] = [[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]]; pose[:3, 3] = rng.normal(0, 50, 3)
        assert np.array_equal(otrack.transform_box(box, pose), ds.transform_box(box, pose))
        lwh = rng.uniform(0.5, 12, 3).astype(np.float32)
        assert codecs.size2class(lwh)[0] == ut.size2class(lwh)[0]


def test_oracle_motion_feature_quirk():
    """np.array of a list of (1,7) boxes is (L,1,7): [0, :3] keeps every box column (tools/motionState.py:47-49)."""
    rng = np.random.default_rng(1)
    lst = [rng.normal(0, 5, (1, 7)) for _ in range(9)]
    f = otrack.motion_features([lst])[0]
    b = np.concatenate(lst, 0)
    assert np.isclose(f[0], np.linalg.norm(b[0] - b[-1])) and np.isclose(f[1], np.linalg.norm(b.var(0)))


@pytest.mark.gpu
def test_regroup_matches_dict_semantics():
    rng = np.random.default_rng(2)
    for n_frames, n_ids in [(40, 60), (200, 250), (3, 2), (1, 1)]:
        ids, frames = _observations(rng, n_frames, n_ids)
        if len(ids) == 0:
            continue
        order, table = otrack.regroup(ids, frames)
        g = tops.regroup(torch.from_numpy(ids).to(DEV), torch.from_numpy(frames).to(DEV), n_frames)
        assert g["track_id"].cpu().tolist() == order
        lens = g["track_len"].cpu().numpy(); obs = g["track_obs"].cpu().numpy()
        for t, tid in enumerate(order):
            assert obs[t, :lens[t]].tolist() == table[tid] and np.all(obs[t, lens[t]:] == -1)
        assert g["track_of_obs"].cpu().tolist() == [order.index(int(i)) for i in ids]
    with pytest.raises(ValueError):
        tops.regroup(torch.tensor([5, 5], device=DEV), torch.tensor([0, 0], dtype=torch.int32, device=DEV), 2)


@pytest.mark.gpu
def test_motion_features_and_writeback_match_oracle():
    rng = np.random.default_rng(3)
    ids, frames = _observations(rng, 50, 40)
    n_obs = len(ids)
    boxes = rng.normal(0, 20, (n_obs, 7))
    order, table = otrack.regroup(ids, frames)
    g = tops.regroup(torch.from_numpy(ids).to(DEV), torch.from_numpy(frames).to(DEV), 50)
    feat = tops.motion_features(g, torch.from_numpy(boxes).to(DEV)).cpu().numpy()
    ref = otrack.motion_features([[boxes[i][None] for i in table[t]] for t in order])
    assert np.allclose(feat, ref, rtol=1e-12, atol=1e-12)
    pred = tops.linear_svc_predict(torch.from_numpy(feat).to(DEV), [[-0.8, -0.01]], 30.0).cpu().numpy()
    assert np.array_equal(pred, (ref @ np.array([-0.8, -0.01]) + 30.0) > 0)
    # write-back
    def rand_pose():
        a = rng.uniform(-3, 3); P = np.eye(4)
        P[:2, :2] = [[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]]; P[:3, 3] = rng.normal(0, 100, 3)
        return P
    frame_pose = [rand_pose() for _ in range(50)]
    T = len(order)
    final = rng.normal(0, 5, (T, 7)).astype(np.float32)
    best = np.stack([frame_pose[frames[table[t][0]]] for t in order])
    inv = np.stack([np.linalg.inv(frame_pose[f]) for f in frames])
    out = tops.box_writeback(torch.from_numpy(final).to(DEV), torch.from_numpy(best).to(DEV), g, torch.from_numpy(inv).to(DEV)).cpu().numpy()
    for t, tid in enumerate(order):
        for o in table[tid]:
            r = otrack.transform_box(otrack.transform_box(final[[t]].astype(np.float64), best[t]), inv[o])[0]
            assert np.allclose(out[o], r, rtol=1e-12, atol=1e-9)


@pytest.mark.gpu
def test_track_labels_match_oracle():
    trackprep = importlib.import_module("3dal_pytorch_b200.trackprep")
    rng = np.random.default_rng(4)
    bs, n = 6, 2048
    rows = 9000
    gt = np.concatenate([rng.normal(0, 2, (bs, 3)), rng.uniform(1, 6, (bs, 3)), rng.uniform(-4, 4, (bs, 1))], 1).astype(np.float32)
    poses = []
    for _ in range(bs):
        a = rng.uniform(-3, 3); P = np.eye(4)
        P[:2, :2] = [[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]]; P[:3, 3] = rng.normal(0, 100, 3)
        poses.append(P)
    inv = np.stack([np.linalg.inv(P) for P in poses])
    # global-frame source points scattered around each GT box (vehicle frame -> global)
    src = np.zeros((rows, 3)); owner = rng.integers(0, bs, rows)
    local = gt[owner, :3].astype(np.float64) + rng.normal(0, 2.0, (rows, 3))
    for i in range(rows):
        src[i] = (poses[owner[i]] @ np.append(local[i], 1.0))[:3]
    choice = np.stack([rng.choice(np.nonzero(owner == b)[0], n, replace=True) for b in range(bs)]).astype(np.int64)
    init_heading = rng.uniform(-3, 3, bs)
    out = tops.track_labels(torch.from_numpy(src).to(DEV), torch.from_numpy(choice).to(DEV), torch.from_numpy(inv).to(DEV),
                            torch.from_numpy(gt).to(DEV), torch.from_numpy(init_heading).to(DEV))
    for b in range(bs):
        pv = (inv[b] @ np.concatenate([src[choice[b]].T, np.ones((1, n))], 0))[:3].T
        mask, c, hc, hr, sc, sr = otrack.static_labels(pv, gt[b], init_heading[b])
        got = out["mask_label"][b].cpu().numpy()
        # the float64 pose transform differs from numpy's BLAS in the last bit: a point may flip only if it sits on a face
        assert (got != mask).sum() <= 1, (b, int((got != mask).sum()))
        assert 0.02 < mask.mean() < 0.98
        assert np.array_equal(out["center_label"][b].cpu().numpy(), c)
        assert int(out["heading_class_label"][b]) == hc and np.float32(out["heading_residuals_label"][b].item()) == np.float32(hr)
        assert int(out["size_class_label"][b]) == sc and np.array_equal(out["size_residual_label"][b].cpu().numpy(), sr.astype(np.float32))


def test_oracle_iou3d_analytic_cases():
    a = np.array([0, 0, 0, 4, 2, 2, 0.0])
    assert abs(otrack.iou3d(a, a) - 1.0) < 1e-12
    assert otrack.iou3d(a, np.array([10, 0, 0, 4, 2, 2, 0.3])) == 0.0
    b = np.array([2, 0, 0, 4, 2, 2, 0.0])                                   # half overlap along the length
    assert abs(otrack.iou3d(a, b) - (2 * 2 * 2) / (16 + 16 - 8)) < 1e-12
    c = np.array([0, 0, 1, 4, 2, 2, np.pi / 2])                             # crossed, half the height
    assert abs(otrack.iou3d(a, c) - (2 * 2 * 1) / (16 + 16 - 4)) < 1e-9
    d = np.array([0, 0, 0, 4, 2, 2, np.pi])                                 # same footprint, flipped heading
    assert abs(otrack.iou3d(a, d) - 1.0) < 1e-9


@pytest.mark.gpu
def test_match_iou3d_against_float64_restatement():
    rng = np.random.default_rng(6)
    F = 12
    dets, frames, gts, gt_off = [], [], [], [0]
    for f in range(F):
        m = int(rng.integers(0, 9))
        g = np.concatenate([rng.uniform(-20, 20, (m, 2)), rng.normal(0.8, 0.3, (m, 1)), rng.uniform(0.6, 6, (m, 3)), rng.uniform(-4, 4, (m, 1))], 1)
        gts.append(g); gt_off.append(gt_off[-1] + m)
        for k in range(int(rng.integers(0, 12))):
            if m and rng.random() < 0.7:
                d = g[rng.integers(0, m)] + np.concatenate([rng.normal(0, 0.15, 3), rng.normal(0, 0.1, 3), rng.normal(0, 0.05, 1)])
            else:
                d = np.concatenate([rng.uniform(-20, 20, 2), [0.8], rng.uniform(0.6, 6, 3), rng.uniform(-4, 4, 1)])
            dets.append(d); frames.append(f)
    det = np.asarray(dets, np.float32); gt = np.concatenate(gts, 0).astype(np.float32)
    match, iou = tops.match_detections(torch.from_numpy(det).to(DEV), torch.tensor(frames, dtype=torch.int32, device=DEV),
                                       torch.from_numpy(gt).to(DEV), torch.tensor(gt_off, dtype=torch.int64, device=DEV))
    match, iou = match.cpu().numpy(), iou.cpu().numpy()
    n_matched = 0
    for i, (d, f) in enumerate(zip(det, frames)):
        cand = gt[gt_off[f]:gt_off[f + 1]]
        ref = np.array([otrack.iou3d(d, g) for g in cand]) if len(cand) else np.zeros(0)
        if len(cand) == 0:
            assert match[i] == -1 and iou[i] == 0.0
            continue
        assert abs(iou[i] - ref.max()) < 2e-4, (i, iou[i], ref.max())
        if ref.max() > 0.75 + 1e-3:
            assert match[i] == int(np.argmax(ref)); n_matched += 1
        elif ref.max() < 0.75 - 1e-3:
            assert match[i] == -1
    assert n_matched > 10
