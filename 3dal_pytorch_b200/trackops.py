"""Track-level glue on the device (host side of csrc/trackops.cu): regrouping per-frame detections into tracks
(tools/trackData.py:25-45), motion-state features and the linear static / dynamic split (tools/motionState.py:30-67,
127-139), training labels (tools/static_model.py:549-566) and the per-frame write-back of refined boxes
(tools/static_eval.py:84-92)."""
import numpy as np
import torch

from . import _lib, crop, ops


def _p(t):
    return None if t is None else t.data_ptr()


def regroup(ids, frame_of_obs, n_frames):
    """ids (n_obs,) i64 CUDA, frame_of_obs (n_obs,) i32 CUDA, both in the reference's iteration order ->
    dict(track_id (T,) i64, track_len (T,) i32, track_obs (T, n_frames) i32 (-1 padded), track_of_obs (n_obs,) i32).
    One host read of the track count (to size the outputs)."""
    ops._need_cuda(ids, frame_of_obs)
    dev = ids.device
    n_obs = int(ids.shape[0])
    cap = max(n_obs, 1)
    hs = 2
    while hs < 2 * max(n_obs, 1):
        hs *= 2
    i32 = lambda *s: torch.empty(s, device=dev, dtype=torch.int32)
    keys = torch.empty((hs,), device=dev, dtype=torch.int64)
    first, rank, presence = i32(hs), i32(hs), i32(cap * n_frames)
    track_of_obs, track_len, track_obs = i32(max(n_obs, 1)), i32(cap), i32(cap, n_frames)
    track_id = torch.empty((cap,), device=dev, dtype=torch.int64)
    n_tracks, error = i32(1), i32(1)
    _lib.check(_lib.lib().al3d_track_regroup(_p(ids.contiguous()), _p(frame_of_obs.contiguous()), n_obs, n_frames, cap, _p(keys), _p(first),
                                             _p(rank), hs, _p(presence), _p(track_of_obs), _p(track_id), _p(track_obs), _p(track_len),
                                             _p(n_tracks), _p(error), ops._stream()), "track_regroup")
    T, err = int(n_tracks.item()), int(error.item())
    if err == 2:
        raise ValueError("track_regroup: a tracking id appears twice in one frame")
    if err:
        raise RuntimeError("track_regroup: index out of range (code %d)" % err)
    return {"track_id": track_id[:T], "track_len": track_len[:T], "track_obs": track_obs[:T], "track_of_obs": track_of_obs[:n_obs]}


def motion_features(groups, boxes, n_cols=7):
    """boxes (n_obs, >=n_cols) f64 CUDA global-frame boxes -> (T, 2) f64 [distance, variance norm] (tools/motionState.py:47-49;
    n_cols=7 reproduces the reference's slicing of its (L,1,7) array, see include/al3d.h)."""
    T, F = groups["track_obs"].shape
    boxes = boxes.contiguous()
    assert boxes.dtype == torch.float64
    feat = torch.empty((T, 2), device=boxes.device, dtype=torch.float64)
    _lib.check(_lib.lib().al3d_motion_features(_p(groups["track_obs"]), _p(groups["track_len"]), T, F, _p(boxes), boxes.shape[1], n_cols,
                                               _p(feat), ops._stream()), "motion_features")
    return feat


def linear_svc_predict(feat, coef, intercept):
    """sklearn SVC(kernel='linear').predict for two classes: 1 (static) where w.x + b > 0 (tools/motionState.py:118-127)."""
    w = torch.as_tensor(coef, dtype=torch.float64, device=feat.device).reshape(-1)
    return ((feat * w[None, :]).sum(1) + float(intercept)) > 0


def track_labels(src_xyz, choice, inv_pose, gt_box, init_heading, want_mask=True):
    """Labels of a batch of tracks.  src_xyz (rows,3) f64, choice (bs,n) i64, inv_pose (bs,4,4) f64 as for
    trackprep.prep_points; gt_box (bs,7) f32 [x y z l w h heading] in the vehicle frame; init_heading (bs,) f64."""
    ops._need_cuda(src_xyz, choice, inv_pose, gt_box, init_heading)
    dev = gt_box.device
    bs, n = choice.shape
    gt_box = gt_box.float().contiguous()
    planes = None
    if want_mask:
        sincos = np.stack([np.sin(gt_box[:, 6].cpu().numpy()), np.cos(gt_box[:, 6].cpu().numpy())], 1).astype(np.float32)
        planes = torch.empty((bs, 6, 4), device=dev, dtype=torch.float32)
        aabb = torch.empty((bs, 6), device=dev, dtype=torch.float32)
        d_sc = torch.from_numpy(sincos).to(dev)
        _lib.check(_lib.lib().al3d_crop_box_setup(_p(gt_box), _p(d_sc), bs, crop.AABB_PAD, 1e-5, _p(planes), _p(aabb), ops._stream()),
                   "crop_box_setup")
    out = {"mask_label": torch.empty((bs, n), device=dev, dtype=torch.float32) if want_mask else None,
           "center_label": torch.empty((bs, 3), device=dev, dtype=torch.float32),
           "heading_class_label": torch.empty((bs,), device=dev, dtype=torch.int64),
           "heading_residuals_label": torch.empty((bs,), device=dev, dtype=torch.float32),
           "size_class_label": torch.empty((bs,), device=dev, dtype=torch.int64),
           "size_residual_label": torch.empty((bs, 3), device=dev, dtype=torch.float32)}
    _lib.check(_lib.lib().al3d_track_labels(_p(src_xyz.contiguous()), _p(choice.contiguous()), bs, n, _p(inv_pose.contiguous()), _p(planes),
                                            _p(gt_box), _p(init_heading.double().contiguous()), _p(out["mask_label"]), _p(out["center_label"]),
                                            _p(out["heading_class_label"]), _p(out["heading_residuals_label"]), _p(out["size_class_label"]),
                                            _p(out["size_residual_label"]), ops._stream()), "track_labels")
    return out


def box_writeback(final_box, best_pose, groups, obs_inv_pose):
    """final_box (T,7) f32, best_pose (T,4,4) f64 (veh_to_global of each track's best frame), obs_inv_pose (n_obs,4,4) f64
    (inverse pose of every observation's frame) -> (n_obs,7) f64 refined boxes in their own frames."""
    T, F = groups["track_obs"].shape
    n_obs = obs_inv_pose.shape[0]
    out = torch.zeros((n_obs, 7), device=final_box.device, dtype=torch.float64)
    _lib.check(_lib.lib().al3d_box_writeback(_p(final_box.float().contiguous()), _p(best_pose.double().contiguous()), _p(groups["track_obs"]),
                                             _p(groups["track_len"]), T, F, _p(obs_inv_pose.double().contiguous()), _p(out), ops._stream()),
               "box_writeback")
    return out


def match_detections(det, det_frame, gt, gt_off, iou_threshold=0.75):
    """det (n,7) f32 CUDA, det_frame (n,) i32, gt (m,7) f32, gt_off (F+1,) i64 -> (match (n,) i32 index of the matched GT
    box within its frame or -1, best_iou (n,) f32): det3d/datasets/waymo/waymo_common.py:173-188 (arg-max IoU3D > 0.75)."""
    ops._need_cuda(det, det_frame, gt, gt_off)
    n = det.shape[0]
    best = torch.empty((n,), device=det.device, dtype=torch.int32)
    iou = torch.empty((n,), device=det.device, dtype=torch.float32)
    _lib.check(_lib.lib().al3d_match_iou3d(_p(det.float().contiguous()), _p(det_frame.contiguous()), n, _p(gt.float().contiguous()),
                                           _p(gt_off.contiguous()), _p(best), _p(iou), ops._stream()), "match_iou3d")
    return torch.where(iou > iou_threshold, best, torch.full_like(best, -1)), iou
