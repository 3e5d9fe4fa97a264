"""Track-level glue.  TEST INFRASTRUCTURE (oracle): plain-Python / numpy restatements of

  regroup          tools/trackData.py:25-45   (dict keyed by tracking id, per-key lists appended in frame order)
  motion_features  tools/motionState.py:36-49 (np.array of a list of (1,7) boxes is (L,1,7); [0, :3] keeps all 7 columns)
  labels           tools/static_model.py:549-566 + tools/utils.py:53-67
  transform_box / write-back   tools/static_eval.py:30-45, 84-92
"""
import numpy as np

from . import codecs, crop


def regroup(ids, frame_of_obs):
    """-> (list of track ids in first-appearance order, {id: [observation indices in iteration order]})."""
    tracking = {}
    for i, tid in enumerate(ids):
        tracking.setdefault(int(tid), []).append(i)
    return list(tracking.keys()), tracking


def motion_features(boxes_per_track):
    """boxes_per_track: list of lists of (1,7) arrays, as trackData stores them."""
    out = []
    for lst in boxes_per_track:
        bbox = np.array(lst)                                   # (L,1,7)
        distance = np.linalg.norm(bbox[0, :3] - bbox[-1, :3])
        var = np.linalg.norm(np.var(bbox[:, :3], axis=0))
        out.append([distance, var])
    return np.array(out)


def transform_box(box, pose):
    transform = pose
    heading = box[..., -1] + np.arctan2(transform[..., 1, 0], transform[..., 0, 0])
    center = np.einsum('...ij,...nj->...ni', transform[..., 0:3, 0:3], box[..., 0:3]) + np.expand_dims(transform[..., 0:3, 3], axis=-2)
    return np.concatenate([center, box[..., 3:6], heading[..., np.newaxis]], axis=-1)


def static_labels(point_vehicle, bbox_gt, init_heading):
    """point_vehicle (n,3) f64 resampled points in the vehicle frame, bbox_gt (7,) f32, init_heading f64 -> labels."""
    mask = crop.points_in_boxes(point_vehicle, bbox_gt[np.newaxis, ...]).astype(np.float64).squeeze()
    hc, hr = codecs.angle2class(bbox_gt[-1] - init_heading, 12)
    sc, sr = codecs.size2class(bbox_gt[3:6])
    return mask, bbox_gt[:3], hc, hr, sc, sr
