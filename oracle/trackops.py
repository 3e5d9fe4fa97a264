"""Track-level glue.  TEST INFRASTRUCTURE (oracle): plain-Python / numpy restatements of

  regroup          tools/trackData.py:25-45   (dict keyed by tracking id, per-key lists appended in frame order)
  motion_features  tools/motionState.py:36-49 (np.array of a list of (1,7) boxes is (L,1,7); [0, :3] keeps all 7 columns)
  labels           tools/static_model.py:549-566 + tools/utils.py:53-67
  transform_box / write-back   tools/static_eval.py:30-45, 84-92
  iou3d            det3d/ops/iou3d_nms/iou3d_nms_utils.py:35-72 (rotated BEV overlap x height overlap / union volume)

Pinning: iou3d is checked against the REAL reference's CPU rotated BEV IoU (det3d/ops/iou3d_nms/src/iou3d_cpu.cpp:232,
compiled from its source by oracle/build_ref.py into oracle/_ref) -- live where that module is built, and through the
golden file tests/golden/iou_bev_ref.npz everywhere (tests/test_oracle_iou_ref.py): < 6e-7 on all but the near-degenerate
pairs, where the reference's float32 corner test carries a 1e-2 margin (worst 2.7e-4).  The other functions are pinned by
the fixtures generated from the reference's Python (tests/golden/make_golden.py).
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


def _rect_corners(b):
    c, s = np.cos(b[6]), np.sin(b[6])
    loc = np.array([[0.5, 0.5], [0.5, -0.5], [-0.5, -0.5], [-0.5, 0.5]]) * b[3:5]
    return np.stack([b[0] + loc[:, 0] * c + loc[:, 1] * s, b[1] - loc[:, 0] * s + loc[:, 1] * c], 1)


def _clip(poly, a, b):
    """Keep the part of `poly` on the left of the directed line a -> b (float64 Sutherland-Hodgman)."""
    out = []
    n = len(poly)
    side = lambda p: (b[0] - a[0]) * (p[1] - a[1]) - (b[1] - a[1]) * (p[0] - a[0])
    for i in range(n):
        p, q = poly[i], poly[(i + 1) % n]
        sp, sq = side(p), side(q)
        if sp >= 0:
            out.append(p)
        if (sp >= 0) != (sq >= 0):
            t = sp / (sp - sq)
            out.append(p + t * (q - p))
    return out


def iou3d(a, b):
    """Rotated 3-D IoU of two [x y z l w h heading] boxes, float64 (det3d/ops/iou3d_nms/iou3d_nms_utils.py:35-72 semantics;
    the BEV overlap by clipping A against the four sides of B in WORLD coordinates -- an independent formulation)."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    oh = max(min(a[2] + a[5] / 2, b[2] + b[5] / 2) - max(a[2] - a[5] / 2, b[2] - b[5] / 2), 0.0)
    poly = list(_rect_corners(a))
    cb = _rect_corners(b)
    # orientation of cb: make it counter-clockwise so that "left of the edge" is inside
    area2 = sum(cb[i][0] * cb[(i + 1) % 4][1] - cb[(i + 1) % 4][0] * cb[i][1] for i in range(4))
    if area2 < 0:
        cb = cb[::-1]
    for i in range(4):
        poly = _clip(poly, cb[i], cb[(i + 1) % 4])
        if len(poly) < 3:
            return 0.0
    ov = 0.5 * abs(sum(poly[i][0] * poly[(i + 1) % len(poly)][1] - poly[(i + 1) % len(poly)][0] * poly[i][1] for i in range(len(poly))))
    o3 = ov * oh
    return o3 / max(a[3] * a[4] * a[5] + b[3] * b[4] * b[5] - o3, 1e-6)
