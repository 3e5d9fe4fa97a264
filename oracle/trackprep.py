"""Dataset-side canonicalisation.  TEST INFRASTRUCTURE (oracle).

numpy float64 restatement of the geometric part of STATICTRACK.__getitem__ (tools/static_model.py:538-547,
569-570) and DYNAMICTRACK.__getitem__ (tools/dynamic_model.py:429-453,489,503-507), with the random resample
made explicit (the caller passes the drawn indices)."""
import numpy as np


def rotz(angle):
    c, s = np.cos(angle), np.sin(angle)
    return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])


def transform_box(box, pose):
    """tools/static_model.py:574-588."""
    heading = box[..., -1] + np.arctan2(pose[..., 1, 0], pose[..., 0, 0])
    center = np.einsum('...ij,...nj->...ni', pose[..., 0:3, 0:3], box[..., 0:3]) + np.expand_dims(pose[..., 0:3, 3], axis=-2)
    return np.concatenate([center, box[..., 3:6], heading[..., np.newaxis]], axis=-1)


def static_item(points_global, choice, veh_to_global, best_box_global):
    """points_global (N,3) f64 merged crops, choice (npoints,) indices, veh_to_global (4,4), best_box_global (7,)
    -> canonical points (npoints,3) f64, init box (1,7) f64 in the vehicle frame."""
    pose = np.linalg.inv(np.reshape(veh_to_global, [4, 4]))
    bbox = transform_box(best_box_global[np.newaxis, ...], pose)
    p = points_global.T
    p = pose @ np.concatenate([p, np.ones((1, p.shape[1]))], axis=0)
    p = p[:3, :].T
    p = p[choice, :]
    p = p - bbox[:, :3]
    p = (rotz(-bbox[0, -1]) @ p.T).T
    return p, bbox


def dynamic_item(frame_points, frame_choice, boxes_global, veh_to_global, npoints=1024, r=2, s=50):
    """frame_points: list of 5 arrays (k,3) f64 or None (missing / empty frame); frame_choice: list of 5 index arrays;
    boxes_global (101,8) f64 with zero rows outside the track and dt in column 7.
    -> point (5*npoints,4) f64, bbox (101,8) f64 (relative), init_box (8,) f64."""
    point = np.zeros((0, 4))
    for j in range(2 * r + 1):
        t = np.full((npoints, 1), 0.1 * (j - r))
        if frame_points[j] is None or len(frame_points[j]) == 0:
            point = np.vstack([point, np.hstack([np.zeros((npoints, 3)), t])])
        else:
            point = np.vstack([point, np.hstack([np.copy(frame_points[j][frame_choice[j]]), t])])
    bbox = np.copy(boxes_global)
    pose = np.linalg.inv(np.reshape(veh_to_global, [4, 4]))
    bbox[:, :7] = transform_box(bbox[:, :7], pose)
    point[:, :3] = (pose @ np.concatenate([point[:, :3].T, np.ones((1, point.shape[0]))], axis=0)).T[:, :3]
    init_box = np.copy(bbox[s])
    point[:, :3] = point[:, :3] - bbox[s, :3]
    point[:, :3] = (rotz(-bbox[s, -2]) @ point[:, :3].T).T
    bbox[:, :3] = bbox[:, :3] - bbox[s, :3]
    bbox[:, -2] = bbox[:, -2] - bbox[s, -2]
    return point, bbox, init_box
