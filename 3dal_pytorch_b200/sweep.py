"""End-to-end auto-label sweep on device (BASELINE.json configs[3]): per-frame points-in-box crop -> regroup the
crops by track -> merge / resample / canonicalise (track prep) -> segmentation -> foreground gather -> box head ->
decoded 7-DoF boxes.  Mirrors the reference chain _create_pd_detection (det3d/datasets/waymo/waymo_common.py:139-171)
-> tools/trackData.py:25-45 -> STATICTRACK.__getitem__ (tools/static_model.py:529-572) -> test_one_epoch
(tools/static_eval.py:256-290).  Box b of every frame is taken to be track b (persistent tracking ids)."""
import numpy as np
import torch

from . import crop, trackprep, spec
from .pipeline import StaticAutoLabeler


def track_segments(offsets, box_off, n_frames, n_tracks):
    """CSR of the per-(frame, box) crops -> per-track segment table: seg_start (T,F) i64 rows into the crop output,
    seg_len (T,F) i64.  offsets: (sum B_f + 1,) i64 CUDA; every frame must hold n_tracks boxes."""
    assert all(int(box_off[f + 1] - box_off[f]) == n_tracks for f in range(n_frames))
    off = offsets[:-1].view(n_frames, n_tracks)
    nxt = offsets[1:].view(n_frames, n_tracks)
    return off.t().contiguous(), (nxt - off).t().contiguous()


def resample_rows(seg_start, seg_len, npoints, policy="strided"):
    """(T, npoints) i64 rows of the crop output: a resample with replacement of each track's merged crop
    (tools/static_model.py:532,546).  'strided' is the deterministic device rule (j*N)//npoints; 'numpy_legacy' replays
    np.random.choice on the host.  Tracks without points get -1 (zero points)."""
    T, F = seg_len.shape
    total = seg_len.sum(1)                                            # (T,)
    cum = torch.cumsum(seg_len, 1)                                    # inclusive
    if policy == "numpy_legacy":
        tot = total.cpu().numpy()
        k = np.stack([np.random.choice(int(n), npoints, replace=True) if n > 0 else np.zeros(npoints, np.int64) for n in tot])
        k = torch.from_numpy(k).to(seg_len.device)
    else:
        j = torch.arange(npoints, device=seg_len.device, dtype=torch.int64)[None, :]
        k = (j * total[:, None]) // npoints
    f = torch.searchsorted(cum, k, right=True).clamp_(max=F - 1)      # frame segment of the k-th merged point
    before = cum.gather(1, f) - seg_len.gather(1, f)
    rows = seg_start.gather(1, f) + (k - before)
    return torch.where(total[:, None] > 0, rows, torch.full_like(rows, -1))


class StaticSweep:
    def __init__(self, model, npoints=spec.NUM_POINT_STATIC, policy="strided"):
        self.labeler = StaticAutoLabeler(model)
        self.npoints, self.policy = npoints, policy

    def run(self, frames, det_scores=None):
        """frames: list of dict(points (N,3) f32, det_boxes (B,7) f32 CenterPoint convention, pose (4,4) f64).
        Returns dict(boxes (T,7) f32 refined boxes in the vehicle frame of each track's best frame, init_box (T,7) f64,
        crop result, pts (T,npoints,3))."""
        dev = next(self.labeler.model.parameters()).device
        F = len(frames)
        # one conversion for the whole sweep (every frame holds the same number of boxes: persistent tracking ids)
        T = int(np.asarray(frames[0]["det_boxes"]).shape[0])
        all_w = crop.detector_to_waymo(np.concatenate([np.asarray(f["det_boxes"]).reshape(-1, 7) for f in frames], 0)).reshape(F, -1, 7)
        poses = np.stack([np.asarray(f["pose"], dtype=np.float64) for f in frames])
        plan = crop.CropPlan([f["points"] for f in frames], all_w, poses, device=dev)
        res = plan.run()
        seg_start, seg_len = track_segments(res["offsets"], res["box_off"], F, T)
        rows = resample_rows(seg_start, seg_len, self.npoints, self.policy)
        # best-score frame per track (tools/static_model.py:535): scores (F,T); default = frame with most points
        if det_scores is None:
            best = seg_len.argmax(1).cpu().numpy()
        else:
            best = np.asarray(det_scores).argmax(0)
        inv_pose = np.linalg.inv(poses[best])                                       # (T,4,4) global -> vehicle frame
        # the detector box of the best frame, in that frame's vehicle coordinates, is the initial box
        init_box = all_w[best, np.arange(T)].astype(np.float64)
        d_inv = torch.from_numpy(np.ascontiguousarray(inv_pose)).to(dev)
        d_init = torch.from_numpy(np.ascontiguousarray(init_box)).to(dev)
        pts = trackprep.prep_points(res["xyz_global"].contiguous(), rows.contiguous(), d_inv, d_init, heading_col=6, c_out=3)
        boxes = self.labeler.label_device(pts.transpose(2, 1), d_init.float())
        return {"boxes": boxes, "init_box": d_init, "pts": pts, "rows": rows, "best": best, "crop": res}
