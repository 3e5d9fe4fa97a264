"""Generates tests/golden/*.npz from the REAL reference (build container only).

    python tests/golden/make_golden.py

Needs /root/reference (imported through oracle/refshim.py, CPU).  The fixtures hold inputs, the
seed of the weights (+ a checksum of the generated tensors) and the reference's outputs; the
weights themselves are regenerated from the seed by 3dal_pytorch_b200.synth.random_state_dict
plus the recorded dconv5 calibration, so the files stay small.
"""
import hashlib
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refshim, models  # noqa: E402

synth = importlib.import_module("3dal_pytorch_b200.synth")
OUT = os.path.dirname(os.path.abspath(__file__))


def sd_checksum(sd):
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


class patched_numpy_rng:
    """Replace np.random.choice / shuffle by the deterministic 'strided' rule (oracle/gather.py)
    while the reference forward runs.  The reference files are untouched."""

    def __enter__(self):
        self.c, self.s = np.random.choice, np.random.shuffle

        def choice(L, n, replace=True):
            j = np.arange(n, dtype=np.int64)
            return (j * L) // n if not replace else j % L

        np.random.choice = choice
        np.random.shuffle = lambda a: None
        return self

    def __exit__(self, *a):
        np.random.choice, np.random.shuffle = self.c, self.s


def arrays_checksum(*arrs):
    h = hashlib.sha256()
    for a in arrs:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def model_case(name, kind, bs, n, wseed, dseed, fg_fraction, calibrate=True, rng_seed=424242, compact=False):
    """compact=True (the BASELINE.json config sizes): the inputs are not stored but regenerated from `dseed` by
    3dal_pytorch_b200.synth (their sha256 is), logits / mask are stored once (they do not depend on the gather
    policy) and the mask as packed bits."""
    sm, dm, _, _ = refshim.load()
    sd = synth.random_state_dict(kind, seed=wseed, randomize_bn=calibrate)
    if kind == "dynamic":
        tr = synth.dynamic_tracks(bs, npoints=n // 5, seed=dseed)
        pts_pm, aux = tr["pts_pm"], tr["box_sm"]
    else:
        tr = synth.static_tracks(bs, n=n, seed=dseed)
        pts_pm, aux = tr["pts_pm"], tr["init_box"]
    pts = torch.from_numpy(pts_pm).transpose(2, 1)            # the strided (bs,C,n) view the eval scripts pass
    aux_t = torch.from_numpy(aux).transpose(2, 1) if kind == "dynamic" else torch.from_numpy(aux)
    gt = torch.from_numpy(tr["bbox_gt"])
    cls = {"static_one": sm.StaticModelOneBoxEst, "static_two": sm.StaticModelTwoBoxEst, "dynamic": dm.DynamicModel}[kind]
    ref = cls().eval()
    q = std = 0.0
    if calibrate:
        ref.load_state_dict(sd)
        with torch.no_grad():
            logits = ref.ins_seg(pts)
        q, std = synth.calibrate_seg_margin(sd, logits, fg_fraction)
    ref.load_state_dict(sd)
    rec = {"kind": kind, "wseed": wseed, "calibrated": int(calibrate), "calib_q": q, "calib_std": std,
           "sd_sha256": sd_checksum(sd), "rng_seed": rng_seed}
    if compact:
        rec.update({"compact": 1, "bs": bs, "n": n, "dseed": dseed,
                    "inputs_sha256": arrays_checksum(pts_pm, aux, tr["bbox_gt"])})
    else:
        rec.update({"pts_pm": pts_pm, "aux": aux, "bbox_gt": tr["bbox_gt"]})
    for policy in ("numpy_legacy", "strided"):
        with torch.no_grad():
            if policy == "numpy_legacy":
                np.random.seed(rng_seed)
                out = ref(pts, aux_t, gt)
            else:
                with patched_numpy_rng():
                    out = ref(pts, aux_t, gt)
        for k, v in out.items():
            v = v.detach().cpu().numpy()
            if compact and k in ("logits", "mask"):
                key = "common/" + k
                v = np.packbits(v, axis=1) if k == "mask" else v
                if key in rec:
                    assert np.array_equal(rec[key], v), "logits / mask depend on the gather policy?"
                rec[key] = v
            else:
                rec["%s/%s" % (policy, k)] = v
    counts = (np.unpackbits(rec["common/mask"], axis=1)[:, :n] if compact else rec["strided/mask"]).sum(1)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
    print(name, "fg counts", counts.tolist(), "bytes", os.path.getsize(os.path.join(OUT, name + ".npz")))


def crop_case():
    _, _, bno, _ = refshim.load()
    fr = synth.lidar_frames(1, n_points=24000, n_boxes=32, seed=11)[0]
    pts, det, pose = fr["points"], fr["det_boxes"], fr["pose"]
    box = np.array(det, copy=True)                              # waymo_common.py:110-111
    box[:, -1] = -box[:, -1] - np.pi / 2
    box = box[:, [0, 1, 2, 4, 3, 5, -1]]
    idx, xyz, off = [], [], [0]
    for i in range(box.shape[0]):
        ind = bno.points_in_rbbox(pts, box[i][np.newaxis, ...])          # waymo_common.py:168
        o = pts[ind.reshape([-1])].T                                     # :169
        o = pose @ np.concatenate([o, np.ones((1, o.shape[1]))], axis=0)  # :170
        idx.append(np.nonzero(ind.reshape(-1))[0])
        xyz.append(o[:3, :].T)
        off.append(off[-1] + len(idx[-1]))
    # also the float64-points label path (tools/static_model.py:556)
    p64 = pts[:4000].astype(np.float64) * 1.0000001
    lab = np.stack([bno.points_in_rbbox(p64, box[i][None]).reshape(-1) for i in range(8)], 1)
    np.savez_compressed(os.path.join(OUT, "crop_frame.npz"), points=pts, det_boxes=det, waymo_boxes=box, pose=pose,
                        indices=np.concatenate(idx), xyz_global=np.concatenate(xyz), offsets=np.asarray(off),
                        points_f64=p64, labels_f64=lab)
    print("crop_frame", "inside per box", np.diff(off).tolist())


def codec_case():
    _, _, _, ut = refshim.load()
    rng = np.random.default_rng(5)
    ang = np.concatenate([rng.uniform(-12, 12, 400), np.arange(-24, 25) * (np.pi / 12), [0.0, np.pi, -np.pi, 2 * np.pi]])
    a2c64 = np.array([ut.angle2class(a, 12) for a in ang])
    a32 = ang.astype(np.float32)
    a2c32 = np.array([[float(v) for v in ut.angle2class(torch.tensor(a), 12)] for a in a32])
    cls = rng.integers(0, 12, len(ang))
    res = rng.normal(0, 0.3, len(ang))
    c2a = np.array([ut.class2angle(int(c), r, 12) for c, r in zip(cls, res)])
    lwh = rng.uniform(0.3, 12, (200, 3))
    s2c = [ut.size2class(x) for x in lwh]
    np.savez_compressed(os.path.join(OUT, "codecs.npz"), angles=ang, a2c64=a2c64, angles32=a32, a2c32=a2c32,
                        cls=cls, res=res, c2a=c2a, lwh=lwh, s2c_cls=np.array([c for c, _ in s2c]),
                        s2c_res=np.stack([r for _, r in s2c]))
    print("codecs ok")


if __name__ == "__main__":
    torch.set_num_threads(8)
    model_case("static_one", "static_one", bs=6, n=2048, wseed=101, dseed=1, fg_fraction=0.25)
    model_case("static_two", "static_two", bs=6, n=2048, wseed=102, dseed=2, fg_fraction=0.25)
    model_case("dynamic", "dynamic", bs=3, n=5120, wseed=103, dseed=3, fg_fraction=0.5)
    model_case("static_one_default_init", "static_one", bs=2, n=512, wseed=104, dseed=4, fg_fraction=0.0, calibrate=False)
    # BASELINE.json config 1 (static 32 x 4096) and config 2 (dynamic 64 x (5 x 1024 points + 101 boxes))
    model_case("static_one_cfg1", "static_one", bs=32, n=4096, wseed=201, dseed=21, fg_fraction=0.3, compact=True)
    model_case("static_two_cfg1", "static_two", bs=32, n=4096, wseed=202, dseed=22, fg_fraction=0.3, compact=True)
    model_case("dynamic_cfg2", "dynamic", bs=64, n=5120, wseed=203, dseed=23, fg_fraction=0.4, compact=True)
    crop_case()
    codec_case()
