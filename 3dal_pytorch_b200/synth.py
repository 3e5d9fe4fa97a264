"""Synthetic tracks and random-init weights for the benchmark and the parity tests.

There is no dataset and no checkpoint on the box (no network), so both sides of every comparison
are fed by this generator:

  * ``random_state_dict``  reference-format ``state_dict`` drawn from a seeded CPU generator:
    conv / linear weights and biases U(-1/sqrt(fan_in), 1/sqrt(fan_in)) (the nn.Conv1d / nn.Linear
    default), BatchNorm statistics randomised (weight U(.5,1.5), bias N(0,.1), running_mean N(0,.1),
    running_var U(.5,1.5)) so that BN folding is actually exercised.
  * ``calibrate_seg_margin``  rescales ``ins_seg.dconv5`` so the logit margin l1-l0 has unit spread
    and a chosen foreground fraction (default-init weights give an all-0 / all-1 mask).
  * ``static_tracks`` / ``dynamic_tracks``  inputs shaped like STATICTRACK / DYNAMICTRACK items
    (tools/static_model.py:529-572, tools/dynamic_model.py:419-509).
  * ``lidar_frames``  Waymo-shaped frames + detector boxes for the crop (waymo_common.py:100-111).
"""
import math

import numpy as np
import torch

from . import spec

REFERENCE_SEED = 10922081   # tools/static_eval.py:303 fixSeed(10922081)


def random_state_dict(kind, seed=REFERENCE_SEED, n_channel=None, randomize_bn=True):
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    sd = {}

    def uni(shape, lo, hi):
        return torch.rand(shape, generator=g, dtype=torch.float32) * (hi - lo) + lo

    def nrm(shape, std):
        return torch.randn(shape, generator=g, dtype=torch.float32) * std

    for block, table in spec.model_blocks(kind, n_channel):
        for lname, bn, cin, cout, lkind in table:
            bound = 1.0 / math.sqrt(cin)
            wshape = (cout, cin, 1) if lkind == "conv" else (cout, cin)
            sd["%s.%s.weight" % (block, lname)] = uni(wshape, -bound, bound)
            sd["%s.%s.bias" % (block, lname)] = uni((cout,), -bound, bound)
            if bn is not None:
                p = "%s.%s." % (block, bn)
                if randomize_bn:
                    sd[p + "weight"] = uni((cout,), 0.5, 1.5)
                    sd[p + "bias"] = nrm((cout,), 0.1)
                    sd[p + "running_mean"] = nrm((cout,), 0.1)
                    sd[p + "running_var"] = uni((cout,), 0.5, 1.5)
                else:
                    sd[p + "weight"] = torch.ones(cout)
                    sd[p + "bias"] = torch.zeros(cout)
                    sd[p + "running_mean"] = torch.zeros(cout)
                    sd[p + "running_var"] = torch.ones(cout)
                sd[p + "num_batches_tracked"] = torch.zeros((), dtype=torch.int64)
    return sd


def calibrate_seg_margin(sd, logits, fg_fraction=0.125, prefix="ins_seg"):
    """Given logits (bs,n,2) produced WITH ``sd``, rewrite ``dconv5`` in place so that the new
    margin is (old margin - q) / std, q being the (1 - fg_fraction) quantile: unit spread and the
    requested share of foreground points.  Returns (q, std)."""
    m = (logits[..., 1] - logits[..., 0]).detach().float().cpu().flatten()
    if m.numel() > 1 << 20:
        m = m[torch.randperm(m.numel(), generator=torch.Generator().manual_seed(0))[: 1 << 20]]
    std = float(m.std()) or 1.0
    q = float(torch.quantile(m, 1.0 - fg_fraction))
    w = sd[prefix + ".dconv5.weight"]
    b = sd[prefix + ".dconv5.bias"]
    w.mul_(1.0 / std)
    b.mul_(1.0 / std)
    b[1] -= q / std
    return q, std


def _anchor_boxes(rng, bs):
    cls = rng.integers(0, 3, size=bs)
    size = np.asarray(spec.MEAN_SIZE_ARR)[cls] + rng.normal(0, 0.1, size=(bs, 3))
    return cls, size


def static_tracks(bs, n=spec.NUM_POINT_STATIC, seed=0, dtype=np.float32):
    """Returns dict(pts_pm (bs,n,3) point-major f32, init_box (bs,7), bbox_gt (bs,7), n_fg (bs,)).
    Per track: an anchor-sized box at the origin with n_fg ~ U{64..3000} points inside it, the
    rest clutter in +-(8,8,2) m and a ground plane; shuffled."""
    rng = np.random.default_rng(seed)
    _, size = _anchor_boxes(rng, bs)
    pts = np.empty((bs, n, 3), dtype=np.float64)
    n_fg = rng.integers(64, min(3000, n - 1) + 1, size=bs) if n > 128 else rng.integers(1, n, size=bs)
    for i in range(bs):
        k = int(n_fg[i])
        fg = (rng.random((k, 3)) - 0.5) * size[i]
        rest = n - k
        g = rest // 3
        ground = np.stack([rng.uniform(-8, 8, g), rng.uniform(-8, 8, g), rng.normal(-size[i, 2] / 2, 0.03, g)], 1)
        clutter = rng.uniform(-1, 1, (rest - g, 3)) * np.array([8.0, 8.0, 2.0])
        p = np.concatenate([fg, ground, clutter], 0)
        pts[i] = p[rng.permutation(n)]
    init_box = np.concatenate([rng.normal(0, 0.3, (bs, 3)), size, rng.normal(0, 0.1, (bs, 1))], 1)
    bbox_gt = init_box + rng.normal(0, 0.05, (bs, 7))
    return {"pts_pm": pts.astype(dtype), "init_box": init_box.astype(dtype), "bbox_gt": bbox_gt.astype(dtype),
            "n_fg": n_fg}


def dynamic_tracks(bs, npoints=spec.NUM_POINT_DYNAMIC, seed=0, dtype=np.float32):
    """Returns dict(pts_pm (bs,5*npoints,4), box_sm (bs,101,8) step-major, bbox_gt (bs,7)).
    Channel 3 of pts is 0.1*(frame-2); channel 7 of box is 0.1*(step-50); frames / steps outside
    the (random-length) track are zero-padded like tools/dynamic_model.py:431-447."""
    rng = np.random.default_rng(seed)
    F, S = spec.NUM_FRAME, spec.NUM_BOX_STEPS
    _, size = _anchor_boxes(rng, bs)
    pts = np.zeros((bs, F * npoints, 4), dtype=np.float64)
    box = np.zeros((bs, S, 8), dtype=np.float64)
    for i in range(bs):
        lo = rng.integers(0, 45)            # first valid step
        hi = rng.integers(56, S + 1)        # one past last valid step
        vel = rng.normal(0, 0.6, 2)
        steps = np.arange(S) - 50
        walk = np.cumsum(rng.normal(0, 0.02, (S, 2)), 0)
        walk -= walk[50]
        ctr = np.concatenate([steps[:, None] * 0.1 * vel[None, :] + walk, rng.normal(0, 0.02, (S, 1))], 1)
        hd = np.cumsum(rng.normal(0, 0.01, S))
        hd -= hd[50]
        b = np.concatenate([ctr, np.tile(size[i], (S, 1)) + rng.normal(0, 0.03, (S, 3)), hd[:, None]], 1)
        valid = (np.arange(S) >= lo) & (np.arange(S) < hi)
        box[i, :, :7] = np.where(valid[:, None], b, 0.0)
        box[i, :, 7] = 0.1 * steps
        for j in range(F):
            sl = slice(j * npoints, (j + 1) * npoints)
            pts[i, sl, 3] = 0.1 * (j - 2)
            step = 50 + (j - 2)
            if not valid[step] or rng.random() < 0.08:
                continue                      # zero xyz block (track edge / empty crop)
            k = int(rng.integers(16, npoints))
            fg = (rng.random((k, 3)) - 0.5) * size[i] + ctr[step]
            cl = rng.uniform(-1, 1, (npoints - k, 3)) * np.array([6.0, 6.0, 1.5]) + ctr[step]
            p = np.concatenate([fg, cl], 0)
            pts[i, sl, :3] = p[rng.permutation(npoints)]
    bbox_gt = np.concatenate([rng.normal(0, 0.2, (bs, 3)), size, rng.normal(0, 0.1, (bs, 1))], 1)
    return {"pts_pm": pts.astype(dtype), "box_sm": box.astype(dtype), "bbox_gt": bbox_gt.astype(dtype)}


def lidar_frames(n_frames, n_points=180000, n_boxes=200, seed=0):
    """Waymo-shaped frames: points (N,3) f32 in +-75 m x +-75 m x [-2,4] m (ground-heavy, with
    point clusters on the objects) and detector boxes (B,7) f32 in the CenterPoint convention
    [x,y,z,w,l,h,r2] that waymo_common.py:110-111 converts (l<->w swap, heading -r2 - pi/2).
    Also returns per-frame 4x4 float64 vehicle->global poses."""
    rng = np.random.default_rng(seed)
    frames = []
    for f in range(n_frames):
        cls = rng.integers(0, 3, n_boxes)
        lwh = np.asarray(spec.MEAN_SIZE_ARR)[cls] * rng.uniform(0.8, 1.2, (n_boxes, 3))
        ctr = np.stack([rng.uniform(-70, 70, n_boxes), rng.uniform(-70, 70, n_boxes), rng.normal(0.5, 0.3, n_boxes)], 1)
        r2 = rng.uniform(-np.pi, np.pi, n_boxes)
        det = np.concatenate([ctr, lwh[:, [1, 0, 2]], r2[:, None]], 1).astype(np.float32)
        n_obj = n_points // 20                               # ~5 % of the points sit on objects
        which = rng.integers(0, n_boxes, n_obj)
        loc = (rng.random((n_obj, 3)) - 0.5) * lwh[which] * 1.1
        h = -r2[which] - np.pi / 2
        c, s = np.cos(h), np.sin(h)
        obj = np.stack([c * loc[:, 0] - s * loc[:, 1], s * loc[:, 0] + c * loc[:, 1], loc[:, 2]], 1) + ctr[which]
        n_gr = (n_points - n_obj) * 2 // 3
        ground = np.stack([rng.uniform(-75, 75, n_gr), rng.uniform(-75, 75, n_gr), rng.normal(-0.2, 0.05, n_gr)], 1)
        n_cl = n_points - n_obj - n_gr
        clutter = np.stack([rng.uniform(-75, 75, n_cl), rng.uniform(-75, 75, n_cl), rng.uniform(-2, 4, n_cl)], 1)
        pts = np.concatenate([obj, ground, clutter], 0)[rng.permutation(n_points)].astype(np.float32)
        yaw = rng.uniform(-np.pi, np.pi)
        pose = np.eye(4)
        pose[:2, :2] = [[np.cos(yaw), -np.sin(yaw)], [np.sin(yaw), np.cos(yaw)]]
        pose[:3, 3] = [rng.uniform(-500, 500), rng.uniform(-500, 500), rng.uniform(-5, 5)]
        frames.append({"points": pts, "det_boxes": det, "pose": pose})
    return frames


def static_tracks_device(bs, n=spec.NUM_POINT_STATIC, seed=0, device="cuda"):
    """Device-side twin of ``static_tracks`` for large batches (the host loop is too slow for 8192
    tracks): same recipe -- anchor-sized box at the origin with a random share of the points inside
    it, a ground plane and clutter, shuffled -- generated with torch ops.  Returns point-major
    pts_pm (bs,n,3) f32, init_box (bs,7), bbox_gt (bs,7)."""
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    r = lambda *s: torch.rand(*s, generator=g, device=device)
    rn = lambda *s: torch.randn(*s, generator=g, device=device)
    anchors = torch.tensor(spec.MEAN_SIZE_ARR, device=device, dtype=torch.float32)
    cls = torch.randint(0, 3, (bs,), generator=g, device=device)
    size = anchors[cls] + 0.1 * rn(bs, 3)
    n_fg = torch.randint(64, min(3000, n - 1) + 1, (bs, 1), generator=g, device=device)
    slot = torch.arange(n, device=device)[None, :]
    kind = torch.where(slot < n_fg, 0, torch.where(r(bs, n) < 1.0 / 3.0, 1, 2))     # 0 fg, 1 ground, 2 clutter
    u = r(bs, n, 3)
    fg = (u - 0.5) * size[:, None, :]
    ground = torch.stack([u[..., 0] * 16 - 8, u[..., 1] * 16 - 8,
                          -size[:, None, 2].expand(bs, n) / 2 + 0.03 * rn(bs, n)], -1)
    clutter = (u * 2 - 1) * torch.tensor([8.0, 8.0, 2.0], device=device)
    pts = torch.where((kind == 0)[..., None], fg, torch.where((kind == 1)[..., None], ground, clutter))
    perm = torch.argsort(r(bs, n), dim=1)
    pts = torch.gather(pts, 1, perm[..., None].expand(bs, n, 3)).contiguous()
    init_box = torch.cat([0.3 * rn(bs, 3), size, 0.1 * rn(bs, 1)], 1)
    bbox_gt = init_box + 0.05 * rn(bs, 7)
    return {"pts_pm": pts, "init_box": init_box, "bbox_gt": bbox_gt}
