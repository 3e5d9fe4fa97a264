"""Generates tests/golden/iou_bev_ref.npz from the REAL reference: random pairs of [x y z l w h heading] boxes (the layout
det3d/ops/iou3d_nms/iou3d_nms_utils.py:35 takes), converted with the reference's own to_pcdet rule (:25-33, restated
below) and pushed through its CPU rotated-BEV-IoU, det3d/ops/iou3d_nms/src/iou3d_cpu.cpp:232, built by oracle/build_ref.py.

    python tests/golden/make_golden_iou.py        # needs the reference tree (this container), not the GPU box
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import build_ref  # noqa: E402


def to_pcdet(b):
    b = b[:, [0, 1, 2, 4, 3, 5, 6]].copy()
    b[:, 6] = -b[:, 6] - np.pi / 2
    return b


def main():
    ref = build_ref.load_iou3d_cpu()
    assert ref is not None, "reference tree not available"
    rng = np.random.default_rng(20260117)
    n = 96
    a = np.zeros((n, 7), np.float32)
    a[:, :2] = rng.uniform(-4, 4, (n, 2)); a[:, 2] = rng.uniform(-1, 1, n)
    a[:, 3] = rng.uniform(0.6, 6.0, n); a[:, 4] = rng.uniform(0.5, 2.5, n); a[:, 5] = rng.uniform(1.0, 2.5, n)
    a[:, 6] = rng.uniform(-np.pi, np.pi, n)
    b = a.copy()
    b[:, :2] += rng.normal(0, 0.8, (n, 2)).astype(np.float32); b[:, 2] += rng.normal(0, 0.3, n).astype(np.float32)
    b[:, 3:6] *= rng.uniform(0.8, 1.25, (n, 3)).astype(np.float32)
    b[:, 6] += rng.normal(0, 0.4, n).astype(np.float32)
    b[:8] = a[:8]                                   # identical boxes
    b[8:12, :2] += 50.0                             # far apart
    b[12:16, 6] = a[12:16, 6] + np.float32(np.pi / 2)   # crossed
    b[16:20, 3:5] = a[16:20, 3:5] * 0.4             # contained
    out = torch.zeros((n, n), dtype=torch.float32)
    ref.boxes_iou_bev_cpu(torch.from_numpy(to_pcdet(a)).contiguous(), torch.from_numpy(to_pcdet(b)).contiguous(), out)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "iou_bev_ref.npz")
    np.savez_compressed(path, boxes_a=a, boxes_b=b, iou_bev=out.numpy())
    print("wrote", path, "mean iou", float(out.mean()), "nonzero", int((out > 0).sum()))


if __name__ == "__main__":
    main()
