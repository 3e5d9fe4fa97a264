"""Oracle vs the committed golden vectors (generated from the real reference by
tests/golden/make_golden.py).  CPU only; this is what pins the oracle on the GPU box."""
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN, load_model_case, rel_err
from oracle import codecs, crop, gather, models

CASES = ["static_one", "static_two", "dynamic", "static_one_default_init",
         "static_one_cfg1", "static_two_cfg1", "dynamic_cfg2"]        # *_cfgN: the BASELINE.json config sizes


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("policy", ["numpy_legacy", "strided"])
def test_model_forward_matches_reference_outputs(name, policy):
    z, sd, pts, aux, gt = load_model_case(name)
    kind = str(z["kind"])
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    if policy == "numpy_legacy":
        np.random.seed(int(z["rng_seed"]))
    out = models.FORWARDS[kind](sd, pts, aux, gt, policy=policy)
    keys = [k.split("/", 1)[1] for k in z if k.startswith(policy + "/")]
    assert keys
    for k in keys:
        ref = z[policy + "/" + k]
        got = out[k].cpu().numpy()
        assert got.shape == ref.shape, k
        if ref.dtype == np.bool_ or np.issubdtype(ref.dtype, np.integer):
            assert np.array_equal(got, ref), k          # masks, class labels: bit-exact
        else:
            # same fp32 algorithm, possibly another BLAS kernel order on this host: 1e-5 of max|ref|
            assert rel_err(got, ref) < 1e-5, (k, rel_err(got, ref))


def test_mask_is_strict_less_than_and_nan_false():
    lg = torch.tensor([[[0.0, 1.0], [1.0, 1.0], [2.0, 1.0], [float("nan"), 1.0], [0.0, float("nan")]]])
    assert gather.mask_from_logits(lg).tolist() == [[True, False, False, False, False]]


def test_gather_policies_edge_cases():
    pts = np.arange(2 * 3 * 10, dtype=np.float32).reshape(2, 3, 10)
    mask = np.zeros((2, 10), dtype=bool)
    mask[1, [2, 5, 7]] = True
    obj, idx = gather.gather_object_pts(pts, mask, 8, "strided")
    assert np.all(obj[0] == 0) and np.all(idx[0] == 0)          # empty object stays zero
    assert idx[1].tolist() == [2, 5, 7, 2, 5, 7, 2, 5]          # cyclic repeat when L < n_pts
    mask[0, :] = True
    _, idx = gather.gather_object_pts(pts, mask, 4, "strided")
    assert idx[0].tolist() == [0, 2, 5, 7]                      # (j*L)//n_pts when L >= n_pts
    np.random.seed(3)
    _, idx = gather.gather_object_pts(pts, mask, 4, "numpy_legacy")
    assert len(set(idx[0].tolist())) == 4                       # without replacement when L >= n_pts
    assert set(idx[1].tolist()) == {2, 5, 7}                    # every fg point kept when L < n_pts


def test_crop_matches_reference_frame():
    z = np.load(os.path.join(GOLDEN, "crop_frame.npz"))
    pts, det, pose = z["points"], z["det_boxes"], z["pose"]
    box = crop.detector_to_waymo(det)
    assert np.array_equal(box, z["waymo_boxes"])
    idx, xyz = crop.crop_frame(pts, box, pose)
    off = z["offsets"]
    for b in range(box.shape[0]):
        assert np.array_equal(idx[b], z["indices"][off[b]:off[b + 1]]), b      # bit-exact, ascending
        assert np.array_equal(xyz[b], z["xyz_global"][off[b]:off[b + 1]]), b   # same f64 matmul
    lab = crop.points_in_boxes(z["points_f64"], box[:8])
    assert np.array_equal(lab, z["labels_f64"])


def test_crop_empty_inputs():
    box = np.array([[0, 0, 0, 2, 1, 1, 0.3]], dtype=np.float32)
    assert crop.points_in_boxes(np.zeros((0, 3), np.float32), box).shape == (0, 1)
    assert crop.points_in_boxes(np.zeros((5, 3), np.float32), np.zeros((0, 7), np.float32)).shape == (5, 0)
    # a point exactly on a face is outside (sign >= 0 rejects)
    on_face = np.array([[1.0, 0.0, 0.0]], dtype=np.float32)
    axis_box = np.array([[0, 0, 0, 2, 1, 1, 0.0]], dtype=np.float32)
    assert not crop.points_in_boxes(on_face, axis_box)[0, 0]
    assert crop.points_in_boxes(np.array([[0.999, 0, 0]], np.float32), axis_box)[0, 0]


def test_codecs_match_reference():
    z = np.load(os.path.join(GOLDEN, "codecs.npz"))
    for a, (c, r) in zip(z["angles"], z["a2c64"]):
        cc, rr = codecs.angle2class(float(a), 12)
        assert cc == int(c) and rr == r
    for a, (c, r) in zip(z["angles32"], z["a2c32"]):
        cc, rr = codecs.angle2class_f32(np.float32(a), 12)
        assert cc == int(c) and np.float32(rr) == np.float32(r), (a, cc, c, rr, r)
    for c, r, ang in zip(z["cls"], z["res"], z["c2a"]):
        assert codecs.class2angle(int(c), float(r), 12) == ang
    for x, c, r in zip(z["lwh"], z["s2c_cls"], z["s2c_res"]):
        cc, rr = codecs.size2class(x)
        assert cc == int(c) and np.array_equal(rr, r)
