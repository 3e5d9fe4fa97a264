"""Live oracle-vs-reference pinning.  Runs only where /root/reference exists (build container);
skipped on the GPU box, where tests/test_oracle_golden.py carries the pin."""
import numpy as np
import pytest
import torch

from helpers import synth
from oracle import crop, models, refshim

pytestmark = pytest.mark.skipif(not refshim.available(), reason="reference tree not mounted")


@pytest.mark.parametrize("kind,bs,n", [("static_one", 3, 768), ("static_two", 3, 768), ("dynamic", 2, 1280)])
def test_forward_bitwise_equal_to_reference(kind, bs, n):
    sm, dm, _, _ = refshim.load()
    sd = synth.random_state_dict(kind, seed=11)
    if kind == "dynamic":
        tr = synth.dynamic_tracks(bs, npoints=n // 5, seed=5)
        pts = torch.from_numpy(tr["pts_pm"]).transpose(2, 1)
        aux = torch.from_numpy(tr["box_sm"]).transpose(2, 1)
    else:
        tr = synth.static_tracks(bs, n=n, seed=5)
        pts = torch.from_numpy(tr["pts_pm"]).transpose(2, 1)
        aux = torch.from_numpy(tr["init_box"])
    gt = torch.from_numpy(tr["bbox_gt"])
    logits, _ = models.seg_forward(sd, pts)
    synth.calibrate_seg_margin(sd, logits, 0.4)
    ref = {"static_one": sm.StaticModelOneBoxEst, "static_two": sm.StaticModelTwoBoxEst,
           "dynamic": dm.DynamicModel}[kind]().eval()
    ref.load_state_dict(sd)
    np.random.seed(9)
    with torch.no_grad():
        r = ref(pts, aux, gt)
    np.random.seed(9)
    o = models.FORWARDS[kind](sd, pts, aux, gt, policy="numpy_legacy")
    assert set(r) == {k for k in o if not k.startswith("_")}
    for k, v in r.items():
        assert v.dtype == o[k].dtype and v.shape == o[k].shape, k
        assert torch.equal(v, o[k]), k


def test_points_in_rbbox_bitwise_equal_to_reference():
    _, _, bno, _ = refshim.load()
    rng = np.random.default_rng(2)
    pts = (rng.uniform(-1, 1, (5000, 3)) * np.array([6, 6, 3])).astype(np.float32)
    boxes = np.concatenate([rng.normal(0, 1.5, (64, 3)), rng.uniform(0.5, 6, (64, 3)),
                            rng.uniform(-7, 7, (64, 1))], 1).astype(np.float32)
    mine = crop.points_in_boxes(pts, boxes)
    for b in range(boxes.shape[0]):
        assert np.array_equal(bno.points_in_rbbox(pts, boxes[b][None]).reshape(-1), mine[:, b])
