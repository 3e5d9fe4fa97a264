"""Pins the training-mode oracle (oracle/train.py) against the REAL reference modules in train() mode: loss, every
parameter gradient and the BatchNorm running statistics after one forward / backward (container only; CPU)."""
import numpy as np
import pytest
import torch

from helpers import synth
from oracle import refshim, train as otrain

pytestmark = pytest.mark.skipif(not refshim.available(), reason="reference tree not mounted")


def _labels(bs, n, seed):
    g = torch.Generator().manual_seed(seed)
    return ((torch.rand((bs, n), generator=g) < 0.3).float(), torch.randn((bs, 3), generator=g) * 0.3,
            torch.randint(0, 12, (bs,), generator=g), torch.randn((bs,), generator=g) * 0.1,
            torch.randint(0, 3, (bs,), generator=g), torch.randn((bs, 3), generator=g) * 0.2)


class _strided_rng:
    """np.random.choice / shuffle -> the deterministic 'strided' rule while the reference forward runs."""

    def __enter__(self):
        self.c, self.s = np.random.choice, np.random.shuffle
        np.random.choice = lambda L, n, replace=True: ((np.arange(n, dtype=np.int64) * L) // n if not replace
                                                       else np.arange(n, dtype=np.int64) % L)
        np.random.shuffle = lambda a: None

    def __exit__(self, *a):
        np.random.choice, np.random.shuffle = self.c, self.s


@pytest.mark.parametrize("kind", ["static_one", "static_two", "dynamic"])
def test_oracle_training_step_matches_reference_autograd(kind):
    rs, rd, _, _ = refshim.load()
    torch.set_num_threads(8)
    bs = 4
    sd = synth.random_state_dict(kind, seed=31)
    if kind == "dynamic":
        tr = synth.dynamic_tracks(bs, npoints=256, seed=5)
        pts = torch.from_numpy(tr["pts_pm"]).transpose(2, 1).contiguous()
        aux = torch.from_numpy(tr["box_sm"]).transpose(2, 1).contiguous()
    else:
        tr = synth.static_tracks(bs, n=1024, seed=5)
        pts = torch.from_numpy(tr["pts_pm"]).transpose(2, 1).contiguous()
        aux = torch.from_numpy(tr["init_box"])
    gt = torch.from_numpy(tr["bbox_gt"])
    labels = _labels(bs, pts.shape[2], 7)
    cls = {"static_one": rs.StaticModelOneBoxEst, "static_two": rs.StaticModelTwoBoxEst, "dynamic": rd.DynamicModel}[kind]
    crit = {"static_one": rs.FrustumPointNetLossOneBoxEst, "static_two": rs.FrustumPointNetLossTwoBoxEst,
            "dynamic": rd.DynamicModelLoss}[kind]()
    ref = cls().train()
    ref.load_state_dict(sd)
    ref.ins_seg.dropout.p = 0.0                      # dropout off on both sides (instance attribute; files untouched)
    with _strided_rng():
        out = ref(pts, aux, gt)
    ls = crit(out, *labels)
    ls["total_loss"].backward()
    if kind == "static_one":
        ols, _, grads, stats = otrain.static_one_step(sd, pts, aux, labels)
    elif kind == "static_two":
        ols, _, grads, stats = otrain.static_two_step(sd, pts, aux, gt, labels)
    else:
        ols, _, grads, stats = otrain.dynamic_step(sd, pts, aux, labels)
    assert abs(float(ols["total_loss"]) - float(ls["total_loss"])) <= 1e-5 * abs(float(ls["total_loss"]))
    for name, p in ref.named_parameters():
        g, r = grads[name], p.grad
        assert g is not None and r is not None, name
        denom = max(float(r.abs().max()), 1e-8)
        assert float((g - r).abs().max()) / denom < 2e-4, (name, float((g - r).abs().max()) / denom)
    for name, b in ref.named_buffers():
        if name.endswith("running_mean") or name.endswith("running_var"):
            assert torch.allclose(stats[name], b, rtol=1e-5, atol=1e-6), name
