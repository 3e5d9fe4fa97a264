"""Loss forward: oracle vs the reference loss modules (CPU, when mounted); CUDA kernels vs the oracle (GPU)."""
import importlib

import numpy as np
import pytest
import torch

from oracle import losses as olosses, refshim

losses = importlib.import_module("3dal_pytorch_b200.losses")


def _case(bs, n, two, seed=0, device="cpu"):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)
    out = {"logits": r(bs, n, 2) * 2}
    sfxs = ("_one", "_two") if two else ("",)
    for s in sfxs:
        out.update({"center" + s: r(bs, 3), "heading_scores" + s: r(bs, 12), "heading_residuals_normalized" + s: r(bs, 12) * 0.3,
                    "heading_residuals" + s: r(bs, 12) * 0.1, "size_scores" + s: r(bs, 3),
                    "size_residuals_normalized" + s: r(bs, 3, 3) * 0.2, "size_residuals" + s: r(bs, 3, 3) * 0.5})
    if two:
        out["heading_class_label_two"] = torch.randint(0, 12, (bs,), generator=g)
        out["heading_residuals_label_two"] = r(bs) * 0.1
    labels = [(torch.rand(bs, n, generator=g) < 0.3).float(), r(bs, 3) * 2.5, torch.randint(0, 12, (bs,), generator=g), r(bs) * 0.1,
              torch.randint(0, 3, (bs,), generator=g), r(bs, 3) * 0.4]
    mv = lambda t: t.to(device)
    return {k: mv(v) for k, v in out.items()}, [mv(t) for t in labels]


@pytest.mark.skipif(not refshim.available(), reason="reference tree not mounted")
@pytest.mark.parametrize("two", [False, True])
def test_oracle_losses_equal_reference(two):
    sm, dm, _, _ = refshim.load()
    out, labels = _case(16, 300, two)
    ref = (sm.FrustumPointNetLossTwoBoxEst() if two else sm.FrustumPointNetLossOneBoxEst())(out, *labels, w_box=0.7)
    got = (olosses.two_box if two else olosses.one_box)(out, *labels, w_box=0.7)
    assert set(ref) == set(got)
    for k in ref:
        assert torch.allclose(ref[k], got[k], rtol=1e-6, atol=1e-7), k
    if not two:
        refd = dm.DynamicModelLoss()(out, *labels, w_box=0.7)
        for k in refd:
            assert torch.allclose(refd[k], got[k], rtol=1e-6, atol=1e-7), k


@pytest.mark.gpu
@pytest.mark.parametrize("two", [False, True])
@pytest.mark.parametrize("bs,n", [(1, 7), (64, 4096), (300, 1000)])
def test_cuda_losses_match_oracle(two, bs, n):
    out, labels = _case(bs, n, two, seed=bs)
    ref = (olosses.two_box if two else olosses.one_box)(out, *labels, w_box=0.5)
    dout, dlabels = _case(bs, n, two, seed=bs, device="cuda:0")
    mod = losses.FrustumPointNetLossTwoBoxEst() if two else losses.FrustumPointNetLossOneBoxEst()
    got = mod(dout, *dlabels, w_box=0.5)
    assert set(got) == set(ref)
    for k in ref:
        assert abs(float(got[k]) - float(ref[k])) <= 1e-4 * max(1.0, abs(float(ref[k]))), (k, float(got[k]), float(ref[k]))
    again = mod(dout, *dlabels, w_box=0.5)
    assert all(torch.equal(got[k], again[k]) for k in got)          # deterministic reduction order
