"""CPU side of the "mixed" precision mode's parity claim: the numerics model of the kernels (helpers.emulate_seg_mixed: float64
on operands rounded exactly the way csrc/chain_split.cu rounds them -- the GPU tests pin the kernels to this model) against
the golden fixtures' fp32 reference logits.  BASELINE.json's bar is 1e-3 of max|ref|; a mask bit may differ only where the
reference margin is inside the logit error.  The BASELINE-size cases are in profiles/r2_precision_study_mixed.txt
(scripts/precision_study_mixed.py); here the small ones, in seconds."""
import numpy as np
import pytest
import torch

from helpers import bf16x2_round, emulate_seg_mixed, f16_round, fold_state_dict, load_model_case, rel_err, spec, synth


@pytest.mark.parametrize("name,bound", [("static_one", 1e-4), ("dynamic", 3e-4), ("static_one_default_init", 1e-4)])
def test_mixed_numerics_model_is_inside_the_bar_on_the_golden_cases(name, bound):
    z, sd, pts, aux, gt = load_model_case(name)
    fw = fold_state_dict(sd, "ins_seg", spec.seg_layers(pts.shape[1]))
    ref = z["strided/logits"]
    margin = ref[..., 1] - ref[..., 0]
    errs = {}
    for label, c5, d2 in (("bf16x3", False, 0), ("mixed", True, 2), ("mixed_d2_single", True, 1)):
        lg = emulate_seg_mixed(fw, pts, conv5_f16=c5, d2_mode=d2)[0].float().numpy()
        errs[label] = rel_err(lg, ref)
        flips = (lg[..., 0] < lg[..., 1]) != z["strided/mask"]
        # a differing bit sits inside the guard band of tests/test_gpu_parity.py (2 * tol * max|logit|) with room to spare
        assert np.all(np.abs(margin[flips]) <= 2 * 1e-3 * np.abs(ref).max()), (label, int(flips.sum()))
    assert errs["bf16x3"] < 2e-5, errs
    assert errs["mixed"] < bound, errs                      # measured 2.8e-5 / 1.1e-4 / 2.7e-5: >= 9x inside the 1e-3 bar
    assert errs["mixed_d2_single"] < 1e-3, errs
    assert errs["bf16x3"] < errs["mixed"] <= errs["mixed_d2_single"] * 1.05, errs


def test_fp16_last_layer_of_the_box_head_trunk_keeps_the_heads_inside_the_bar():
    """The mixed mode runs conv4 (256 -> 512) of the box-estimation trunk as f16(a) * f16(w): effect on the pooled feature
    and on the 39 head outputs (fc1-3 in fp32 / float64 here), three weight seeds."""
    for seed in (5, synth.REFERENCE_SEED, 3):
        sd = synth.random_state_dict("static_one", seed=seed)
        fw = fold_state_dict(sd, "box_est", spec.static_est_layers())
        torch.manual_seed(0)
        x = (torch.randn(16, 512, 3) * torch.tensor([1.5, 0.8, 0.6])).transpose(2, 1)
        h = x.transpose(2, 1)
        o_ref = torch.relu(h.double() @ fw["conv1"][0].double().t() + fw["conv1"][1].double())
        o = torch.relu(h @ fw["conv1"][0].t() + fw["conv1"][1])
        for nm in ("conv2", "conv3"):
            o_ref = torch.relu(o_ref @ fw[nm][0].double().t() + fw[nm][1].double())
            o = torch.relu(bf16x2_round(o).double() @ bf16x2_round(fw[nm][0]).double().t() + fw[nm][1]).float()
        g_ref = torch.relu((o_ref @ fw["conv4"][0].double().t()).max(dim=1)[0] + fw["conv4"][1].double())
        g = torch.relu((f16_round(o).double() @ f16_round(fw["conv4"][0]).double().t()).max(dim=1)[0] + fw["conv4"][1])
        assert rel_err(g, g_ref) < 6e-4, (seed, rel_err(g, g_ref))

        def heads(v):
            for nm in ("fc1", "fc2", "fc3"):
                v = v @ fw[nm][0].double().t() + fw[nm][1].double()
                if nm != "fc3":
                    v = torch.relu(v)
            return v
        assert rel_err(heads(g), heads(g_ref)) < 2e-4, (seed, rel_err(heads(g), heads(g_ref)))
