"""Loss forward.  TEST INFRASTRUCTURE (oracle).  torch restatement of FrustumPointNetLossOneBoxEst /
...TwoBoxEst (tools/static_model.py:341-517) and DynamicModelLoss (tools/dynamic_model.py:314-398)."""
import numpy as np
import torch
import torch.nn.functional as F

from .codecs import MEAN_SIZE_ARR, NUM_HEADING_BIN


def huber(err, delta):
    a = err.abs()
    q = a.clamp(max=delta)
    return (0.5 * q ** 2 + delta * (a - q)).mean()


def head_terms(center, hs, hrn, ss, srn, center_label, hcls, hres, scls, sres):
    c = huber((center - center_label).norm(dim=1), 2.0)
    h = F.nll_loss(F.log_softmax(hs, dim=1), hcls.long())
    lab = hres / (np.pi / NUM_HEADING_BIN)
    hr = huber(hrn.gather(1, hcls.long()[:, None])[:, 0] - lab, 1.0)
    s = F.nll_loss(F.log_softmax(ss, dim=1), scls.long())
    anchors = torch.from_numpy(MEAN_SIZE_ARR).to(center.dtype).to(center.device)[scls.long()]
    pred = srn[torch.arange(srn.shape[0]), scls.long()]
    sr = huber((sres / anchors - pred).norm(dim=1), 1.0)
    return c, h, s, hr, sr


def mask_term(logits, mask_label):
    return F.nll_loss(F.log_softmax(logits.reshape(-1, 2), dim=1), mask_label.reshape(-1).long())


def one_box(output, mask_label, center_label, hcls, hres, scls, sres, w_box=1.0, suffix=""):
    m = mask_term(output["logits"], mask_label)
    c, h, s, hr, sr = head_terms(output["center" + suffix], output["heading_scores" + suffix],
                                 output["heading_residuals_normalized" + suffix], output["size_scores" + suffix],
                                 output["size_residuals_normalized" + suffix], center_label, hcls, hres, scls, sres)
    return {"total_loss": m + w_box * (c * 10 + h + s + hr * 20 + sr * 20), "mask_loss": m, "center_loss": w_box * c * 10,
            "heading_class_loss": w_box * h, "size_class_loss": w_box * s, "heading_residuals_normalized_loss": w_box * hr * 20,
            "size_residuals_normalized_loss": w_box * sr * 20}


def two_box(output, mask_label, center_label, hcls, hres, scls, sres, w_box=1.0):
    m = mask_term(output["logits"], mask_label)
    a = head_terms(output["center_one"], output["heading_scores_one"], output["heading_residuals_normalized_one"],
                   output["size_scores_one"], output["size_residuals_normalized_one"], center_label, hcls, hres, scls, sres)
    b = head_terms(output["center_two"], output["heading_scores_two"], output["heading_residuals_normalized_two"],
                   output["size_scores_two"], output["size_residuals_normalized_two"], center_label,
                   output["heading_class_label_two"], output["heading_residuals_label_two"], scls, sres)
    total = m + w_box * sum(x * k for t in (a, b) for x, k in zip(t, (10, 1, 1, 20, 20)))
    out = {"total_loss": total, "mask_loss": m}
    for t, sfx in ((a, "_one"), (b, "_two")):
        for x, k, name in zip(t, (10, 1, 1, 20, 20), ("center_loss", "heading_class_loss", "size_class_loss",
                                                       "heading_residuals_normalized_loss", "size_residuals_normalized_loss")):
            out[name + sfx] = w_box * x * k
    return out
