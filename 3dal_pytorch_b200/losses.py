"""Loss modules with the reference's names and forward signatures (forward only; training backward is not
built yet).  FrustumPointNetLossOneBoxEst / FrustumPointNetLossTwoBoxEst (tools/static_model.py:348-517) and
DynamicModelLoss (tools/dynamic_model.py:321-398): same arguments, same keys in the returned dict."""
import torch
import torch.nn as nn

from . import _lib, ops

N_PARTIAL = 1024


def _six(output, suffix, logits, mask_label, center_label, hcls, hres, scls, sres):
    """-> (6,) tensor [mask, centre, heading-class, size-class, heading-residual, size-residual] (unweighted)."""
    f = lambda t: t.float().contiguous()
    center = f(output["center" + suffix])
    bs = center.shape[0]
    dev = center.device
    out = torch.empty((6,), device=dev, dtype=torch.float32)
    ws = torch.empty((N_PARTIAL,), device=dev, dtype=torch.float32)
    lg = f(logits) if logits is not None else None
    ml = f(mask_label).view(-1) if logits is not None else None
    M = lg.shape[0] * lg.shape[1] if lg is not None else 0
    args = [center, f(center_label), f(output["heading_scores" + suffix]), hcls.long().contiguous(),
            f(output["heading_residuals_normalized" + suffix]), f(hres), f(output["size_scores" + suffix]),
            scls.long().contiguous(), f(output["size_residuals_normalized" + suffix]), f(sres)]
    ops._need_cuda(*args)
    _lib.check(_lib.lib().al3d_loss_forward(lg.data_ptr() if lg is not None else None, ml.data_ptr() if ml is not None else None,
                                            M, *[a.data_ptr() for a in args], bs, ws.data_ptr(), N_PARTIAL, out.data_ptr(),
                                            ops._stream()), "loss_forward")
    return out


class FrustumPointNetLossOneBoxEst(nn.Module):
    def forward(self, output, mask_label, center_label, heading_class_label, heading_residuals_label, size_class_label,
                size_residuals_label, w_box=1.0):
        t = _six(output, "", output["logits"], mask_label, center_label, heading_class_label, heading_residuals_label,
                 size_class_label, size_residuals_label)
        mask, c, h, s, hr, sr = t.unbind(0)
        total = mask + w_box * (c * 10 + h + s + hr * 20 + sr * 20)
        return {"total_loss": total, "mask_loss": mask, "center_loss": w_box * c * 10, "heading_class_loss": w_box * h,
                "size_class_loss": w_box * s, "heading_residuals_normalized_loss": w_box * hr * 20,
                "size_residuals_normalized_loss": w_box * sr * 20}


class DynamicModelLoss(FrustumPointNetLossOneBoxEst):
    pass


class FrustumPointNetLossTwoBoxEst(nn.Module):
    def forward(self, output, mask_label, center_label, heading_class_label, heading_residuals_label, size_class_label,
                size_residuals_label, w_box=1.0):
        one = _six(output, "_one", output["logits"], mask_label, center_label, heading_class_label, heading_residuals_label,
                   size_class_label, size_residuals_label)
        two = _six(output, "_two", None, None, center_label, output["heading_class_label_two"],
                   output["heading_residuals_label_two"], size_class_label, size_residuals_label)
        mask, c1, h1, s1, hr1, sr1 = one.unbind(0)
        _, c2, h2, s2, hr2, sr2 = two.unbind(0)
        total = mask + w_box * (c1 * 10 + h1 + s1 + hr1 * 20 + sr1 * 20 + c2 * 10 + h2 + s2 + hr2 * 20 + sr2 * 20)
        return {"total_loss": total, "mask_loss": mask,
                "center_loss_one": w_box * c1 * 10, "center_loss_two": w_box * c2 * 10,
                "heading_class_loss_one": w_box * h1, "heading_class_loss_two": w_box * h2,
                "size_class_loss_one": w_box * s1, "size_class_loss_two": w_box * s2,
                "heading_residuals_normalized_loss_one": w_box * hr1 * 20, "heading_residuals_normalized_loss_two": w_box * hr2 * 20,
                "size_residuals_normalized_loss_one": w_box * sr1 * 20, "size_residuals_normalized_loss_two": w_box * sr2 * 20}
