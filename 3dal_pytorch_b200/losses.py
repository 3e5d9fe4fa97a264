"""Loss modules with the reference's names and forward signatures.  FrustumPointNetLossOneBoxEst /
FrustumPointNetLossTwoBoxEst (tools/static_model.py:348-517) and DynamicModelLoss (tools/dynamic_model.py:321-398): same
arguments, same keys in the returned dict.  Forward and backward are fused CUDA kernels (csrc/loss.cu, csrc/train.cu);
the modules are differentiable through ``torch.autograd`` (``losses['total_loss'].backward()`` as in
tools/static_train.py:85-89)."""
import torch
import torch.nn as nn

from . import _lib, ops, spec

N_PARTIAL = 1024


def _six_raw(center, hs, hrn, ss, srn, logits, mask_label, center_label, hcls, hres, scls, sres):
    """-> (6,) tensor [mask, centre, heading-class, size-class, heading-residual, size-residual] (unweighted)."""
    bs = center.shape[0]
    dev = center.device
    out = torch.empty((6,), device=dev, dtype=torch.float32)
    ws = torch.empty((N_PARTIAL,), device=dev, dtype=torch.float32)
    M = logits.shape[0] * logits.shape[1] if logits is not None else 0
    args = [center, center_label, hs, hcls, hrn, hres, ss, scls, srn, sres]
    ops._need_cuda(*args)
    _lib.check(_lib.lib().al3d_loss_forward(logits.data_ptr() if logits is not None else None,
                                            mask_label.data_ptr() if logits is not None else None,
                                            M, *[a.data_ptr() for a in args], bs, ws.data_ptr(), N_PARTIAL, out.data_ptr(),
                                            ops._stream()), "loss_forward")
    return out


class _SixFn(torch.autograd.Function):
    """The six unweighted loss terms as one differentiable op: backward is one fused kernel that evaluates the gradient
    of sum_t w[t] * term_t for the incoming weights w = d(total)/d(term)."""

    @staticmethod
    def forward(ctx, logits, center, hs, hrn, ss, srn, mask_label, center_label, hcls, hres, scls, sres):
        six = _six_raw(center, hs, hrn, ss, srn, logits, mask_label, center_label, hcls, hres, scls, sres)
        ctx.save_for_backward(*[t for t in (logits, center, hs, hrn, ss, srn, mask_label, center_label, hcls, hres, scls, sres)
                                if t is not None])
        ctx.has_logits = logits is not None
        return six

    @staticmethod
    def backward(ctx, dsix):
        saved = list(ctx.saved_tensors)
        if ctx.has_logits:
            logits, center, hs, hrn, ss, srn, mask_label, center_label, hcls, hres, scls, sres = saved
        else:
            center, hs, hrn, ss, srn, center_label, hcls, hres, scls, sres = saved
            logits = mask_label = None
        bs = center.shape[0]
        dev = center.device
        w6 = dsix.float().contiguous()
        dlogits = torch.empty_like(logits) if logits is not None else None
        dbox = torch.empty((bs, spec.HEAD_WIDTH), device=dev, dtype=torch.float32)
        M = logits.shape[0] * logits.shape[1] if logits is not None else 0
        p = lambda t: None if t is None else t.data_ptr()
        _lib.check(_lib.lib().al3d_loss_backward(p(logits), p(mask_label), M, p(center), p(center_label), p(hs), p(hcls), p(hrn),
                                                 p(hres), p(ss), p(scls), p(srn), p(sres), bs, p(w6), p(dlogits), p(dbox),
                                                 ops._stream()), "loss_backward")
        H, S = spec.NUM_HEADING_BIN, spec.NUM_SIZE_CLUSTER
        return (dlogits, dbox[:, :3], dbox[:, 3:3 + H], dbox[:, 3 + H:3 + 2 * H], dbox[:, 3 + 2 * H:3 + 2 * H + S],
                dbox[:, 3 + 2 * H + S:].reshape(bs, S, 3), None, None, None, None, None, None)


def _six(output, suffix, logits, mask_label, center_label, hcls, hres, scls, sres):
    """-> (6,) tensor [mask, centre, heading-class, size-class, heading-residual, size-residual] (unweighted)."""
    f = lambda t: t.float().contiguous()
    lg = f(logits) if logits is not None else None
    ml = f(mask_label).view(-1) if logits is not None else None
    return _SixFn.apply(lg, f(output["center" + suffix]), f(output["heading_scores" + suffix]),
                        f(output["heading_residuals_normalized" + suffix]), f(output["size_scores" + suffix]),
                        f(output["size_residuals_normalized" + suffix]), ml, f(center_label), hcls.long().contiguous(), f(hres),
                        scls.long().contiguous(), f(sres))


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
