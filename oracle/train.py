"""Training-mode forward of the auto-label models with torch autograd.  TEST INFRASTRUCTURE (oracle).

Functional restatement, driven by a dict of leaf tensors in the reference's state_dict layout, of the reference
modules under ``model.train()``: BatchNorm with batch statistics (and the running-stat update), Dropout as an injected
multiplier (tools/static_model.py:264,293), the non-differentiable foreground gather (:33-47), the losses
(oracle/losses.py).  tests/test_train_oracle_vs_reference.py pins it against the real reference modules
(loss, every parameter gradient, running statistics) when /root/reference is mounted; the GPU tests compare the CUDA
training step with it.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import gather, losses, models

BN_EPS = 1e-5
MOMENTUM = 0.1


def _bn_train(P, name, y, stats):
    """nn.BatchNorm1d training forward on (bs,C,n) or (bs,C); records the updated running statistics in `stats`."""
    rm, rv = P[name + ".running_mean"].clone(), P[name + ".running_var"].clone()
    out = F.batch_norm(y, rm, rv, P[name + ".weight"], P[name + ".bias"], training=True, momentum=MOMENTUM, eps=BN_EPS)
    stats[name + ".running_mean"], stats[name + ".running_var"] = rm, rv
    return out


def _conv_bn_relu(P, conv, bn, x, stats):
    return F.relu(_bn_train(P, bn, F.conv1d(x, P[conv + ".weight"], P[conv + ".bias"]), stats))


def _fc_bn_relu(P, fc, bn, x, stats):
    return F.relu(_bn_train(P, bn, F.linear(x, P[fc + ".weight"], P[fc + ".bias"]), stats))


def seg_forward(P, pts, drop_mult, stats, prefix="ins_seg"):
    p = prefix + "."
    o1 = _conv_bn_relu(P, p + "conv1", p + "bn1", pts, stats)
    o2 = _conv_bn_relu(P, p + "conv2", p + "bn2", o1, stats)
    o3 = _conv_bn_relu(P, p + "conv3", p + "bn3", o2, stats)
    o4 = _conv_bn_relu(P, p + "conv4", p + "bn4", o3, stats)
    o5 = _conv_bn_relu(P, p + "conv5", p + "bn5", o4, stats)
    g = o5.max(dim=2, keepdim=True)[0]
    cat = torch.cat([o2, g.expand(-1, -1, pts.shape[2])], dim=1)
    x = _conv_bn_relu(P, p + "dconv1", p + "dbn1", cat, stats)
    x = _conv_bn_relu(P, p + "dconv2", p + "dbn2", x, stats)
    x = _conv_bn_relu(P, p + "dconv3", p + "dbn3", x, stats)
    x = _conv_bn_relu(P, p + "dconv4", p + "dbn4", x, stats)
    if drop_mult is not None:
        x = x * drop_mult
    x = F.conv1d(x, P[p + "dconv5.weight"], P[p + "dconv5.bias"])
    return x.transpose(2, 1).contiguous()


def trunk(P, prefix, x, stats):
    p = prefix + "."
    for i in (1, 2, 3, 4):
        x = _conv_bn_relu(P, p + "conv%d" % i, p + "bn%d" % i, x, stats)
    return x.max(dim=2)[0]


def static_est(P, prefix, obj, stats):
    p = prefix + "."
    g = trunk(P, prefix, obj, stats)
    x = _fc_bn_relu(P, p + "fc1", p + "fcbn1", g, stats)
    x = _fc_bn_relu(P, p + "fc2", p + "fcbn2", x, stats)
    return F.linear(x, P[p + "fc3.weight"], P[p + "fc3.bias"])


def embedding(P, prefix, x, stats):
    p = prefix + "."
    g = trunk(P, prefix, x, stats)
    x = _fc_bn_relu(P, p + "fc1", p + "fcbn1", g, stats)
    return _fc_bn_relu(P, p + "fc2", p + "fcbn2", x, stats)


def _gather(pts, logits, n_obj, policy):
    mask = gather.mask_from_logits(logits.detach())
    obj, _ = gather.gather_object_pts(pts.detach().cpu().numpy(), mask.cpu().numpy(), n_obj, policy)
    return torch.from_numpy(obj).to(pts.device).to(pts.dtype), mask


def _heads(pred):
    c, hs, hrn, hr, ss, srn, sr = models.parse_heads(pred)
    return {"center_boxnet": c, "heading_scores": hs, "heading_residuals_normalized": hrn, "heading_residuals": hr,
            "size_scores": ss, "size_residuals_normalized": srn, "size_residuals": sr}


def leaf_params(sd):
    """state_dict -> dict of tensors where every floating-point parameter (not buffer) is a leaf requiring grad."""
    P = {}
    for k, v in sd.items():
        t = v.clone()
        if t.is_floating_point() and not (k.endswith("running_mean") or k.endswith("running_var")):
            t.requires_grad_(True)
        P[k] = t
    return P


def static_one_step(sd, pts, init_box, labels, drop_mult=None, policy="strided", w_box=1.0):
    """-> (losses dict, output dict, {param name: grad}, {buffer name: updated running stat})."""
    P, stats = leaf_params(sd), {}
    logits = seg_forward(P, pts, drop_mult, stats)
    obj, mask = _gather(pts[:, :3, :], logits, models.NUM_OBJECT_POINT, policy)
    out = _heads(static_est(P, "box_est", obj, stats))
    out.update({"logits": logits, "mask": mask, "center": out["center_boxnet"] + init_box[:, :3]})
    ls = losses.one_box(out, *labels, w_box=w_box)
    ls["total_loss"].backward()
    return ls, out, {k: v.grad for k, v in P.items() if v.requires_grad}, stats


def static_two_step(sd, pts, init_box, bbox_gt, labels, drop_mult=None, policy="strided", w_box=1.0):
    P, stats = leaf_params(sd), {}
    logits = seg_forward(P, pts, drop_mult, stats)
    obj, mask = _gather(pts[:, :3, :], logits, models.NUM_OBJECT_POINT, policy)
    one = _heads(static_est(P, "box_est_one", obj, stats))
    c1 = one["center_boxnet"] + init_box[:, :3]
    with torch.no_grad():
        box_one64, _, _ = models.decode_box(c1, one["heading_scores"], one["heading_residuals"], one["size_scores"],
                                            one["size_residuals"], init_box[:, 6])
        box_one = torch.from_numpy(box_one64).float()          # the reference casts box_one to float32 (:190)
        bs = pts.shape[0]
        obj2 = obj.clone()
        cls2 = np.zeros((bs,), dtype=np.int64)
        res2 = np.zeros((bs,), dtype=np.float32)
        from . import codecs
        for i in range(bs):
            dt = obj2.dtype
            p = models._rotz(init_box[i, 6]).to(dt) @ obj2[i] + init_box[i, :3][:, None] - box_one[i, :3][:, None].to(dt)
            obj2[i] = models._rotz(-box_one[i, 6]).to(dt) @ p
            cls2[i], res2[i] = codecs.angle2class_f32((bbox_gt[i, 6].float() - box_one[i, 6]).numpy().astype(np.float32), 12)
    two = _heads(static_est(P, "box_est_two", obj2, stats))
    c2 = two["center_boxnet"] + c1
    out = {"logits": logits, "mask": mask, "center_one": c1, "center_two": c2, "box_one": box_one,
           "heading_class_label_two": torch.from_numpy(cls2), "heading_residuals_label_two": torch.from_numpy(res2).to(pts.dtype)}
    for k in ("heading_scores", "heading_residuals_normalized", "heading_residuals", "size_scores", "size_residuals_normalized",
              "size_residuals"):
        out[k + "_one"], out[k + "_two"] = one[k], two[k]
    ls = losses.two_box(out, *labels, w_box=w_box)
    ls["total_loss"].backward()
    return ls, out, {k: v.grad for k, v in P.items() if v.requires_grad}, stats


def dynamic_step(sd, pts, box, labels, drop_mult=None, policy="strided", w_box=1.0):
    P, stats = leaf_params(sd), {}
    logits = seg_forward(P, pts, drop_mult, stats)
    obj, mask = _gather(pts[:, :4, :], logits, models.NUM_FRAME * models.NUM_OBJECT_POINT, policy)
    pe = embedding(P, "point_emb", obj, stats)
    be = embedding(P, "box_emb", box, stats)
    x = torch.cat([pe, be], dim=1)
    x = _fc_bn_relu(P, "box_est.fc1", "box_est.fcbn1", x, stats)
    x = _fc_bn_relu(P, "box_est.fc2", "box_est.fcbn2", x, stats)
    out = _heads(F.linear(x, P["box_est.fc3.weight"], P["box_est.fc3.bias"]))
    out.update({"logits": logits, "mask": mask, "center": out["center_boxnet"]})
    ls = losses.one_box(out, *labels, w_box=w_box)
    ls["total_loss"].backward()
    return ls, out, {k: v.grad for k, v in P.items() if v.requires_grad}, stats
