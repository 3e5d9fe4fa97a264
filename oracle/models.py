"""fp32 restatement of the three auto-label model forwards.  TEST INFRASTRUCTURE (oracle).

Functional torch code driven by a reference-format ``state_dict`` (same keys/shapes as the
reference's nn.Modules), eval-mode semantics only: BatchNorm uses running
statistics, Dropout is the identity.  Follows, layer for layer:

  seg net        PointNetInstanceSeg.forward     tools/static_model.py:271-296 (dynamic_model.py:187-212)
  static head    PointNetEstimation.forward      tools/static_model.py:320-339
  head parsing   parse_output_to_tensors         tools/static_model.py:64-96
  one-box model  StaticModelOneBoxEst.forward    tools/static_model.py:117-146
  two-box model  StaticModelTwoBoxEst.forward    tools/static_model.py:158-239
  point emb.     PointEmbedding.forward          tools/dynamic_model.py:234-249
  box emb.       BoxEmbedding.forward            tools/dynamic_model.py:271-286
  dynamic head   PointNetEstimation.forward      tools/dynamic_model.py:300-312
  dynamic model  DynamicModel.forward            tools/dynamic_model.py:121-155

Runs on whatever device the inputs live on (CPU in the tests / cpu_baseline).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import codecs, gather

NUM_HEADING_BIN = codecs.NUM_HEADING_BIN
NUM_SIZE_CLUSTER = codecs.NUM_SIZE_CLUSTER
NUM_OBJECT_POINT = 512          # tools/static_model.py:14
NUM_FRAME = 5                   # tools/dynamic_model.py:16
BN_EPS = 1e-5                   # nn.BatchNorm1d default


def _conv_bn_relu(sd, conv, bn, x):
    """Conv1d(k=1) -> BatchNorm1d(eval) -> ReLU on (bs, C, n)."""
    y = F.conv1d(x, sd[conv + ".weight"], sd[conv + ".bias"])
    y = F.batch_norm(y, sd[bn + ".running_mean"], sd[bn + ".running_var"], sd[bn + ".weight"], sd[bn + ".bias"],
                     training=False, eps=BN_EPS)
    return F.relu(y)


def _fc_bn_relu(sd, fc, bn, x):
    y = F.linear(x, sd[fc + ".weight"], sd[fc + ".bias"])
    y = F.batch_norm(y, sd[bn + ".running_mean"], sd[bn + ".running_var"], sd[bn + ".weight"], sd[bn + ".bias"],
                     training=False, eps=BN_EPS)
    return F.relu(y)


def seg_forward(sd, pts, prefix="ins_seg"):
    """pts (bs,C,n) -> logits (bs,n,2).  Also returns the per-object global feature (bs,1024)."""
    p = prefix + "."
    o1 = _conv_bn_relu(sd, p + "conv1", p + "bn1", pts)
    o2 = _conv_bn_relu(sd, p + "conv2", p + "bn2", o1)
    o3 = _conv_bn_relu(sd, p + "conv3", p + "bn3", o2)
    o4 = _conv_bn_relu(sd, p + "conv4", p + "bn4", o3)
    o5 = _conv_bn_relu(sd, p + "conv5", p + "bn5", o4)
    g = o5.max(dim=2, keepdim=True)[0]                      # (bs,1024,1)
    cat = torch.cat([o2, g.expand(-1, -1, pts.shape[2])], dim=1)   # (bs,1088,n)
    x = _conv_bn_relu(sd, p + "dconv1", p + "dbn1", cat)
    x = _conv_bn_relu(sd, p + "dconv2", p + "dbn2", x)
    x = _conv_bn_relu(sd, p + "dconv3", p + "dbn3", x)
    x = _conv_bn_relu(sd, p + "dconv4", p + "dbn4", x)
    x = F.conv1d(x, sd[p + "dconv5.weight"], sd[p + "dconv5.bias"])
    return x.transpose(2, 1).contiguous(), g.squeeze(2)


def trunk_maxpool(sd, prefix, x):
    """conv1-4 (+BN+ReLU) and max over the point/step axis: (bs,C,m) -> (bs,512)."""
    p = prefix + "."
    for i in (1, 2, 3, 4):
        x = _conv_bn_relu(sd, p + "conv%d" % i, p + "bn%d" % i, x)
    return x.max(dim=2)[0]


def static_est_forward(sd, prefix, obj_pts):
    """(bs,3,512) -> (bs,39)."""
    p = prefix + "."
    g = trunk_maxpool(sd, prefix, obj_pts)
    x = _fc_bn_relu(sd, p + "fc1", p + "fcbn1", g)
    x = _fc_bn_relu(sd, p + "fc2", p + "fcbn2", x)
    return F.linear(x, sd[p + "fc3.weight"], sd[p + "fc3.bias"])


def embedding_forward(sd, prefix, x):
    """PointEmbedding (bs,4,2560)->(bs,256) / BoxEmbedding (bs,8,101)->(bs,128)."""
    p = prefix + "."
    g = trunk_maxpool(sd, prefix, x)
    x = _fc_bn_relu(sd, p + "fc1", p + "fcbn1", g)
    return _fc_bn_relu(sd, p + "fc2", p + "fcbn2", x)


def dynamic_est_forward(sd, prefix, emb):
    p = prefix + "."
    x = _fc_bn_relu(sd, p + "fc1", p + "fcbn1", emb)
    x = _fc_bn_relu(sd, p + "fc2", p + "fcbn2", x)
    return F.linear(x, sd[p + "fc3.weight"], sd[p + "fc3.bias"])


def parse_heads(box_pred):
    """Slice the 39-vector; residuals are scaled by pi/12 and by the anchor sizes."""
    bs = box_pred.shape[0]
    H, S = NUM_HEADING_BIN, NUM_SIZE_CLUSTER
    center = box_pred[:, 0:3]
    h_scores = box_pred[:, 3:3 + H]
    h_res_n = box_pred[:, 3 + H:3 + 2 * H]
    h_res = h_res_n * (np.pi / H)
    s_scores = box_pred[:, 3 + 2 * H:3 + 2 * H + S]
    s_res_n = box_pred[:, 3 + 2 * H + S:3 + 2 * H + 4 * S].contiguous().view(bs, S, 3)
    anchors = torch.from_numpy(codecs.MEAN_SIZE_ARR).to(box_pred.dtype).to(box_pred.device)
    s_res = s_res_n * anchors.unsqueeze(0)
    return center, h_scores, h_res_n, h_res, s_scores, s_res_n, s_res


def _mask_and_gather(pts, logits, n_obj_pts, policy, mask_override=None):
    """mask_override: evaluate the stages AFTER the mask on a given (bs,n) bool mask instead of the oracle's own --
    the parity tests use it to check the box heads of a reduced-precision run whose mask differs from the fp32
    mask in a few boundary points (the gather, and with it every head, depends on the exact foreground set)."""
    mask = gather.mask_from_logits(logits) if mask_override is None else torch.as_tensor(mask_override).bool()
    obj, idx = gather.gather_object_pts(pts.cpu().numpy(), mask.cpu().numpy(), n_obj_pts, policy)
    return torch.from_numpy(obj).to(pts.device), mask, idx


@torch.no_grad()
def static_one_forward(sd, pts, init_box, bbox_gt=None, policy="numpy_legacy", mask_override=None):
    logits, _ = seg_forward(sd, pts)
    obj, mask, idx = _mask_and_gather(pts[:, :3, :], logits, NUM_OBJECT_POINT, policy, mask_override)
    pred = static_est_forward(sd, "box_est", obj)
    c, hs, hrn, hr, ss, srn, sr = parse_heads(pred)
    return {
        "logits": logits, "mask": mask, "center_boxnet": c, "heading_scores": hs,
        "heading_residuals_normalized": hrn, "heading_residuals": hr, "size_scores": ss,
        "size_residuals_normalized": srn, "size_residuals": sr, "center": c + init_box[:, :3],
        "_object_pts": obj, "_indices": torch.from_numpy(idx),
    }


def decode_box(center, h_scores, h_res, s_scores, s_res, base_heading):
    """The host decode both the two-box forward (tools/static_model.py:178-190) and the eval loops
    (tools/static_eval.py:270-288) perform: argmax heads, class2size / class2angle in float64,
    add the base heading; returns (bs,7) float64 [cx,cy,cz,l,w,h,heading] and the two class ids."""
    hs = h_scores.cpu().numpy()
    ss = s_scores.cpu().numpy()
    hr = h_res.cpu().numpy()
    sr = s_res.cpu().numpy()
    base = base_heading.cpu().numpy()
    bs = hs.shape[0]
    hcls = np.argmax(hs, 1)
    scls = np.argmax(ss, 1)
    out = np.zeros((bs, 7))
    out[:, 0:3] = center.cpu().numpy()
    for i in range(bs):
        out[i, 3:6] = codecs.class2size(scls[i], sr[i, scls[i], :])
        ang = codecs.class2angle(hcls[i], hr[i, hcls[i]], NUM_HEADING_BIN)
        out[i, 6] = ang + base[i]
    return out, hcls, scls


def _rotz(angle):
    c, s = torch.cos(angle), torch.sin(angle)
    z, o = torch.zeros_like(c), torch.ones_like(c)
    return torch.stack([torch.stack([c, -s, z]), torch.stack([s, c, z]), torch.stack([z, z, o])])


@torch.no_grad()
def static_two_forward(sd, pts, init_box, bbox_gt, policy="numpy_legacy", mask_override=None):
    logits, _ = seg_forward(sd, pts)
    obj, mask, idx = _mask_and_gather(pts[:, :3, :], logits, NUM_OBJECT_POINT, policy, mask_override)
    pred1 = static_est_forward(sd, "box_est_one", obj)
    c1, hs1, hrn1, hr1, ss1, srn1, sr1 = parse_heads(pred1)
    c1 = c1 + init_box[:, :3]
    box_one64, _, _ = decode_box(c1, hs1, hr1, ss1, sr1, init_box[:, 6])
    box_one = torch.from_numpy(box_one64).float().to(pts.device)

    bs = pts.shape[0]
    obj2 = obj.clone()
    cls2 = np.zeros((bs,), dtype=np.int64)
    res2 = np.zeros((bs,), dtype=np.float32)
    for i in range(bs):
        # back to the frame of init_box, then into the frame of box_one (tools/static_model.py:195-200)
        p = _rotz(init_box[i, 6]) @ obj2[i]
        p = p + init_box[i, :3][:, None]
        p = p - box_one[i, :3][:, None]
        obj2[i] = _rotz(-box_one[i, 6]) @ p
        d = (bbox_gt[i, 6] - box_one[i, 6]).cpu().numpy().astype(np.float32)
        cls2[i], res2[i] = codecs.angle2class_f32(d, NUM_HEADING_BIN)

    pred2 = static_est_forward(sd, "box_est_two", obj2)
    c2, hs2, hrn2, hr2, ss2, srn2, sr2 = parse_heads(pred2)
    c2 = c2 + c1
    return {
        "logits": logits, "mask": mask,
        "heading_scores_one": hs1, "heading_residuals_normalized_one": hrn1, "heading_residuals_one": hr1,
        "size_scores_one": ss1, "size_residuals_normalized_one": srn1, "size_residuals_one": sr1,
        "center_one": c1, "box_one": box_one,
        "heading_scores_two": hs2, "heading_residuals_normalized_two": hrn2, "heading_residuals_two": hr2,
        "size_scores_two": ss2, "size_residuals_normalized_two": srn2, "size_residuals_two": sr2,
        "center_two": c2,
        "heading_class_label_two": torch.from_numpy(cls2).to(pts.device),
        "heading_residuals_label_two": torch.from_numpy(res2).to(pts.device),
        "center": c2, "heading_scores": hs2, "heading_residuals": hr2, "size_scores": ss2, "size_residuals": sr2,
        "_object_pts": obj, "_object_pts_two": obj2, "_indices": torch.from_numpy(idx),
    }


@torch.no_grad()
def dynamic_forward(sd, pts, box, bbox_gt=None, policy="numpy_legacy", mask_override=None):
    logits, _ = seg_forward(sd, pts)
    obj, mask, idx = _mask_and_gather(pts[:, :4, :], logits, NUM_FRAME * NUM_OBJECT_POINT, policy, mask_override)
    pe = embedding_forward(sd, "point_emb", obj)
    be = embedding_forward(sd, "box_emb", box)
    pred = dynamic_est_forward(sd, "box_est", torch.cat([pe, be], dim=1))
    c, hs, hrn, hr, ss, srn, sr = parse_heads(pred)
    return {
        "logits": logits, "mask": mask, "center": c, "heading_scores": hs,
        "heading_residuals_normalized": hrn, "heading_residuals": hr, "size_scores": ss,
        "size_residuals_normalized": srn, "size_residuals": sr,
        "_object_pts": obj, "_indices": torch.from_numpy(idx),
    }


FORWARDS = {"static_one": static_one_forward, "static_two": static_two_forward, "dynamic": dynamic_forward}


def flops_per_object(kind, n_points):
    """Factored algorithmic FLOPs per object (1 MAC = 2 FLOP; the 1024-wide
    global-feature half of ins_seg.dconv1 is counted once per object)."""
    c = 4 if kind == "dynamic" else 3
    per_pt = c * 64 + 64 * 64 + 64 * 64 + 64 * 128 + 128 * 1024 + 64 * 512 + 512 * 256 + 256 * 128 + 128 * 128 + 128 * 2
    macs = per_pt * n_points + 1024 * 512
    est = (3 * 128 + 128 * 128 + 128 * 256 + 256 * 512) * 512 + 512 * 512 + 512 * 256 + 256 * 39
    if kind == "static_one":
        macs += est
    elif kind == "static_two":
        macs += 2 * est
    else:
        macs += (4 * 64 + 64 * 128 + 128 * 256 + 256 * 512) * 2560 + 512 * 512 + 512 * 256
        macs += (8 * 64 + 64 * 64 + 64 * 128 + 128 * 512) * 101 + 512 * 128 + 128 * 128
        macs += 384 * 128 + 128 * 128 + 128 * 39
    return 2.0 * macs
