"""Shared test helpers: golden-fixture loading and weight regeneration."""
import hashlib
import importlib
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

pkg = importlib.import_module("3dal_pytorch_b200")
synth = importlib.import_module("3dal_pytorch_b200.synth")
spec = importlib.import_module("3dal_pytorch_b200.spec")


def sd_checksum(sd):
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def load_model_case(name):
    """Returns (fixture dict, state_dict, pts (bs,C,n) strided view, aux tensor, bbox_gt)."""
    z = dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))
    kind = str(z["kind"])
    sd = synth.random_state_dict(kind, seed=int(z["wseed"]), randomize_bn=bool(int(z["calibrated"])))
    if int(z["calibrated"]):
        std, q = float(z["calib_std"]), float(z["calib_q"])
        sd["ins_seg.dconv5.weight"].mul_(1.0 / std)
        sd["ins_seg.dconv5.bias"].mul_(1.0 / std)
        sd["ins_seg.dconv5.bias"][1] -= q / std
    assert sd_checksum(sd) == str(z["sd_sha256"]), "weight generator drifted from the golden fixture"
    if "compact" in z:
        # BASELINE-size fixtures: inputs regenerated from the seed (checked against the stored hash); logits and the
        # bit-packed mask are stored once and shown under both policies
        bs, n = int(z["bs"]), int(z["n"])
        if kind == "dynamic":
            tr = synth.dynamic_tracks(bs, npoints=n // 5, seed=int(z["dseed"]))
            z["pts_pm"], z["aux"] = tr["pts_pm"], tr["box_sm"]
        else:
            tr = synth.static_tracks(bs, n=n, seed=int(z["dseed"]))
            z["pts_pm"], z["aux"] = tr["pts_pm"], tr["init_box"]
        z["bbox_gt"] = tr["bbox_gt"]
        h = hashlib.sha256()
        for a in (z["pts_pm"], z["aux"], z["bbox_gt"]):
            h.update(np.ascontiguousarray(a).tobytes())
        assert h.hexdigest() == str(z["inputs_sha256"]), "input generator drifted from the golden fixture"
        mask = np.unpackbits(z.pop("common/mask"), axis=1)[:, :n].astype(bool)
        logits = z.pop("common/logits")
        for policy in ("numpy_legacy", "strided"):
            z[policy + "/mask"], z[policy + "/logits"] = mask, logits
    pts = torch.from_numpy(z["pts_pm"]).transpose(2, 1)
    aux = torch.from_numpy(z["aux"])
    if kind == "dynamic":
        aux = aux.transpose(2, 1)
    return z, sd, pts, aux, torch.from_numpy(z["bbox_gt"])


def rel_err(a, b):
    """max |a-b| / max(|b|, tiny) over the whole tensor (the 'relative' of BASELINE.json's
    north_star is per tensor: logits and heads have common-mode offsets far above their spread)."""
    a = torch.as_tensor(a).double()
    b = torch.as_tensor(b).double()
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-30))


def bf16_round(t):
    return t.to(torch.bfloat16).to(torch.float32)


def fold_state_dict(sd, block, table):
    """BN-fold a reference-format state_dict block -> {layer: (W', b')} (fp32, same maths as
    3dal_pytorch_b200.engine.fold_block, restated here so the tests do not depend on it)."""
    out = {}
    for lname, bn, cin, cout, kind in table:
        w = sd["%s.%s.weight" % (block, lname)].reshape(cout, cin).float()
        b = sd["%s.%s.bias" % (block, lname)].float()
        if bn is not None:
            p = "%s.%s." % (block, bn)
            a = sd[p + "weight"].float() / torch.sqrt(sd[p + "running_var"].float() + 1e-5)
            w = w * a[:, None]
            b = a * (b - sd[p + "running_mean"].float()) + sd[p + "bias"].float()
        out[lname] = (w.contiguous(), b.contiguous())
    return out


def emulate_chain_bf16(fw, names, x):
    """Numerics model of chain_max_kernel: fp32 first layer, bf16 activations / weights with fp32
    accumulation in the MMA layers, max over points, then bias + ReLU.  x (bs,C,n) -> (bs,last)."""
    h = x.transpose(2, 1).float()                                    # (bs,n,C)
    w, b = fw[names[0]]
    h = bf16_round(torch.relu(h @ w.t() + b))
    for nm in names[1:-1]:
        w, b = fw[nm]
        h = bf16_round(torch.relu(h @ bf16_round(w).t() + b))
    w, b = fw[names[-1]]
    return torch.relu((h @ bf16_round(w).t()).max(dim=1)[0] + b)


def emulate_seg_bf16(fw, x):
    """Numerics model of the bf16 segmentation path -> logits (bs,n,2)."""
    g = emulate_chain_bf16(fw, ["conv1", "conv2", "conv3", "conv4", "conv5"], x)
    h = x.transpose(2, 1).float()
    h = bf16_round(torch.relu(h @ fw["conv1"][0].t() + fw["conv1"][1]))
    o2 = bf16_round(torch.relu(h @ bf16_round(fw["conv2"][0]).t() + fw["conv2"][1]))
    wd1, bd1 = fw["dconv1"]
    gb = g @ wd1[:, 64:].t() + bd1                                    # fp32 per-object bias
    d = bf16_round(torch.relu(o2 @ bf16_round(wd1[:, :64]).t() + gb[:, None, :]))
    d = bf16_round(torch.relu(d @ bf16_round(fw["dconv2"][0]).t() + fw["dconv2"][1]))
    d = bf16_round(torch.relu(d @ bf16_round(fw["dconv3"][0]).t() + fw["dconv3"][1]))
    d = torch.relu(d @ bf16_round(fw["dconv4"][0]).t() + fw["dconv4"][1])     # stays fp32
    return d @ fw["dconv5"][0].t() + fw["dconv5"][1]


def _round_mantissa(t, bits):
    """Round-to-nearest-even to `bits` explicit mantissa bits (fp32 in, fp32 out)."""
    i = t.contiguous().view(torch.int32)
    drop = 23 - bits
    bias = ((i >> drop) & 1) + (1 << (drop - 1)) - 1
    return (((i + bias) >> drop) << drop).view(torch.float32)


def bf16x2_round(t):
    """hi + lo of the bf16x3 kernels: bf16(x) + bf16(x - bf16(x))."""
    hi = bf16_round(t)
    return hi + bf16_round(t - hi)


def f16_round(t):
    return t.clamp(max=65504.0).half().float()


def f16x2_round(t):
    hi = f16_round(t)
    return hi + (t.clamp(max=65504.0) - hi).half().float()


def emulate_seg_mixed(fw, x, conv5_f16=True, d2_mode=2):
    """Numerics model of the split-precision segmentation kernels (csrc/chain_split.cu) evaluated in float64 on rounded
    operands: every layer bf16 hi+lo x hi+lo, except conv5 (fp16 x fp16 when conv5_f16) and dconv2 (d2_mode 1: fp16 x
    fp16, 2: fp16 hi+lo x fp16).  -> logits (bs,n,2) float64, global feature (bs,1024) float64."""
    A, W = bf16x2_round, bf16x2_round

    def lin(a, w, ra=A, rw=W):
        return ra(a).double() @ rw(w).double().t()

    h = x.transpose(2, 1).float()
    o1 = torch.relu(h @ fw["conv1"][0].t() + fw["conv1"][1])
    o2 = torch.relu(lin(o1, fw["conv2"][0]) + fw["conv2"][1]).float()
    o3 = torch.relu(lin(o2, fw["conv3"][0]) + fw["conv3"][1]).float()
    o4 = torch.relu(lin(o3, fw["conv4"][0]) + fw["conv4"][1]).float()
    r5 = (f16_round, f16_round) if conv5_f16 else (A, W)
    g = torch.relu(lin(o4, fw["conv5"][0], *r5).max(dim=1)[0] + fw["conv5"][1])
    wd1, bd1 = fw["dconv1"]
    gb = g @ wd1[:, 64:].double().t() + bd1
    d = torch.relu(lin(o2, wd1[:, :64]) + gb[:, None, :]).float()
    r2 = {0: (A, W), 1: (f16_round, f16_round), 2: (f16x2_round, f16_round)}[d2_mode]
    d = torch.relu(lin(d, fw["dconv2"][0], *r2) + fw["dconv2"][1]).float()
    d = torch.relu(lin(d, fw["dconv3"][0]) + fw["dconv3"][1]).float()
    d = torch.relu(lin(d, fw["dconv4"][0]) + fw["dconv4"][1])
    return d @ fw["dconv5"][0].double().t() + fw["dconv5"][1], g
