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
