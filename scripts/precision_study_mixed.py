"""CPU emulation of the "mixed" precision mode (tests/helpers.emulate_seg_mixed) on the golden cases of the parity tests:
error of the logits against the fixture's fp32 reference logits, mask bits that differ and the largest reference margin
among them, for the candidate operand formats of conv5 / dconv2.

    python scripts/precision_study_mixed.py > profiles/r2_precision_study_mixed.txt
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import emulate_seg_mixed, fold_state_dict, load_model_case, spec  # noqa: E402

torch.set_num_threads(os.cpu_count() or 8)
MODES = [("bf16x3 (all layers hi+lo)", False, 0), ("conv5 f16", True, 0), ("dconv2 f16 hi+lo x f16", False, 2),
         ("mixed: conv5 f16, dconv2 f16 hi+lo x f16", True, 2), ("conv5 f16, dconv2 f16 x f16", True, 1)]
for name in ("static_one", "static_one_cfg1", "static_one_default_init", "dynamic", "dynamic_cfg2"):
    z, sd, pts, aux, gt = load_model_case(name)
    C = pts.shape[1]
    fw = fold_state_dict(sd, "ins_seg", spec.seg_layers(C))
    ref = z["strided/logits"]
    ref_mask = z["strided/mask"]
    margin = ref[..., 1] - ref[..., 0]
    print("%s: %d x %d points, max|logit| %.3f, foreground %.3f" % (name, pts.shape[0], pts.shape[2], np.abs(ref).max(), ref_mask.mean()))
    for label, c5, d2 in MODES:
        lg, _ = emulate_seg_mixed(fw, pts, conv5_f16=c5, d2_mode=d2)
        lg = lg.float().numpy()
        err = np.abs(lg - ref).max() / np.abs(ref).max()
        flips = (lg[..., 0] < lg[..., 1]) != ref_mask
        print("  %-42s logits rel %.2e  mask bits differing %d / %d  largest |l1-l0| among them %.2e"
              % (label, err, int(flips.sum()), flips.size, float(np.abs(margin[flips]).max()) if flips.any() else 0.0))

# ---- robustness over weight seeds: random-init weights with randomised BatchNorm statistics, margin calibrated to std 1 and
# 30 % foreground (the generator of the parity measurements), 8 tracks x 4096 points each
from helpers import synth  # noqa: E402
from oracle import models  # noqa: E402

print("weight seeds, static 8 x 4096 (error relative to max|logit|, which the calibration leaves between ~5 and ~70):")
worst = 0.0
for seed in (synth.REFERENCE_SEED, 1, 2, 3, 4, 5, 6, 7, 11, 13, 17, 19):
    sd = synth.random_state_dict("static_one", seed=seed)
    d = synth.static_tracks(8, n=4096, seed=7)
    pts = torch.from_numpy(d["pts_pm"]).transpose(2, 1)
    lg, _ = models.seg_forward(sd, pts)
    synth.calibrate_seg_margin(sd, lg, 0.3)
    ref, _ = models.seg_forward(sd, pts)
    ref = ref.numpy()
    fw = fold_state_dict(sd, "ins_seg", spec.seg_layers(3))
    row = []
    for label, c5, d2 in MODES[:1] + MODES[3:]:
        lg = emulate_seg_mixed(fw, pts, conv5_f16=c5, d2_mode=d2)[0].float().numpy()
        row.append(np.abs(lg - ref).max() / np.abs(ref).max())
    worst = max(worst, row[1])
    print("  seed %-9d max|logit| %6.2f   bf16x3 %.2e   mixed %.2e   mixed with dconv2 f16 x f16 %.2e" % (seed, np.abs(ref).max(), *row))
print("worst mixed over the seeds: %.2e (bar 1e-3)" % worst)
