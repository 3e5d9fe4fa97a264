"""BASELINE.json configs[2] at its full size on one GPU (8192 tracks x 4096 points, the workload bench.py times), checked
through size-independent properties: the mask is the comparison of the logits the kernel wrote, the forward is
deterministic run to run, and a sample of tracks spread over the whole batch agrees with the fp32 oracle evaluated on
the host (logits within the 1e-3 bar, mask bits differing only inside the guard band of tests/test_gpu_parity.py).
Named to run last: it is the largest case of the suite."""
import importlib

import numpy as np
import pytest
import torch

from helpers import rel_err, synth
from oracle import models

sm = importlib.import_module("3dal_pytorch_b200.static_model")

TRACKS, POINTS, SAMPLE = 8192, 4096, 32
TOL = 1e-3


def check_sample_against_oracle(logits, mask, ref_logits, ref_mask, tol=TOL):
    """logits (s,n,2) / mask (s,n) of the sampled tracks against the oracle's; returns the number of differing mask bits."""
    logits, ref_logits = np.asarray(logits, dtype=np.float32), np.asarray(ref_logits, dtype=np.float32)
    mask, ref_mask = np.asarray(mask, dtype=bool), np.asarray(ref_mask, dtype=bool)
    assert logits.shape == ref_logits.shape and mask.shape == ref_mask.shape
    err = rel_err(logits, ref_logits)
    assert err < tol, err
    flips = mask != ref_mask
    margin = ref_logits[..., 1] - ref_logits[..., 0]
    band = 2 * tol * float(np.abs(ref_logits).max())
    assert np.all(np.abs(margin[flips]) <= band), (int(flips.sum()), float(np.abs(margin[flips]).max()), band)
    return int(flips.sum())


def test_sample_checker_accepts_errors_inside_the_bar_and_rejects_a_wrong_mask():
    """CPU: the checker itself, on the oracle's outputs perturbed inside / outside the tolerance."""
    sd = synth.random_state_dict("static_one", seed=3)
    d = synth.static_tracks(2, n=512, seed=3)
    pts, box = torch.from_numpy(d["pts_pm"]).transpose(2, 1), torch.from_numpy(d["init_box"])
    ref = models.static_one_forward(sd, pts, box, policy="strided")
    lg, mk = ref["logits"].numpy(), ref["mask"].numpy()
    scale = float(np.abs(lg).max())
    rng = np.random.default_rng(0)
    near = lg + (rng.random(lg.shape, dtype=np.float32) - 0.5) * 2 * 2e-4 * scale
    check_sample_against_oracle(near, near[..., 0] < near[..., 1], lg, mk)
    with pytest.raises(AssertionError):
        check_sample_against_oracle(lg + 3e-3 * scale, mk, lg, mk)
    far = np.abs(lg[..., 1] - lg[..., 0]) > 4 * TOL * scale
    if far.any():
        wrong = mk.copy()
        wrong[np.nonzero(far)[0][0], np.nonzero(far)[1][0]] ^= True
        with pytest.raises(AssertionError):
            check_sample_against_oracle(lg, wrong, lg, mk)


@pytest.mark.gpu
@pytest.mark.parametrize("prec", ["mixed", "bf16x3"])
def test_full_size_batch_properties(prec):
    dev = "cuda:0"
    sd = synth.random_state_dict("static_one", seed=synth.REFERENCE_SEED)
    model = sm.StaticModelOneBoxEst().to(dev).eval()
    model.precision = prec
    model.load_state_dict(sd)
    # the calibration of bench.py: unit margin spread, 12.5 % foreground, so that mask and gather do real work
    cal = synth.static_tracks_device(256, n=POINTS, seed=1000, device=dev)
    with torch.no_grad():
        lg = model(cal["pts_pm"].transpose(2, 1), cal["init_box"], None)["logits"]
    synth.calibrate_seg_margin(sd, lg, fg_fraction=0.125)
    model.load_state_dict(sd)
    data = synth.static_tracks_device(TRACKS, n=POINTS, seed=1001, device=dev)
    pts, box = data["pts_pm"].transpose(2, 1), data["init_box"]          # the strided (T,3,n) view of the eval scripts
    with torch.no_grad():
        out = model(pts, box, None)
        torch.cuda.synchronize()
        assert out["logits"].shape == (TRACKS, POINTS, 2) and out["mask"].shape == (TRACKS, POINTS)
        assert out["mask"].dtype == torch.bool and out["center"].shape == (TRACKS, 3)
        # the mask is exactly the comparison of the logits that were written, for every one of the 33.5 M points
        assert torch.equal(out["mask"], out["logits"][..., 0] < out["logits"][..., 1])
        assert bool(torch.isfinite(out["logits"]).all()) and bool(torch.isfinite(out["center"]).all())
        # deterministic run to run (no atomics on floating-point sums anywhere on the path)
        again = model(pts, box, None)
        torch.cuda.synchronize()
        for k, v in out.items():
            if torch.is_tensor(v):
                assert torch.equal(v, again[k]), k
        del again
    # a sample spread over the whole batch against the oracle on the host
    idx = torch.arange(0, TRACKS, TRACKS // SAMPLE, device=dev)[:SAMPLE]
    ref = models.static_one_forward(sd, pts[idx].cpu(), box[idx].cpu(), policy="strided")
    check_sample_against_oracle(out["logits"][idx].cpu().numpy(), out["mask"][idx].cpu().numpy(),
                                ref["logits"].numpy(), ref["mask"].numpy())
