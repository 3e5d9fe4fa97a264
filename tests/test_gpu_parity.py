"""GPU parity tests (run on the B200 box with -m gpu): the CUDA path, called through the C ABI,
against the oracle and the committed golden vectors."""
import importlib
import os

import numpy as np
import pytest
import torch

from helpers import load_model_case, rel_err, synth
from oracle import codecs, gather, models

pytestmark = pytest.mark.gpu

ops = importlib.import_module("3dal_pytorch_b200.ops")
sm = importlib.import_module("3dal_pytorch_b200.static_model")
dm = importlib.import_module("3dal_pytorch_b200.dynamic_model")

DEV = "cuda:0"
MODEL = {"static_one": sm.StaticModelOneBoxEst, "static_two": sm.StaticModelTwoBoxEst, "dynamic": dm.DynamicModel}
# Float tolerance of BASELINE.json's north_star: 1e-3 relative (per tensor, to max|ref|).  The fp32 SIMT mode is held
# to 1e-4 and the split-precision tensor-core mode (bf16x3) to the 1e-3 bar itself (measured ~5e-5,
# profiles/r2_parity_per_tensor.jsonl); plain bf16 (8-bit mantissa) states its own, looser tolerance.
# "mixed" (bf16x3 with conv5 / dconv2 of the segmentation net in fp16) is held to the same 1e-3 bar (emulated 4e-5 .. 1.1e-4
# on these cases, profiles/r2_precision_study_mixed.txt).
TOL = {"fp32": 1e-4, "bf16x3": 1e-3, "mixed": 1e-3, "bf16": 3e-2}
# A mask bit is `l0 < l1` of OUR logits; against the fp32 reference it may differ only where the reference margin
# |l1 - l0| is inside the logit tolerance (2 * TOL * max|logit|: both logits can move).
PRECISIONS = ["fp32", "bf16x3", "mixed", "bf16"]


def _check_outputs(out, z, policy, prec, sd, pts, aux, gt):
    kind = str(z["kind"])
    keys = [k.split("/", 1)[1] for k in z if k.startswith(policy + "/")]
    assert set(keys) == set(out), (sorted(keys), sorted(out))
    ref_logits, ref_mask = z[policy + "/logits"], z[policy + "/mask"]
    got_mask = out["mask"].cpu().numpy()
    assert got_mask.dtype == ref_mask.dtype and got_mask.shape == ref_mask.shape
    assert rel_err(out["logits"].cpu().numpy(), ref_logits) < TOL[prec], ("logits", rel_err(out["logits"].cpu().numpy(), ref_logits))
    # the mask is exactly the comparison of the logits the kernel wrote ...
    lg = out["logits"].cpu().numpy()
    assert np.array_equal(got_mask, lg[..., 0] < lg[..., 1])
    # ... and differs from the reference mask only inside the guard band
    flips = got_mask != ref_mask
    margin = ref_logits[..., 1] - ref_logits[..., 0]
    band = TOL[prec] * float(np.abs(ref_logits).max()) * 2
    assert np.all(np.abs(margin[flips]) <= band), (int(flips.sum()), float(np.abs(margin[flips]).max()), band)
    # Everything after the mask (gather, box heads, decode, labels) is a function of the exact foreground set, so it
    # is compared with the reference algorithm evaluated ON THE KERNEL'S MASK (the oracle, bit-identical to the
    # reference on the reference's own mask: tests/test_oracle_golden.py); when no bit flipped that is the golden
    # fixture itself.
    if flips.any():
        if policy == "numpy_legacy":
            np.random.seed(int(z["rng_seed"]))
        ref = models.FORWARDS[kind](sd, pts, aux, gt, policy=policy, mask_override=torch.from_numpy(got_mask))
        ref = {k: torch.as_tensor(v).cpu().numpy() for k, v in ref.items()}
    else:
        ref = {k: z[policy + "/" + k] for k in keys}
    for k in keys:
        if k in ("mask", "logits"):
            continue
        r, g = ref[k], out[k].cpu().numpy()
        assert g.shape == r.shape and g.dtype == r.dtype, (k, g.shape, r.shape, g.dtype, r.dtype)
        if np.issubdtype(r.dtype, np.integer):
            if prec != "bf16":
                assert np.array_equal(g, r), k                    # class labels: exact
            else:
                assert np.mean(g != r) <= 0.1, k                 # a label may move when the heading sits on a bin edge
        else:
            assert rel_err(g, r) < TOL[prec], (k, rel_err(g, r))
    return int(flips.sum())


@pytest.mark.parametrize("name", ["static_one", "static_two", "dynamic", "static_one_default_init",
                                  "static_one_cfg1", "static_two_cfg1", "dynamic_cfg2"])
@pytest.mark.parametrize("policy", ["strided", "numpy_legacy"])
@pytest.mark.parametrize("prec", PRECISIONS)
def test_forward_matches_reference_golden(name, policy, prec):
    z, sd, pts, aux, gt = load_model_case(name)
    model = MODEL[str(z["kind"])]().to(DEV).eval()
    model.load_state_dict(sd)
    model.precision = prec
    model.gather_policy = policy
    if policy == "numpy_legacy":
        np.random.seed(int(z["rng_seed"]))
    out = model(pts.to(DEV), aux.to(DEV), gt.to(DEV))     # pts keeps the strided (bs,C,n) view
    torch.cuda.synchronize()
    _check_outputs(out, z, policy, prec, sd, pts, aux, gt)


@pytest.mark.parametrize("prec", PRECISIONS)
def test_contiguous_and_strided_inputs_agree_bitwise(prec):
    z, sd, pts, aux, gt = load_model_case("static_one")
    model = sm.StaticModelOneBoxEst().to(DEV).eval()
    model.load_state_dict(sd)
    model.precision = prec
    a = model(pts.to(DEV), aux.to(DEV), gt.to(DEV))
    b = model(pts.to(DEV).contiguous(), aux.to(DEV), gt.to(DEV))
    for k in a:
        assert torch.equal(a[k], b[k]), k


@pytest.mark.parametrize("prec", PRECISIONS)
def test_sharding_by_track_is_bitwise_equivalent(prec):
    """Tracks are independent: running two halves (what two ranks do) == running the whole batch."""
    z, sd, pts, aux, gt = load_model_case("static_two")
    model = sm.StaticModelTwoBoxEst().to(DEV).eval()
    model.load_state_dict(sd)
    model.precision = prec
    p, a, g = pts.to(DEV), aux.to(DEV), gt.to(DEV)
    whole = model(p, a, g)
    h = p.shape[0] // 2
    lo, hi = model(p[:h], a[:h], g[:h]), model(p[h:], a[h:], g[h:])
    for k in whole:
        assert torch.equal(whole[k], torch.cat([lo[k], hi[k]], 0)), k


def test_linear_f32_against_torch():
    torch.manual_seed(0)
    for M, K, N in [(1, 8, 4), (300, 64, 64), (1000, 1088, 512), (257, 128, 2), (64, 256, 39), (513, 100, 130)]:
        a = torch.randn(M, K, device=DEV)
        w = torch.randn(N, K, device=DEV) / K ** 0.5
        b = torch.randn(N, device=DEV)
        y = ops.linear(a, w, b, act=ops.ACT_RELU)
        ref = torch.relu(a.double() @ w.double().t() + b.double())
        assert rel_err(y.cpu(), ref.cpu()) < 1e-5, (M, K, N)
    # per-group bias + max-pool epilogue, groups that straddle tiles
    a = torch.randn(5 * 100, 64, device=DEV)
    w = torch.randn(96, 64, device=DEV) / 8
    rb = torch.randn(5, 96, device=DEV)
    g = torch.zeros(5, 96, device=DEV)
    ops.linear(a, w, None, rowbias=rb, rows_per_group=100, max_out=g)
    ref = torch.relu((a.double() @ w.double().t()).view(5, 100, 96) + rb.double()[:, None, :]).max(1)[0]
    assert rel_err(g.cpu(), ref.cpu()) < 1e-5


def test_pointwise_first_strided_and_contiguous():
    torch.manual_seed(1)
    for C, cout in [(3, 64), (4, 64), (8, 64), (3, 128)]:
        x_pm = torch.randn(3, 200, C, device=DEV)
        w = torch.randn(cout, C, device=DEV)
        b = torch.randn(cout, device=DEV)
        ref = torch.relu(x_pm.double() @ w.double().t() + b.double()).view(-1, cout)
        y1 = ops.pointwise_first(x_pm.transpose(2, 1), w, b)
        y2 = ops.pointwise_first(x_pm.transpose(2, 1).contiguous(), w, b)
        assert torch.equal(y1, y2)
        assert rel_err(y1.cpu(), ref.cpu()) < 1e-6


def test_mask_compact_and_gather_bit_exact():
    rng = np.random.default_rng(0)
    for bs, n, n_pts, C in [(5, 1000, 512, 3), (3, 4096, 512, 3), (2, 5120, 2560, 4), (4, 37, 16, 3), (1, 1, 4, 3)]:
        logits = rng.normal(size=(bs, n, 2)).astype(np.float32)
        logits[0, :, 1] = -10                                   # empty object
        if bs > 1:
            logits[1, :, 1] = 10                                # all foreground
        logits[-1, 0, 0] = np.nan
        pts = rng.normal(size=(bs, n, C)).astype(np.float32)
        ref_mask = gather.mask_from_logits(torch.from_numpy(logits)).numpy()
        x = torch.from_numpy(pts).to(DEV).transpose(2, 1)
        mask, pos, count = ops.mask_compact(logits=torch.from_numpy(logits).to(DEV))
        assert np.array_equal(mask.cpu().numpy(), ref_mask)
        assert np.array_equal(count.cpu().numpy(), ref_mask.sum(1))
        for i in range(bs):
            assert np.array_equal(pos[i, :int(count[i])].cpu().numpy(), np.nonzero(ref_mask[i])[0])
        # strided device rule
        out, idx = ops.gather_fg(x, pos, count, n_pts, want_indices=True)
        ro, ri = gather.gather_object_pts(pts.transpose(0, 2, 1), ref_mask, n_pts, "strided")
        assert np.array_equal(out.cpu().numpy(), ro) and np.array_equal(idx.cpu().numpy(), ri)
        # numpy legacy replay through a choice table
        eng = importlib.import_module("3dal_pytorch_b200.engine")
        np.random.seed(77)
        table = torch.from_numpy(eng.choice_table_numpy_legacy(count.cpu().numpy(), n_pts)).to(DEV)
        out, idx = ops.gather_fg(x, pos, count, n_pts, choice=table, want_indices=True)
        np.random.seed(77)
        ro, ri = gather.gather_object_pts(pts.transpose(0, 2, 1), ref_mask, n_pts, "numpy_legacy")
        assert np.array_equal(out.cpu().numpy(), ro) and np.array_equal(idx.cpu().numpy(), ri)


def test_decode_and_retransform_match_oracle():
    torch.manual_seed(3)
    bs, m = 64, 512
    center = torch.randn(bs, 3)
    hs, hr = torch.randn(bs, 12), torch.randn(bs, 12) * 0.3
    ss, sr = torch.randn(bs, 3), torch.randn(bs, 3, 3) * 0.2
    init_box = torch.randn(bs, 7)
    gt = torch.randn(bs, 7) * 3
    hs[0, 3] = hs[0, 7] = 9.0                                   # tie -> first maximum
    ref64, hcls, scls = models.decode_box(center, hs, hr, ss, sr, init_box[:, 6])
    box, cls = ops.decode_boxes(center.to(DEV), hs.to(DEV), hr.to(DEV), ss.to(DEV), sr.to(DEV),
                                base_heading=init_box.to(DEV)[:, 6])
    assert np.array_equal(cls.cpu().numpy()[:, 0], hcls) and np.array_equal(cls.cpu().numpy()[:, 1], scls)
    assert np.array_equal(box.cpu().numpy(), ref64.astype(np.float32))      # f64 math, one final rounding
    obj = torch.randn(bs, 3, m)
    box_one = torch.from_numpy(ref64.astype(np.float32))
    o2, c2, r2 = ops.twostage_retransform(obj.to(DEV), init_box.to(DEV), box_one.to(DEV), gt.to(DEV))
    for i in range(bs):
        p = models._rotz(init_box[i, 6]) @ obj[i] + init_box[i, :3][:, None] - box_one[i, :3][:, None]
        p = models._rotz(-box_one[i, 6]) @ p
        assert rel_err(o2[i].cpu(), p) < 1e-5
        cid, res = codecs.angle2class_f32((gt[i, 6] - box_one[i, 6]).numpy(), 12)
        assert int(c2[i]) == cid and np.float32(r2[i].item()) == np.float32(res), i


def test_fused_fc_heads_match_the_layer_by_layer_path():
    """csrc/heads.cu (one launch: fc1 -> fc2 -> fc3 -> parse -> centre add -> decode) against al3d_linear_f32 per layer +
    al3d_parse_heads + al3d_decode_boxes.  Same fp32 arithmetic, another summation order: floats to 1e-5, classes exact."""
    from helpers import fold_state_dict, spec
    eng = importlib.import_module("3dal_pytorch_b200.engine")
    torch.manual_seed(7)
    for bs in (1, 16, 37, 300):
        sd = synth.random_state_dict("static_one", seed=9)
        fw = {k: (w.to(DEV), b.to(DEV)) for k, (w, b) in fold_state_dict(sd, "box_est", spec.static_est_layers()).items()}
        g = torch.relu(torch.randn(bs, 512, device=DEV))
        init_box = torch.randn(bs, 7, device=DEV)
        ref_pred = eng.fc_chain(fw, g, ("fc1", "fc2", "fc3"))
        ref = ops.parse_heads(ref_pred, add=init_box)
        ref_box, ref_cls = ops.decode_boxes(ref["center"], ref["heading_scores"], ref["heading_residuals"], ref["size_scores"],
                                            ref["size_residuals"], base_heading=init_box[:, 6])
        got = ops.fc_chain(g, eng.fc_layers_t(fw, ("fc1", "fc2", "fc3")), heads=True, add=init_box, base_heading=init_box[:, 6],
                           want_box=True)
        for k in ref:
            assert rel_err(got[k].cpu(), ref[k].cpu()) < 1e-5, (bs, k)
        assert torch.equal(got["_cls"], ref_cls) and rel_err(got["_box"].cpu(), ref_box.cpu()) < 1e-5
    # dynamic head: two concatenated inputs, embeddings without a head epilogue
    sd = synth.random_state_dict("dynamic", seed=9)
    fwe = {k: (w.to(DEV), b.to(DEV)) for k, (w, b) in fold_state_dict(sd, "box_est", spec.dynamic_est_layers()).items()}
    fwp = {k: (w.to(DEV), b.to(DEV)) for k, (w, b) in fold_state_dict(sd, "point_emb", spec.point_emb_layers()).items()}
    gp = torch.relu(torch.randn(50, 512, device=DEV))
    pe_ref = eng.fc_chain(fwp, gp, ("fc1", "fc2"))
    pe = ops.fc_chain(gp, eng.fc_layers_t(fwp, ("fc1", "fc2")))
    assert rel_err(pe.cpu(), pe_ref.cpu()) < 1e-5
    be = torch.relu(torch.randn(50, 128, device=DEV))
    ref = ops.parse_heads(eng.fc_chain(fwe, torch.cat([pe_ref, be], 1), ("fc1", "fc2", "fc3")))
    got = ops.fc_chain(pe_ref, eng.fc_layers_t(fwe, ("fc1", "fc2", "fc3")), x1=be, heads=True)
    for k in ref:
        assert rel_err(got[k].cpu(), ref[k].cpu()) < 1e-5, k
