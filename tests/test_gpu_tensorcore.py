"""GPU tests of the tcgen05 kernels against a bf16 numerics model (tight) and the fp32 oracle (loose)."""
import importlib
import os

import numpy as np
import pytest
import torch

from helpers import (bf16x2_round, emulate_chain_bf16, emulate_seg_bf16, emulate_seg_mixed, f16_round, fold_state_dict, rel_err, spec,
                     synth)  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
eb = importlib.import_module("3dal_pytorch_b200.engine_bf16")


@pytest.mark.parametrize("N,K", [(64, 64), (128, 64), (128, 128), (256, 256), (16, 16), (96, 48), (32, 512)])
def test_umma_selftest(N, K):
    torch.manual_seed(N * 1000 + K)
    a = torch.randn(128, K, device=DEV)
    b = torch.randn(N, K, device=DEV)
    ref = a.to(torch.bfloat16).float() @ b.to(torch.bfloat16).float().t()
    d = eb.umma_selftest(a, b, swap=False)
    err = rel_err(d.cpu(), ref.cpu())
    if err >= 1e-5:
        d2 = eb.umma_selftest(a, b, swap=True)
        pytest.fail("descriptor convention wrong: err(lbo=R*16,sbo=128)=%g err(swapped)=%g" % (err, rel_err(d2.cpu(), ref.cpu())))


@pytest.mark.parametrize("N,K", [(64, 64), (128, 64), (128, 128), (256, 256), (64, 512)])
def test_umma_selftest_a_from_tmem(N, K):
    torch.manual_seed(N * 1000 + K + 1)
    a = torch.randn(128, K, device=DEV)
    b = torch.randn(N, K, device=DEV)
    ref = a.to(torch.bfloat16).float() @ b.to(torch.bfloat16).float().t()
    d = eb.umma_selftest_ts(a, b)
    assert rel_err(d.cpu(), ref.cpu()) < 1e-5, rel_err(d.cpu(), ref.cpu())


@pytest.mark.parametrize("N,K", [(64, 64), (128, 64), (128, 128), (256, 256)])
def test_umma_selftest_cta_pair(N, K):
    torch.manual_seed(N * 1000 + K + 2)
    a = torch.randn(256, K, device=DEV)
    b = torch.randn(N, K, device=DEV)
    ref = a.to(torch.bfloat16).float() @ b.to(torch.bfloat16).float().t()
    d = eb.umma_selftest_pair(a, b)
    assert rel_err(d.cpu(), ref.cpu()) < 1e-5, rel_err(d.cpu(), ref.cpu())


@pytest.mark.parametrize("K", [64, 128])
def test_umma_selftest_cta_pair_ss(K):
    torch.manual_seed(K + 3)
    a = torch.randn(256, K, device=DEV)
    b = torch.randn(128, K, device=DEV)
    ref = a.to(torch.bfloat16).float() @ b.to(torch.bfloat16).float().t()
    d = eb.umma_selftest_pair_ss(a, b)
    assert rel_err(d.cpu(), ref.cpu()) < 1e-5, rel_err(d.cpu(), ref.cpu())


def _dev(fw):
    return {k: (w.to(DEV), b.to(DEV)) for k, (w, b) in fw.items()}


@pytest.mark.parametrize("kind,block,table,C,n,bs", [
    ("static_one", "box_est", "static_est_layers", 3, 512, 5),
    ("dynamic", "point_emb", "point_emb_layers", 4, 2560, 3),
    ("dynamic", "box_emb", "box_emb_layers", 8, 101, 7),
    ("static_one", "box_est", "static_est_layers", 3, 130, 300),     # ragged tile, many objects (no split)
])
def test_chain_maxpool_trunks(kind, block, table, C, n, bs):
    sd = synth.random_state_dict(kind, seed=5)
    fw = _dev(fold_state_dict(sd, block, getattr(spec, table)()))
    torch.manual_seed(0)
    x_pm = torch.randn(bs, n, C, device=DEV)
    x = x_pm.transpose(2, 1)                                  # strided view
    pack = eb.pack_trunk(fw)
    got = eb.chain_maxpool(pack, x)
    emu = emulate_chain_bf16(fw, ["conv1", "conv2", "conv3", "conv4"], x)
    assert rel_err(got.cpu(), emu.cpu()) < 2e-3, rel_err(got.cpu(), emu.cpu())
    got2 = eb.chain_maxpool(pack, x.contiguous())
    assert torch.equal(got, got2)


@pytest.mark.parametrize("C,n,bs", [(3, 4096, 2), (4, 5120, 2), (3, 1000, 3), (3, 77, 40), (3, 300, 310), (3, 100, 1), (3, 129, 1),
                                    (3, 384, 600), (4, 640, 300), (3, 2000, 149)])   # many steps per CTA, odd tile counts, item changes
def test_seg_bf16_against_numerics_model(C, n, bs):
    kind = "dynamic" if C == 4 else "static_one"
    sd = synth.random_state_dict(kind, seed=6)
    fw = _dev(fold_state_dict(sd, "ins_seg", spec.seg_layers(C)))
    torch.manual_seed(1)
    x = (torch.randn(bs, n, C, device=DEV) * torch.tensor([2.0, 2.0, 0.7, 0.2][:C], device=DEV)).transpose(2, 1)
    pack = eb.pack_seg(fw, C)
    g = eb.chain_maxpool(pack.pass1, x)                     # generic chain kernel on the segmentation widths
    emu_g = emulate_chain_bf16(fw, ["conv1", "conv2", "conv3", "conv4", "conv5"], x)
    assert rel_err(g.cpu(), emu_g.cpu()) < 2e-3
    g1 = eb.seg_pass1(pack, x)                              # specialised tile-pair kernel
    assert rel_err(g1.cpu(), emu_g.cpu()) < 2e-3, rel_err(g1.cpu(), emu_g.cpu())
    logits, mask = eb.seg_forward(pack, fw, x)
    emu = emulate_seg_bf16(fw, x)
    # The kernel and the emulation round the same fp32 sums to bf16 after different accumulation orders, so single
    # activations differ by one bf16 ulp (2^-8) and the difference propagates through dconv2-5; observed <= 6e-3 of the
    # largest logit over these shapes (a wrong tile, bias or weight block shows up as >= 1e-1).
    assert rel_err(logits.cpu(), emu.cpu()) < 1e-2, rel_err(logits.cpu(), emu.cpu())
    assert torch.equal(mask, logits[..., 0] < logits[..., 1])          # mask is exact w.r.t. our logits


# ------------------------------------------------------------------------------------------------ split precision
es = importlib.import_module("3dal_pytorch_b200.engine_split")


def _fp32_chain(fw, names, x):
    h = x.transpose(2, 1).double()
    for nm in names[:-1]:
        w, b = fw[nm]
        h = torch.relu(h @ w.double().t() + b.double())
    w, b = fw[names[-1]]
    return torch.relu((h @ w.double().t()).max(dim=1)[0] + b.double())


@pytest.mark.parametrize("kind,block,table,C,n,bs", [
    ("static_one", "box_est", "static_est_layers", 3, 512, 5),
    ("dynamic", "point_emb", "point_emb_layers", 4, 2560, 3),
    ("dynamic", "box_emb", "box_emb_layers", 8, 101, 7),
    ("static_one", "box_est", "static_est_layers", 3, 130, 300),
])
def test_split_chain_trunks_match_fp64(kind, block, table, C, n, bs):
    """bf16x3 carries 16 significant bits per operand: the pooled features match an fp64 evaluation to ~1e-5."""
    sd = synth.random_state_dict(kind, seed=5)
    fw = _dev(fold_state_dict(sd, block, getattr(spec, table)()))
    torch.manual_seed(0)
    x = torch.randn(bs, n, C, device=DEV).transpose(2, 1)
    pack = es.pack_trunk(fw)
    got = es.chain_maxpool(pack, x)
    ref = _fp32_chain(fw, ["conv1", "conv2", "conv3", "conv4"], x)
    assert rel_err(got.cpu(), ref.cpu()) < 1e-4, rel_err(got.cpu(), ref.cpu())
    assert torch.equal(got, es.chain_maxpool(pack, x.contiguous()))


@pytest.mark.parametrize("C,n,bs", [(3, 4096, 2), (4, 5120, 2), (3, 1000, 3), (3, 77, 40), (3, 300, 310), (3, 100, 1), (3, 129, 1),
                                    (3, 384, 600), (3, 2000, 149)])
def test_split_seg_matches_fp64(C, n, bs):
    kind = "dynamic" if C == 4 else "static_one"
    sd = synth.random_state_dict(kind, seed=6)
    fw = _dev(fold_state_dict(sd, "ins_seg", spec.seg_layers(C)))
    torch.manual_seed(1)
    x = (torch.randn(bs, n, C, device=DEV) * torch.tensor([2.0, 2.0, 0.7, 0.2][:C], device=DEV)).transpose(2, 1)
    pack = es.pack_seg(fw, C)
    g = es.chain_maxpool(pack.pass1, x)
    ref_g = _fp32_chain(fw, ["conv1", "conv2", "conv3", "conv4", "conv5"], x)
    assert rel_err(g.cpu(), ref_g.cpu()) < 1e-4, rel_err(g.cpu(), ref_g.cpu())
    logits, mask = es.seg_forward(pack, fw, x)
    # fp64 evaluation of the second half
    h = x.transpose(2, 1).double()
    o1 = torch.relu(h @ fw["conv1"][0].double().t() + fw["conv1"][1].double())
    o2 = torch.relu(o1 @ fw["conv2"][0].double().t() + fw["conv2"][1].double())
    wd1, bd1 = fw["dconv1"]
    gb = ref_g @ wd1[:, 64:].double().t() + bd1.double()
    d = torch.relu(o2 @ wd1[:, :64].double().t() + gb[:, None, :])
    for nm in ("dconv2", "dconv3", "dconv4"):
        d = torch.relu(d @ fw[nm][0].double().t() + fw[nm][1].double())
    ref = d @ fw["dconv5"][0].double().t() + fw["dconv5"][1].double()
    assert rel_err(logits.cpu(), ref.cpu()) < 3e-4, rel_err(logits.cpu(), ref.cpu())
    assert torch.equal(mask, logits[..., 0] < logits[..., 1])


@pytest.mark.parametrize("conv5_f16,d2_mode", [(True, 2), (True, 1), (False, 2), (True, 0)])
@pytest.mark.parametrize("C,n,bs", [(3, 4096, 2), (4, 5120, 2), (3, 1000, 3), (3, 300, 310), (3, 129, 1)])
def test_mixed_seg_matches_its_numerics_model(C, n, bs, conv5_f16, d2_mode):
    """"mixed" mode: conv5 as one fp16 MMA per product, dconv2 with fp16 (hi + lo) activations x fp16 weights, the other
    layers bf16x3.  Against the float64 evaluation on operands rounded the same way (tests/helpers.emulate_seg_mixed) the
    kernels agree to ~1e-4 of the largest logit (a wrong block, format or descriptor shows up as >= 1e-1); against plain
    float64 they stay inside the 1e-3 bar."""
    kind = "dynamic" if C == 4 else "static_one"
    sd = synth.random_state_dict(kind, seed=6)
    fw = fold_state_dict(sd, "ins_seg", spec.seg_layers(C))
    torch.manual_seed(1)
    x = (torch.randn(bs, n, C) * torch.tensor([2.0, 2.0, 0.7, 0.2][:C])).transpose(2, 1)
    emu, emu_g = emulate_seg_mixed(fw, x, conv5_f16=conv5_f16, d2_mode=d2_mode)
    exact, exact_g = _fp64_seg(fw, x)
    fwd = _dev(fw)
    pack = es.SplitSegPack(fwd, C, conv5_f16=conv5_f16, d2_mode=d2_mode)
    assert pack.pass1.struct.n_blocks == 6 + (16 if conv5_f16 else 32)       # conv2-4: one (hi, lo) block each; conv5: 8 chunks x 2 k blocks
    xd = x.to(DEV)
    g = es.chain_maxpool(pack.pass1, xd)
    # (an activation that sits on an fp16 rounding boundary may round the other way than in the model: one fp16 ulp, 5e-4 of
    # that value; measured 1.6e-4 on the pooled feature)
    assert rel_err(g.cpu(), emu_g) < 5e-4, rel_err(g.cpu(), emu_g)
    logits, mask = es.seg_forward(pack, fwd, xd)
    torch.cuda.synchronize()
    assert rel_err(logits.cpu(), emu) < 3e-4, rel_err(logits.cpu(), emu)
    assert rel_err(logits.cpu(), exact) < 1e-3, rel_err(logits.cpu(), exact)
    assert torch.equal(mask, logits[..., 0] < logits[..., 1])
    # strided and contiguous inputs give the same bits
    l2, m2 = es.seg_forward(pack, fwd, xd.contiguous())
    assert torch.equal(l2, logits) and torch.equal(m2, mask)


def _fp64_seg(fw, x):
    h = x.transpose(2, 1).double()
    o1 = torch.relu(h @ fw["conv1"][0].double().t() + fw["conv1"][1].double())
    o2 = torch.relu(o1 @ fw["conv2"][0].double().t() + fw["conv2"][1].double())
    o = o2
    for nm in ("conv3", "conv4"):
        o = torch.relu(o @ fw[nm][0].double().t() + fw[nm][1].double())
    g = torch.relu((o @ fw["conv5"][0].double().t()).max(dim=1)[0] + fw["conv5"][1].double())
    wd1, bd1 = fw["dconv1"]
    gb = g @ wd1[:, 64:].double().t() + bd1.double()
    d = torch.relu(o2 @ wd1[:, :64].double().t() + gb[:, None, :])
    for nm in ("dconv2", "dconv3", "dconv4"):
        d = torch.relu(d @ fw[nm][0].double().t() + fw[nm][1].double())
    return d @ fw["dconv5"][0].double().t() + fw["dconv5"][1].double(), g


@pytest.mark.parametrize("kind,block,table,C,n,bs", [
    ("static_one", "box_est", "static_est_layers", 3, 512, 5),
    ("dynamic", "point_emb", "point_emb_layers", 4, 2560, 3),
    ("dynamic", "box_emb", "box_emb_layers", 8, 101, 7),
    ("static_one", "box_est", "static_est_layers", 3, 130, 300),
])
def test_split_chain_trunk_with_fp16_last_layer(kind, block, table, C, n, bs):
    """The single-tile trunk kernel with its max-pooled last layer as one fp16 MMA per product (last_f16): against a float64
    evaluation with that layer's operands rounded to fp16 and against plain float64."""
    sd = synth.random_state_dict(kind, seed=5)
    fw = fold_state_dict(sd, block, getattr(spec, table)())
    torch.manual_seed(0)
    x = torch.randn(bs, n, C).transpose(2, 1)
    h = x.transpose(2, 1).float()
    o = torch.relu(h @ fw["conv1"][0].t() + fw["conv1"][1])
    for nm in ("conv2", "conv3"):
        o = torch.relu(bf16x2_round(o).double() @ bf16x2_round(fw[nm][0]).double().t() + fw[nm][1]).float()
    emu = torch.relu((f16_round(o).double() @ f16_round(fw["conv4"][0]).double().t()).max(dim=1)[0] + fw["conv4"][1])
    fwd = _dev(fw)
    pack = es.pack_trunk(fwd, last_f16=True)
    assert pack.struct.last_f16 == 1
    got = es.chain_maxpool(pack, x.to(DEV))
    torch.cuda.synchronize()
    assert rel_err(got.cpu(), emu) < 5e-4, rel_err(got.cpu(), emu)
    assert rel_err(got.cpu(), _fp32_chain(fw, ["conv1", "conv2", "conv3", "conv4"], x)) < 1e-3
    assert torch.equal(got, es.chain_maxpool(pack, x.to(DEV).contiguous()))


def test_mixed_mode_saturates_above_the_fp16_range():
    """Activations above 65504 saturate in the two fp16 layers (documented); finite inputs never produce inf / NaN logits."""
    sd = synth.random_state_dict("static_one", seed=6)
    fw = _dev(fold_state_dict(sd, "ins_seg", spec.seg_layers(3)))
    x = (torch.randn(2, 512, 3, device=DEV) * 1e6).transpose(2, 1)
    pack = es.SplitSegPack(fw, 3, conv5_f16=True, d2_mode=2)
    logits, mask = es.seg_forward(pack, fw, x)
    torch.cuda.synchronize()
    assert torch.isfinite(logits).all()


def test_split_tail_pair_kernel_equals_single_cta_kernel():
    """The CTA-pair (cta_group::2) tail kernel issues the same MMAs in the same order as the single-CTA one: bit-identical
    logits, for even / odd tile counts (ghost CTA) and ragged last tiles."""
    sd = synth.random_state_dict("static_one", seed=6)
    fw = _dev(fold_state_dict(sd, "ins_seg", spec.seg_layers(3)))
    torch.manual_seed(2)
    for n, bs in [(4096, 3), (1000, 3), (129, 1), (300, 311), (128, 1)]:
        x = (torch.randn(bs, n, 3, device=DEV) * torch.tensor([2.0, 2.0, 0.7], device=DEV)).transpose(2, 1)
        keep = es.USE_PAIR_KERNEL
        try:
            es.USE_PAIR_KERNEL = True
            lp, mp = es.seg_forward(es.pack_seg(fw, 3), fw, x)
            es.USE_PAIR_KERNEL = False
            ls, ms = es.seg_forward(es.pack_seg(fw, 3), fw, x)
        finally:
            es.USE_PAIR_KERNEL = keep
        torch.cuda.synchronize()
        assert torch.equal(lp, ls) and torch.equal(mp, ms), (n, bs)
