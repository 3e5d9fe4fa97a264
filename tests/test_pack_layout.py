"""CPU check of the tensor-core weight packing: decode the packed block stream exactly the way the
kernels consume it (block order of csrc/chain_bf16.cu, KP plane layout of csrc/umma.cuh) and make
sure it reproduces the folded weights."""
import importlib

import torch

from helpers import fold_state_dict, spec, synth

eb = importlib.import_module("3dal_pytorch_b200.engine_bf16")


def kp_unpack(flat, R, K):
    """Inverse of the KP layout: element (r,k) lives at (k//8)*(R*8) + r*8 + k%8."""
    out = torch.empty(R, K)
    for k in range(K):
        for r in range(R):
            out[r, k] = float(flat[(k // 8) * (R * 8) + r * 8 + (k % 8)])
    return out


def _blocks(stream):
    return stream.view(-1, eb.BLOCK_ELEMS)


def test_kp_pack_is_plane_major():
    w = (torch.arange(16 * 64) % 251).float().view(16, 64)                 # exactly representable in bf16
    assert torch.equal(kp_unpack(eb.kp_pack(w).float(), 16, 64), w)


def test_chain_stream_order_matches_kernel_loops():
    sd = synth.random_state_dict("static_one", seed=1)
    fw = fold_state_dict(sd, "box_est", spec.static_est_layers())
    pack = eb.pack_trunk(fw)
    blocks = _blocks(pack.t["wstream"])
    blk = 0
    prev = fw["conv1"][0].shape[0]
    for name in ("conv2", "conv3", "conv4"):          # kernel: for row-chunk: for k-block
        w = fw[name][0].to(torch.bfloat16).float()
        N, K = w.shape
        assert K == prev
        rows = min(N, 128)
        for nc in range(N // rows):
            for kb in range(K // 64):
                got = kp_unpack(blocks[blk][: rows * 64].float(), rows, 64)
                assert torch.equal(got, w[nc * rows:(nc + 1) * rows, kb * 64:(kb + 1) * 64]), (name, nc, kb)
                blk += 1
        prev = N
    assert blk == pack.struct.n_blocks == blocks.shape[0]
    assert (pack.struct.w0, pack.struct.n_mid, list(pack.struct.mid)[:2], pack.struct.last) == (128, 2, [128, 256], 512)
    assert pack.t["w0_w"].shape == (8, 128) and torch.equal(pack.t["w0_w"][:3, :], fw["conv1"][0].t())


def test_pass2_stream_order_matches_kernel_schedule():
    sd = synth.random_state_dict("dynamic", seed=2)
    fw = fold_state_dict(sd, "ins_seg", spec.seg_layers(4))
    pack = eb.pack_seg(fw, 4)
    bf = lambda n: fw[n][0].to(torch.bfloat16).float()
    wd1, wd2, wd3, wd4 = bf("dconv1"), bf("dconv2"), bf("dconv3"), bf("dconv4")
    expect = [("conv2", bf("conv2"))]
    d1 = lambda c: [("d1_%d" % c, wd1[c * 64:(c + 1) * 64, :64])]
    d2 = lambda pc: [("d2_%d" % pc, wd2[:, pc * 64:(pc + 1) * 64])]
    # order of use by the MMA thread in seg_pass2_kernel
    expect += d1(0) + d1(1) + d1(2)
    for kc in range(8):
        expect += d2(kc)
        if kc + 3 < 8:
            expect += d1(kc + 3)
    expect += [("d3_%d" % kb, wd3[:, kb * 64:(kb + 1) * 64]) for kb in range(4)]
    expect += [("d4_%d" % kb, wd4[:, kb * 64:(kb + 1) * 64]) for kb in range(2)]
    assert len(expect) == 23

    # p2_block_bytes / p2_half_off of csrc/chain_bf16.cu
    def block_bytes(blk):
        if blk <= 3:
            return 8192
        if blk < 14:
            return 32768 if (blk - 4) % 2 == 0 else 8192
        if blk < 17:
            return 32768
        return 16384
    assert [w.shape[0] * 128 for _, w in expect] == [block_bytes(i) for i in range(23)]
    half_total = sum(block_bytes(i) // 2 for i in range(23))
    assert half_total == 217088
    stream = pack.t["wstream"]
    assert stream.numel() * 2 == 2 * half_total
    # CTA r of a pair holds rows [r*R/2, (r+1)*R/2) of every block (cta_group::2 splits B's N rows over the two CTAs)
    for r in range(2):
        off = r * half_total // 2
        for i, (name, w) in enumerate(expect):
            rows = w.shape[0] // 2
            got = kp_unpack(stream[off: off + rows * 64].float(), rows, 64)
            assert torch.equal(got, w[r * rows:(r + 1) * rows]), (name, r)
            off += rows * 64
    assert pack.struct.c_in == 4 and pack.w_glob.shape == (512, 1024)


def test_mixed_mode_streams_single_fp16_slots_for_conv5_and_dconv2():
    """"mixed" mode: conv5 (pass 1) and the sixteen dconv2 partial-sum blocks (tail) are ONE slot of IEEE fp16 each, in the
    consumption order of csrc/chain_split.cu; every other block keeps its (hi, lo) bf16 slots."""
    es = importlib.import_module("3dal_pytorch_b200.engine_split")
    sd = synth.random_state_dict("static_one", seed=4)
    fw = fold_state_dict(sd, "ins_seg", spec.seg_layers(3))
    ref = es.SplitSegPack(fw, 3)
    mix = es.mixed.pack_seg(fw, 3)
    assert (ref.pass1.struct.last_f16, ref.struct.d2_mode) == (0, 0)
    assert (mix.pass1.struct.last_f16, mix.struct.d2_mode) == (1, es.D2_F16X2)
    # pass 1: conv2-4 as before (one block x (hi, lo) each), then conv5's 8 row chunks x 2 k blocks, single slots
    a, b = _blocks(ref.pass1.t["wstream"]), _blocks(mix.pass1.t["wstream"])
    assert (a.shape[0], b.shape[0]) == (6 + 32, 6 + 16) and mix.pass1.struct.n_blocks == 22
    assert torch.equal(a[:6], b[:6])
    w5 = fw["conv5"][0]
    for cc in range(8):
        for kb in range(2):
            got = kp_unpack(b[6 + cc * 2 + kb].view(torch.float16).float(), 128, 64)
            assert torch.equal(got, w5[cc * 128:(cc + 1) * 128, kb * 64:(kb + 1) * 64].half().float()), (cc, kb)
    # tail: conv2 | d1(0) d1(1) d1(2) | p(0) | d1(3) | p(1) p(2) p(3) | dconv3 | dconv4 with p(c) = 4 single fp16 slots
    a, b = _blocks(ref.t["wstream"]), _blocks(mix.t["wstream"])
    assert (a.shape[0], b.shape[0]) == (54, 38)
    wd2 = fw["dconv2"][0]
    ia = ib = 0
    for kind, c in [("x", 0)] * 4 + [("p", 0), ("x", 0), ("p", 1), ("p", 2), ("p", 3)] + [("x", 0)] * 6:
        if kind == "x":
            assert torch.equal(a[ia:ia + 2], b[ib:ib + 2])
            ia, ib = ia + 2, ib + 2
        else:
            for nc in range(2):
                for kb in range(2):
                    got = kp_unpack(b[ib].view(torch.float16).float(), 128, 64)
                    assert torch.equal(got, wd2[nc * 128:(nc + 1) * 128, c * 128 + kb * 64:c * 128 + (kb + 1) * 64].half().float()), (c, nc, kb)
                    ia, ib = ia + 2, ib + 1
    assert (ia, ib) == (54, 38)


def test_mixed_mode_keeps_a_layer_bf16x3_when_its_weights_exceed_the_fp16_range():
    es = importlib.import_module("3dal_pytorch_b200.engine_split")
    sd = synth.random_state_dict("static_one", seed=4)
    fw = fold_state_dict(sd, "ins_seg", spec.seg_layers(3))
    w, b = fw["conv5"]
    w = w.clone()
    w[3, 5] = 7.0e4
    fw["conv5"] = (w, b)
    pk = es.mixed.pack_seg(fw, 3)
    assert (pk.pass1.struct.last_f16, pk.pass1.struct.n_blocks, pk.struct.d2_mode) == (0, 6 + 32, es.D2_F16X2)
    w2, b2 = fw["dconv2"]
    fw["dconv2"] = (w2 * 1e6, b2)
    pk = es.mixed.pack_seg(fw, 3)
    assert (pk.pass1.struct.last_f16, pk.struct.d2_mode, pk.t["wstream"].numel()) == (0, es.D2_BF16X3, 54 * eb.BLOCK_ELEMS)
