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
