"""bf16x3 split-precision tensor-core mode (csrc/chain_split.cu): weight packing and launch sequences.

Every folded fp32 weight matrix is split into hi = bf16(W) and lo = bf16(W - hi); each (<=128 rows x 64 K) block is
stored as a hi slot followed by a lo slot (16 KB each, KP layout of engine_bf16.kp_pack), in the order the kernel's
MMA thread consumes them.  The kernels split the activations the same way on the fly and evaluate
a_hi*w_hi + a_lo*w_hi + a_hi*w_lo with fp32 accumulation: 16 significant bits per operand.

"mixed" mode (``engine_split.mixed``): the two widest layers of the segmentation net multiply IEEE fp16 operands instead --
conv5 (128 -> 1024) with ONE MMA per product, dconv2 (512 -> 256) with fp16 hi + lo activations x fp16 weights (two MMAs) --
and every other layer stays bf16x3: 63 % of the MMAs of bf16x3.  Their weight blocks are single fp16 slots.
"""
import ctypes
import os

import torch

from . import _lib, ops
from .engine_bf16 import BLOCK_ELEMS, _pad8, _timed, check_abort, kp_pack


class SplitChainWeightsStruct(ctypes.Structure):
    _fields_ = [("c_in", ctypes.c_int32), ("w0", ctypes.c_int32), ("n_mid", ctypes.c_int32),
                ("mid", ctypes.c_int32 * 3), ("last", ctypes.c_int32), ("n_blocks", ctypes.c_int32),
                ("pair", ctypes.c_int32), ("last_f16", ctypes.c_int32),
                ("w0_w", ctypes.c_void_p), ("w0_b", ctypes.c_void_p), ("mid_b", ctypes.c_void_p),
                ("last_b", ctypes.c_void_p), ("wstream", ctypes.c_void_p)]


class SplitTailWeightsStruct(ctypes.Structure):
    _fields_ = [("c_in", ctypes.c_int32), ("d2_mode", ctypes.c_int32),
                ("w1_w", ctypes.c_void_p), ("w1_b", ctypes.c_void_p), ("b2", ctypes.c_void_p),
                ("bd2", ctypes.c_void_p), ("bd3", ctypes.c_void_p), ("bd4", ctypes.c_void_p),
                ("w5", ctypes.c_void_p), ("b5", ctypes.c_void_p), ("wstream", ctypes.c_void_p), ("wstream_pair", ctypes.c_void_p)]

# dconv2 operand modes of the tail kernel (al3d_split_tail_weights.d2_mode)
D2_BF16X3, D2_F16, D2_F16X2 = 0, 1, 2
F16_MAX = 65504.0

# The CTA-pair (cta_group::2) variant of the tail kernel halves the weight bytes per SM and doubles the ring depth; measured
# it is NOT faster (39.6 ms vs 37.3 ms at 8192 x 4096: the kernel is bound by its unhidden epilogues, not by the weight
# stream, and every hand-over gains a cluster round trip), so it is off by default.  AL3D_SPLIT_PAIR=1 selects it; the
# tests keep it bit-identical to the single-CTA kernel.
USE_PAIR_KERNEL = os.environ.get("AL3D_SPLIT_PAIR", "0") == "1"


def split_hi_lo(w):
    hi = w.to(torch.bfloat16)
    lo = (w - hi.float()).to(torch.bfloat16)
    return hi, lo


def _slot(w_bf16):
    flat = kp_pack(w_bf16)
    out = torch.zeros(BLOCK_ELEMS, dtype=torch.bfloat16, device=w_bf16.device)
    out[: flat.numel()] = flat
    return out


def _slots(w_blk):
    """(rows <= 128, 64) fp32 -> [hi slot, lo slot]."""
    hi, lo = split_hi_lo(w_blk.float())
    return [_slot(hi), _slot(lo)]


def _slot_f16(w_blk):
    """(rows <= 128, 64) fp32 -> ONE slot of IEEE fp16 in the same KP order (carried in a bf16-typed tensor: the stream is
    a byte image)."""
    R, K = w_blk.shape
    if float(w_blk.abs().max()) > F16_MAX:
        raise ValueError("a folded weight exceeds the fp16 range")
    flat = w_blk.to(torch.float16).view(R, K // 8, 8).permute(1, 0, 2).contiguous().view(-1).view(torch.bfloat16)
    out = torch.zeros(BLOCK_ELEMS, dtype=torch.bfloat16, device=w_blk.device)
    out[: flat.numel()] = flat
    return [out]


def _half_slot(w_bf16):
    flat = kp_pack(w_bf16)
    out = torch.zeros(BLOCK_ELEMS // 2, dtype=torch.bfloat16, device=w_bf16.device)
    out[: flat.numel()] = flat
    return out


def _pair_images(blocks):
    """blocks: list of (rows <= 128, 64) fp32 weight blocks in consumption order -> the two per-CTA images of the CTA-pair
    kernel: CTA r keeps rows [r*R/2, (r+1)*R/2) of every block, as 8 KB half slots (hi then lo)."""
    imgs = []
    for r in range(2):
        parts = []
        for w in blocks:
            R = w.shape[0]
            hi, lo = split_hi_lo(w[r * (R // 2):(r + 1) * (R // 2)].float())
            parts += [_half_slot(hi), _half_slot(lo)]
        imgs.append(torch.cat(parts))
    return torch.cat(imgs).contiguous()


def _layer_slots(w, f16=False):
    """Slots of one layer in (row-chunk, k-block) order."""
    N, K = w.shape
    rows = min(N, 128)
    out = []
    for r in range(0, N, rows):
        for k in range(0, K, 64):
            out += (_slot_f16 if f16 else _slots)(w[r:r + rows, k:k + 64].float())
    return out


class SplitChainPack:
    def __init__(self, fw, names, pair, last_f16=False):
        first, mids, last = names[0], names[1:-1], names[-1]
        w0, b0 = fw[first]
        self.c_in = w0.shape[1]
        self.t = {"w0_w": _pad8(w0), "w0_b": b0.contiguous(),
                  "mid_b": torch.cat([fw[m][1] for m in mids]).contiguous(), "last_b": fw[last][1].contiguous()}
        slots = []
        for m in mids:
            slots += _layer_slots(fw[m][0])
        slots += _layer_slots(fw[last][0], f16=last_f16)
        self.t["wstream"] = torch.cat(slots).contiguous()
        s = SplitChainWeightsStruct()
        s.c_in, s.w0, s.n_mid = self.c_in, w0.shape[0], len(mids)
        for i in range(3):
            s.mid[i] = fw[mids[i]][0].shape[0] if i < len(mids) else 0
        s.last, s.n_blocks, s.pair, s.last_f16 = fw[last][0].shape[0], len(slots), int(pair), int(last_f16)
        for k in ("w0_w", "w0_b", "mid_b", "last_b", "wstream"):
            setattr(s, k, self.t[k].data_ptr())
        self.struct = s
        self.last = s.last
        self.pair = bool(pair)


class SplitSegPack:
    def __init__(self, fw, c_in, conv5_f16=False, d2_mode=D2_BF16X3):
        self.pass1 = SplitChainPack(fw, ["conv1", "conv2", "conv3", "conv4", "conv5"], pair=True, last_f16=conv5_f16)
        wd1, wd2, wd3, wd4 = (fw[k][0] for k in ("dconv1", "dconv2", "dconv3", "dconv4"))

        def d1(c):                     # dconv1 output channels c*128..+128 on the 64 per-point input channels
            return [wd1[c * 128:(c + 1) * 128, 0:64]]

        def p(c):                      # dconv2 partial sum over input channels c*128..+128: (row half, k block)
            return [wd2[nc * 128:(nc + 1) * 128, c * 128 + kb * 64:c * 128 + (kb + 1) * 64] for nc in range(2) for kb in range(2)]

        pcs = [p(c) for c in range(4)]
        blocks = [fw["conv2"][0]] + d1(0) + d1(1) + d1(2) + pcs[0] + d1(3) + pcs[1] + pcs[2] + pcs[3]
        blocks += [wd3[:, kb * 64:(kb + 1) * 64] for kb in range(4)] + [wd4[:, kb * 64:(kb + 1) * 64] for kb in range(2)]
        is_d2 = {id(b) for pc in pcs for b in pc}
        slots = [sl for blk in blocks for sl in (_slot_f16(blk.float()) if (d2_mode and id(blk) in is_d2) else _slots(blk))]
        assert len(slots) == (38 if d2_mode else 54)
        assert not (d2_mode and USE_PAIR_KERNEL), "the CTA-pair tail kernel has no fp16 dconv2"
        self.t = {"w1_w": _pad8(fw["conv1"][0]), "w1_b": fw["conv1"][1].contiguous(), "b2": fw["conv2"][1].contiguous(),
                  "bd2": fw["dconv2"][1].contiguous(), "bd3": fw["dconv3"][1].contiguous(),
                  "bd4": fw["dconv4"][1].contiguous(), "w5": fw["dconv5"][0].contiguous(),
                  "b5": fw["dconv5"][1].contiguous(), "wstream": torch.cat(slots).contiguous()}
        s = SplitTailWeightsStruct()
        s.c_in, s.d2_mode = c_in, int(d2_mode)
        for k, v in self.t.items():
            setattr(s, k, v.data_ptr())
        if USE_PAIR_KERNEL:
            self.t["wstream_pair"] = _pair_images(blocks)
            s.wstream_pair = self.t["wstream_pair"].data_ptr()
        self.struct = s
        self.w_glob = wd1[:, 64:]          # the 1024-wide half of dconv1 acts on the per-object global feature: fp32
        self.b_d1 = fw["dconv1"][1]


def pack_seg(fw, c_in):
    return SplitSegPack(fw, c_in)


def pack_trunk(fw, last_f16=False):
    return SplitChainPack(fw, ["conv1", "conv2", "conv3", "conv4"], pair=False, last_f16=last_f16)


class _MixedEngine:
    """precision = "mixed": the engine interface of this module with conv5 / dconv2 of the segmentation net and the max-pooled
    last layer of the box-head / embedding trunks in fp16."""
    CONV5_F16 = os.environ.get("AL3D_MIXED_CONV5", "1") == "1"
    D2_MODE = int(os.environ.get("AL3D_MIXED_D2", str(D2_F16X2)))

    def pack_seg(self, fw, c_in):
        # a layer whose folded weights do not fit the fp16 range (BatchNorm with a tiny running variance) stays bf16x3
        fits = lambda name: float(fw[name][0].abs().max()) <= F16_MAX
        return SplitSegPack(fw, c_in, conv5_f16=self.CONV5_F16 and fits("conv5"), d2_mode=self.D2_MODE if fits("dconv2") else D2_BF16X3)

    # the max-pooled last layer of the box-head / embedding trunks (72-76 % of their MACs) as one fp16 MMA per product
    TRUNK_F16 = os.environ.get("AL3D_MIXED_TRUNK", "1") == "1"

    def pack_trunk(self, fw):
        return pack_trunk(fw, last_f16=self.TRUNK_F16 and float(fw["conv4"][0].abs().max()) <= F16_MAX)

    def seg_forward(self, pack, fw, pts):
        return seg_forward(pack, fw, pts)

    def trunk_maxpool(self, pack, fw, x):
        return trunk_maxpool(pack, fw, x)


mixed = _MixedEngine()


def chain_maxpool(pack, x, name=None):
    """x (bs,C,n) any strides -> (bs,last) fp32 = relu(max over points of the chain)."""
    ops._need_cuda(x)
    bs, C, n = x.shape
    assert C == pack.c_in, (C, pack.c_in)
    out = torch.zeros((bs, pack.last), device=x.device, dtype=torch.float32)
    sb, sc, sp = x.stride()
    check_abort("chain_maxpool_bf16x3 launch", x.device)
    name = name or ("split_chain_pair_kernel" if pack.pair else "split_chain_kernel")
    with _timed("%s[last=%d]" % (name, pack.last)):
        _lib.check(_lib.lib().al3d_chain_maxpool_bf16x3(ctypes.byref(pack.struct), x.data_ptr(), sb, sc, sp, bs, n,
                                                        out.data_ptr(), ops._stream()), "chain_maxpool_bf16x3")
    return out


def seg_forward(pack, fw, pts):
    """-> logits (bs,n,2) f32, mask (bs,n) bool."""
    bs, C, n = pts.shape
    g = chain_maxpool(pack.pass1, pts)
    # the 1024-wide half of dconv1 on the per-object global feature: a (bs x 1024) . (1024 x 512) GEMM.  From 1024 objects on
    # it runs on the tensor cores in three-way split precision (fp32-grade, csrc/gemm_split.cu; 0.36 -> 0.07 ms at 8192
    # objects), below that on the fp32 kernel.
    if bs >= 1024:
        from . import train
        gbias = train.linear_split(g, pack.w_glob, pack.b_d1, parts=3)
    else:
        gbias = ops.linear(g, pack.w_glob, pack.b_d1, act=ops.ACT_NONE, K=1024)
    logits = torch.empty((bs, n, 2), device=pts.device, dtype=torch.float32)
    mask = torch.empty((bs, n), device=pts.device, dtype=torch.bool)
    sb, sc, sp = pts.stride()
    check_abort("seg_pass2_bf16x3 launch", pts.device)
    with _timed("split_tail_kernel"):        # (the CTA-pair variant when pack.struct.wstream_pair is set)
        _lib.check(_lib.lib().al3d_seg_pass2_bf16x3(ctypes.byref(pack.struct), pts.data_ptr(), sb, sc, sp, bs, n,
                                                    gbias.data_ptr(), logits.data_ptr(), mask.data_ptr(), ops._stream()),
                   "seg_pass2_bf16x3")
    return logits, mask


def trunk_maxpool(pack, fw, x):
    return chain_maxpool(pack, x)
