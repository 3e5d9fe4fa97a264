"""bf16 tensor-core mode: weight packing for, and launch sequences of, the tcgen05 kernels
(csrc/chain_bf16.cu).  See engine.py for the fp32 mode and the shared mask / gather / head stages.

Packing: every folded weight matrix W' (N, K) is cut into blocks of (<=128 rows) x (64 K), each
converted to bf16 and laid out K-plane-major -- ``blk.view(R, 8, 8).permute(1, 0, 2)`` -- which is
byte-for-byte the shared-memory operand layout the MMA descriptors describe, so the kernel's producer
warp only issues linear 16 KB bulk copies.  Blocks are stored in consumption order.
"""
import ctypes
import os

import torch

from . import _lib, ops

BLOCK_ELEMS = 8192           # 16 KB of bf16


class ChainWeightsStruct(ctypes.Structure):
    _fields_ = [("c_in", ctypes.c_int32), ("w0", ctypes.c_int32), ("n_mid", ctypes.c_int32),
                ("mid", ctypes.c_int32 * 3), ("last", ctypes.c_int32), ("n_blocks", ctypes.c_int32),
                ("w0_w", ctypes.c_void_p), ("w0_b", ctypes.c_void_p), ("mid_b", ctypes.c_void_p),
                ("last_b", ctypes.c_void_p), ("wstream", ctypes.c_void_p)]


class Pass1WeightsStruct(ctypes.Structure):
    _fields_ = [("c_in", ctypes.c_int32), ("reserved", ctypes.c_int32),
                ("w1_w", ctypes.c_void_p), ("w1_b", ctypes.c_void_p), ("b2", ctypes.c_void_p), ("b3", ctypes.c_void_p),
                ("b4", ctypes.c_void_p), ("b5", ctypes.c_void_p), ("wfront", ctypes.c_void_p), ("w5stream", ctypes.c_void_p),
                ("consts_host", ctypes.c_void_p)]


class Pass2WeightsStruct(ctypes.Structure):
    _fields_ = [("c_in", ctypes.c_int32), ("reserved", ctypes.c_int32),
                ("w1_w", ctypes.c_void_p), ("w1_b", ctypes.c_void_p), ("b2", ctypes.c_void_p),
                ("bd2", ctypes.c_void_p), ("bd3", ctypes.c_void_p), ("bd4", ctypes.c_void_p),
                ("w5", ctypes.c_void_p), ("b5", ctypes.c_void_p), ("wstream", ctypes.c_void_p)]


def kp_pack(w):
    """(R, K) float -> bf16 tensor of R*K elements in KP order (K/8 planes of R x 8)."""
    R, K = w.shape
    return w.to(torch.bfloat16).view(R, K // 8, 8).permute(1, 0, 2).contiguous().view(-1)


def _block(w_blk):
    flat = kp_pack(w_blk)
    out = torch.zeros(BLOCK_ELEMS, dtype=torch.bfloat16, device=w_blk.device)
    out[: flat.numel()] = flat
    return out


def _pad8(w):
    """(w0, c_in) first-layer weights -> (8, w0) transposed, zero rows for c >= c_in."""
    out = torch.zeros((8, w.shape[0]), dtype=torch.float32, device=w.device)
    out[: w.shape[1], :] = w.t()
    return out.contiguous()


def _layer_blocks(w):
    """Blocks of one layer in (row-chunk, k-block) order."""
    N, K = w.shape
    rows = min(N, 128)
    return [_block(w[r:r + rows, k:k + 64]) for r in range(0, N, rows) for k in range(0, K, 64)]


class ChainPack:
    """Packed weights of a first-layer + mid-layers + max-pooled-last-layer chain."""

    def __init__(self, fw, names):
        first, mids, last = names[0], names[1:-1], names[-1]
        w0, b0 = fw[first]
        self.c_in = w0.shape[1]
        self.t = {"w0_w": _pad8(w0), "w0_b": b0.contiguous(),
                  "mid_b": torch.cat([fw[m][1] for m in mids]).contiguous(), "last_b": fw[last][1].contiguous()}
        blocks = []
        for m in mids + [last]:
            blocks += _layer_blocks(fw[m][0])
        self.t["wstream"] = torch.cat(blocks).contiguous()
        s = ChainWeightsStruct()
        s.c_in, s.w0, s.n_mid = self.c_in, w0.shape[0], len(mids)
        for i in range(3):
            s.mid[i] = fw[mids[i]][0].shape[0] if i < len(mids) else 0
        s.last, s.n_blocks = fw[last][0].shape[0], len(blocks)
        for k in ("w0_w", "w0_b", "mid_b", "last_b", "wstream"):
            setattr(s, k, self.t[k].data_ptr())
        self.struct = s
        self.last = s.last


class SegPack:
    def __init__(self, fw, c_in):
        self.pass1 = ChainPack(fw, ["conv1", "conv2", "conv3", "conv4", "conv5"])
        wd1, wd2, wd3, wd4 = (fw[k][0] for k in ("dconv1", "dconv2", "dconv3", "dconv4"))
        mats = [fw["conv2"][0]]                     # weight sub-matrices (rows x 64 K) in the MMA thread's order of use

        def d1(c):                     # dconv1 chunk c: 64 output channels x the 64 per-point input channels
            return [wd1[c * 64:(c + 1) * 64, 0:64]]

        def d2(pc):                    # dconv2 partial sum over input channels pc*64..+64: all 256 rows, one N = 256 MMA
            return [wd2[:, pc * 64:(pc + 1) * 64]]

        mats += d1(0) + d1(1) + d1(2)
        for kc in range(8):
            mats += d2(kc)
            if kc + 3 < 8:
                mats += d1(kc + 3)
        mats += [wd3[:, kb * 64:(kb + 1) * 64] for kb in range(4)]
        mats += [wd4[:, kb * 64:(kb + 1) * 64] for kb in range(2)]
        assert len(mats) == 23
        self.p2_mats = mats
        # the kernel runs as CTA pairs (cta_group::2): CTA r keeps rows [r*R/2, (r+1)*R/2) of every block, KP-packed,
        # tightly concatenated; the two per-CTA images follow each other
        halves = [torch.cat([kp_pack(m[r * (m.shape[0] // 2):(r + 1) * (m.shape[0] // 2)]) for m in mats]) for r in range(2)]
        assert halves[0].numel() * 2 == 217088
        wstream2 = torch.cat(halves).contiguous()
        self.t = {"w1_w": _pad8(fw["conv1"][0]), "w1_b": fw["conv1"][1].contiguous(), "b2": fw["conv2"][1].contiguous(),
                  "bd2": fw["dconv2"][1].contiguous(), "bd3": fw["dconv3"][1].contiguous(),
                  "bd4": fw["dconv4"][1].contiguous(), "w5": fw["dconv5"][0].contiguous(),
                  "b5": fw["dconv5"][1].contiguous(), "wstream": wstream2}
        s = Pass2WeightsStruct()
        s.c_in = c_in
        for k, v in self.t.items():
            setattr(s, k, v.data_ptr())
        self.struct = s
        # pass 1 (conv1-5 + max), specialised kernel: conv2-4 resident, conv5 streamed in (chunk, k-block) order
        self.t1 = {"w1_w": self.t["w1_w"], "w1_b": self.t["w1_b"], "b2": self.t["b2"], "b3": fw["conv3"][1].contiguous(),
                   "b4": fw["conv4"][1].contiguous(), "b5": fw["conv5"][1].contiguous(),
                   "wfront": torch.cat([_block(fw["conv2"][0]), _block(fw["conv3"][0]), _block(fw["conv4"][0])]).contiguous(),
                   "w5stream": torch.cat(_layer_blocks(fw["conv5"][0])).contiguous()}
        s1 = Pass1WeightsStruct()
        s1.c_in = c_in
        for k, v in self.t1.items():
            setattr(s1, k, v.data_ptr())
        # small per-model constants the kernel takes in its parameter block (constant-bank operands): HOST copy
        self.consts_host = torch.cat([self.t1[k].detach().float().cpu().reshape(-1) for k in ("w1_w", "w1_b", "b2", "b3", "b4")]).contiguous()
        assert self.consts_host.numel() == 832
        s1.consts_host = self.consts_host.data_ptr()
        self.struct1 = s1
        # the 1024-wide half of dconv1 acts on the per-object global feature: kept fp32
        self.w_glob = wd1[:, 64:]
        self.b_d1 = fw["dconv1"][1]


def pack_seg(fw, c_in):
    return SegPack(fw, c_in)


def pack_trunk(fw):
    return ChainPack(fw, ["conv1", "conv2", "conv3", "conv4"])


# Optional per-kernel CUDA-event timing (bench.py's roofline leg): {"name": [(start_event, end_event), ...]}
KERNEL_EVENTS = None


class _timed:
    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if KERNEL_EVENTS is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *a):
        if KERNEL_EVENTS is not None:
            self.e1.record()
            KERNEL_EVENTS.setdefault(self.name, []).append((self.e0, self.e1))


WATCHDOG_ROLES = {
    0x0: "split_tail_pair_kernel epilogue warps", 0x1: "split_tail_pair_kernel producer / notifier / MMA issuer",
    0x2: "split_tail_kernel epilogue warps", 0x3: "split_tail_kernel producer / MMA issuer",
    0x4: "split_chain(_pair)_kernel epilogue warps", 0x5: "split_chain(_pair)_kernel producer / MMA issuer",
    0x6: "trunk_pair_kernel epilogue warps", 0x7: "trunk_pair_kernel producer / MMA issuer",
    0x8: "seg_pass1_kernel reducer / front warps", 0x9: "seg_pass1_kernel producer / MMA issuers",
    0xA: "seg_pass2_kernel loader / MMA issuers", 0xB: "seg_pass2_kernel epilogue warps",
    0xC: "chain_max_kernel producer / MMA issuer", 0xD: "chain_max_kernel epilogue warps", 0xE: "umma selftest",
}
_status_words = {}


def _status_word(dev_index):
    """ctypes view of the device's host-mapped watchdog status word (csrc/tcstatus.cu): reading it is a plain
    host load -- no CUDA call, no synchronisation."""
    w = _status_words.get(dev_index)
    if w is None:
        ptr = ctypes.c_void_p()
        with torch.cuda.device(dev_index):
            _lib.check(_lib.lib().al3d_tc_status_word_host(ctypes.byref(ptr)), "tc_status_word_host")
        w = ctypes.c_uint32.from_address(ptr.value)
        _status_words[dev_index] = w
    return w


def check_abort(what, device=None):
    """Raises if a tensor-core kernel on `device` gave up on an mbarrier wait since the last check.  Called by
    default before every tensor-core launch and at every host synchronisation point of the pipeline; costs one
    host load.  (The kernel also traps, so the caller's next CUDA synchronisation fails as well.)"""
    idx = torch.cuda.current_device() if device is None else torch.device(device).index
    w = _status_word(idx)
    code = w.value
    if code != 0:
        w.value = 0
        raise RuntimeError("libal3d: watchdog code 0x%X (%s) -- a tensor-core kernel gave up on an mbarrier wait; "
                           "outputs since the last check are invalid (noticed at: %s)"
                           % (code, WATCHDOG_ROLES.get(code >> 12, "?"), what))


def configure_watchdog(trap=True, stress_ns=0, legacy_pass1_release=False):
    """trap=False: record the code only (diagnostics); stress_ns > 0: inject pseudo-random per-role delays of up to
    that many ns before the kernels' barrier waits (protocol stress tests); legacy_pass1_release: run seg_pass1_kernel
    with the round-1 single out4_free barrier (known to deadlock; lets the stress test prove it can detect that)."""
    _lib.check(_lib.lib().al3d_tc_configure(int(bool(trap)) | (2 if legacy_pass1_release else 0), int(stress_ns)),
               "tc_configure")


def chain_maxpool(pack, x):
    """x (bs,C,n) any strides -> (bs,last) fp32 = relu(max over points of the chain)."""
    ops._need_cuda(x)
    bs, C, n = x.shape
    assert C == pack.c_in, (C, pack.c_in)
    out = torch.zeros((bs, pack.last), device=x.device, dtype=torch.float32)
    sb, sc, sp = x.stride()
    check_abort("chain_maxpool launch", x.device)
    with _timed("chain_max_kernel[last=%d]" % pack.last):
        _lib.check(_lib.lib().al3d_chain_maxpool_bf16(ctypes.byref(pack.struct), x.data_ptr(), sb, sc, sp, bs, n,
                                                      out.data_ptr(), ops._stream()), "chain_maxpool_bf16")
    return out


def seg_pass1(pack, pts):
    """ins_seg conv1-5 + max: pts (bs,C,n) any strides -> (bs,1024) fp32 global feature."""
    ops._need_cuda(pts)
    bs, C, n = pts.shape
    out = torch.zeros((bs, 1024), device=pts.device, dtype=torch.float32)
    sb, sc, sp = pts.stride()
    check_abort("seg_pass1 launch", pts.device)
    with _timed("seg_pass1_kernel"):
        _lib.check(_lib.lib().al3d_seg_pass1_bf16(ctypes.byref(pack.struct1), pts.data_ptr(), sb, sc, sp, bs, n,
                                                  out.data_ptr(), ops._stream()), "seg_pass1_bf16")
    return out


def seg_forward(pack, fw, pts):
    """-> logits (bs,n,2) f32, mask (bs,n) bool."""
    bs, C, n = pts.shape
    g = seg_pass1(pack, pts)
    gbias = ops.linear(g, pack.w_glob, pack.b_d1, act=ops.ACT_NONE, K=1024)
    logits = torch.empty((bs, n, 2), device=pts.device, dtype=torch.float32)
    mask = torch.empty((bs, n), device=pts.device, dtype=torch.bool)
    sb, sc, sp = pts.stride()
    check_abort("seg_pass2 launch", pts.device)
    with _timed("seg_pass2_kernel"):
        _lib.check(_lib.lib().al3d_seg_pass2_bf16(ctypes.byref(pack.struct), pts.data_ptr(), sb, sc, sp, bs, n,
                                                  gbias.data_ptr(), logits.data_ptr(), mask.data_ptr(), ops._stream()),
                   "seg_pass2_bf16")
    return logits, mask


def trunk_maxpool(pack, fw, x):
    return chain_maxpool(pack, x)


def umma_selftest(a, b, swap=False):
    """a (128,K), b (N,K) float CUDA tensors -> (128,N) fp32 computed by one tcgen05.mma chain."""
    N, K = b.shape
    d = torch.empty((128, N), device=a.device, dtype=torch.float32)
    ak, bk = kp_pack(a), kp_pack(b)
    _lib.check(_lib.lib().al3d_umma_selftest(ak.data_ptr(), bk.data_ptr(), N, K, d.data_ptr(), int(swap), ops._stream()),
               "umma_selftest")
    torch.cuda.synchronize()
    check_abort("umma_selftest_kernel")
    return d


def umma_selftest_ts(a, b):
    """Like umma_selftest but the kernel stages A in TMEM (A-from-TMEM MMA)."""
    N, K = b.shape
    d = torch.empty((128, N), device=a.device, dtype=torch.float32)
    bk = kp_pack(b)
    a = a.contiguous().float()
    _lib.check(_lib.lib().al3d_umma_selftest_ts(a.data_ptr(), bk.data_ptr(), N, K, d.data_ptr(), ops._stream()),
               "umma_selftest_ts")
    torch.cuda.synchronize()
    check_abort("umma_selftest_ts_kernel")
    return d


def umma_selftest_pair(a, b):
    """CTA-pair MMA check: a (256,K), b (N,K) -> (256,N) via cta_group::2 (B split in two N/2-row halves)."""
    N, K = b.shape
    d = torch.empty((256, N), device=a.device, dtype=torch.float32)
    bk = torch.cat([kp_pack(b[: N // 2]), kp_pack(b[N // 2:])]).contiguous()
    a = a.contiguous().float()
    _lib.check(_lib.lib().al3d_umma_selftest_pair(a.data_ptr(), bk.data_ptr(), N, K, d.data_ptr(), ops._stream()),
               "umma_selftest_pair")
    torch.cuda.synchronize()
    check_abort("umma_selftest_pair_kernel")
    return d


def umma_selftest_pair_ss(a, b):
    """CTA-pair MMA with both operands in shared memory: a (256,K), b (128,K) -> (256,128)."""
    K = b.shape[1]
    d = torch.empty((256, 128), device=a.device, dtype=torch.float32)
    ak = torch.cat([kp_pack(a[:128]), kp_pack(a[128:])]).contiguous()
    bk = torch.cat([kp_pack(b[:64]), kp_pack(b[64:])]).contiguous()
    _lib.check(_lib.lib().al3d_umma_selftest_pair_ss(ak.data_ptr(), bk.data_ptr(), K, d.data_ptr(), ops._stream()),
               "umma_selftest_pair_ss")
    torch.cuda.synchronize()
    check_abort("umma_selftest_pair_ss_kernel")
    return d
