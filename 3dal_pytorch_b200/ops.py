"""Tensor-level wrappers over the C ABI (include/al3d.h).  Every function takes CUDA tensors,
launches on the current PyTorch stream and returns tensors; PyTorch only provides device memory
and streams here -- all arithmetic happens inside libal3d.so."""
import ctypes

import numpy as np
import torch

from . import _lib, spec

ACT_NONE, ACT_RELU = 0, 1
GATHER_STRIDED, GATHER_TABLE = 0, 1


def _p(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("libal3d ops need CUDA tensors; there is no CPU path")


def pointwise_first(x, w, b, act=ACT_RELU):
    """x (bs,C,n) f32 with arbitrary strides -> (bs*n, cout)."""
    _need_cuda(x, w, b)
    bs, C, n = x.shape
    cout = w.shape[0]
    y = torch.empty((bs * n, cout), device=x.device, dtype=torch.float32)
    sb, sc, sp = x.stride()
    _lib.check(_lib.lib().al3d_pointwise_first_f32(_p(x), sb, sc, sp, bs, C, n, _p(w), _p(b), cout, act, _p(y), _stream()),
               "pointwise_first")
    return y


ACT_ACCUMULATE = 2


def linear(a, w, bias=None, rowbias=None, rows_per_group=0, act=ACT_RELU, K=None, max_out=None, out=None, accumulate=False):
    """a (M, >=K) row-major (row stride a.stride(0)), w (cout, >=K) -> (M,cout); with ``max_out``
    (groups,cout) zero-initialised the result is max-pooled into it instead of being stored.  ``out``: write into
    (accumulate=True: add onto) an existing (M,cout) tensor."""
    _need_cuda(a, w, bias, rowbias, max_out)
    M = a.shape[0]
    K = K if K is not None else a.shape[1]
    cout = w.shape[0]
    assert a.stride(1) == 1 and w.stride(1) == 1
    y = None
    if max_out is None:
        y = out if out is not None else torch.empty((M, cout), device=a.device, dtype=torch.float32)
        assert y.shape == (M, cout) and y.is_contiguous()
    if accumulate:
        assert out is not None and max_out is None
        act = act | ACT_ACCUMULATE
    _lib.check(_lib.lib().al3d_linear_f32(_p(a), a.stride(0), M, K, _p(w), w.stride(0), _p(bias), _p(rowbias),
                                          rows_per_group, cout, act, _p(y), cout, _p(max_out), _stream()), "linear")
    return y if max_out is None else max_out


def mask_compact(logits=None, mask=None):
    """logits (bs,n,2) f32 contiguous -> mask (bs,n) bool, pos (bs,n) i32, count (bs,) i32;
    or compact an existing uint8/bool mask."""
    if logits is not None:
        _need_cuda(logits)
        bs, n = logits.shape[0], logits.shape[1]
        assert logits.is_contiguous()
        mask = torch.empty((bs, n), device=logits.device, dtype=torch.bool)
    else:
        _need_cuda(mask)
        bs, n = mask.shape
        assert mask.is_contiguous()
    pos = torch.empty((bs, n), device=mask.device, dtype=torch.int32)
    count = torch.empty((bs,), device=mask.device, dtype=torch.int32)
    _lib.check(_lib.lib().al3d_mask_compact(_p(logits), _p(mask), bs, n, _p(pos), _p(count), _stream()), "mask_compact")
    return mask, pos, count


def gather_fg(x, pos, count, n_pts, choice=None, want_indices=False):
    """x (bs,C,n) any strides -> (bs,C,n_pts) f32 (+ (bs,n_pts) i64 indices)."""
    _need_cuda(x, pos, count, choice)
    bs, C, n = x.shape
    out = torch.empty((bs, C, n_pts), device=x.device, dtype=torch.float32)
    idx = torch.empty((bs, n_pts), device=x.device, dtype=torch.int64) if want_indices else None
    sb, sc, sp = x.stride()
    policy = GATHER_TABLE if choice is not None else GATHER_STRIDED
    _lib.check(_lib.lib().al3d_gather_fg(_p(x), sb, sc, sp, bs, C, n, _p(pos), _p(count), policy, _p(choice), n_pts,
                                         _p(out), _p(idx), _stream()), "gather_fg")
    return (out, idx) if want_indices else out


def parse_heads(box_pred, add=None):
    """(bs,39) -> dict of the 7 head tensors (+ 'center' = centre + add[:, :3] when given)."""
    _need_cuda(box_pred, add)
    bs = box_pred.shape[0]
    dev = box_pred.device
    f = lambda *s: torch.empty(s, device=dev, dtype=torch.float32)
    o = {"center_boxnet": f(bs, 3), "center": f(bs, 3), "heading_scores": f(bs, 12),
         "heading_residuals_normalized": f(bs, 12), "heading_residuals": f(bs, 12), "size_scores": f(bs, 3),
         "size_residuals_normalized": f(bs, 3, 3), "size_residuals": f(bs, 3, 3)}
    assert box_pred.is_contiguous() and box_pred.shape[1] == spec.HEAD_WIDTH
    _lib.check(_lib.lib().al3d_parse_heads(_p(box_pred), bs, _p(add), add.stride(0) if add is not None else 0,
                                           _p(o["center_boxnet"]), _p(o["center"]), _p(o["heading_scores"]),
                                           _p(o["heading_residuals_normalized"]), _p(o["heading_residuals"]),
                                           _p(o["size_scores"]), _p(o["size_residuals_normalized"]),
                                           _p(o["size_residuals"]), _stream()), "parse_heads")
    return o


class FcChainDesc(ctypes.Structure):
    _fields_ = [("x0", ctypes.c_void_p), ("k0", ctypes.c_int32), ("pad0", ctypes.c_int32), ("ld0", ctypes.c_int64),
                ("x1", ctypes.c_void_p), ("k1", ctypes.c_int32), ("pad1", ctypes.c_int32), ("ld1", ctypes.c_int64),
                ("n_layers", ctypes.c_int32), ("width", ctypes.c_int32 * 3), ("relu", ctypes.c_int32 * 3), ("heads", ctypes.c_int32),
                ("wt", ctypes.c_void_p * 3), ("bias", ctypes.c_void_p * 3),
                ("out", ctypes.c_void_p), ("ldo", ctypes.c_int64),
                ("add", ctypes.c_void_p), ("add_stride", ctypes.c_int64),
                ("base_heading", ctypes.c_void_p), ("base_stride", ctypes.c_int64),
                ("center_boxnet", ctypes.c_void_p), ("center", ctypes.c_void_p), ("heading_scores", ctypes.c_void_p),
                ("heading_res_norm", ctypes.c_void_p), ("heading_res", ctypes.c_void_p), ("size_scores", ctypes.c_void_p),
                ("size_res_norm", ctypes.c_void_p), ("size_res", ctypes.c_void_p), ("box", ctypes.c_void_p), ("cls", ctypes.c_void_p)]


def fc_chain(x0, layers, x1=None, heads=False, add=None, base_heading=None, want_box=False):
    """The fused FC chain (csrc/heads.cu): x0 (bs,k0) [+ x1 (bs,k1) concatenated] through up to three layers given as
    (W^T (K,N) contiguous, bias (N), relu) in ONE launch.  heads=False -> (bs, N_last) tensor.  heads=True (39-wide last
    layer) -> dict of the parsed head tensors (centre = centre_boxnet + add[:, :3]) and, with want_box, the decoded
    (bs,7) box (heading base = base_heading) and (bs,2) classes."""
    _need_cuda(x0, x1, add, base_heading)
    bs = x0.shape[0]
    dev = x0.device
    assert x0.stride(1) == 1 and (x1 is None or x1.stride(1) == 1) and 1 <= len(layers) <= 3
    d = FcChainDesc()
    d.x0, d.k0, d.ld0 = x0.data_ptr(), x0.shape[1], x0.stride(0)
    if x1 is not None:
        d.x1, d.k1, d.ld1 = x1.data_ptr(), x1.shape[1], x1.stride(0)
    d.n_layers = len(layers)
    for i, (wt, b, relu) in enumerate(layers):
        assert wt.is_contiguous() and b.is_contiguous()
        d.width[i], d.relu[i], d.wt[i], d.bias[i] = wt.shape[1], int(relu), wt.data_ptr(), b.data_ptr()
    d.heads = int(heads)
    f = lambda *s: torch.empty(s, device=dev, dtype=torch.float32)
    if not heads:
        out = f(bs, layers[-1][0].shape[1])
        d.out, d.ldo = out.data_ptr(), out.stride(0)
        _lib.check(_lib.lib().al3d_fc_chain(ctypes.byref(d), bs, _stream()), "fc_chain")
        return out
    o = {"center_boxnet": f(bs, 3), "center": f(bs, 3), "heading_scores": f(bs, 12),
         "heading_residuals_normalized": f(bs, 12), "heading_residuals": f(bs, 12), "size_scores": f(bs, 3),
         "size_residuals_normalized": f(bs, 3, 3), "size_residuals": f(bs, 3, 3)}
    d.center_boxnet, d.center, d.heading_scores = o["center_boxnet"].data_ptr(), o["center"].data_ptr(), o["heading_scores"].data_ptr()
    d.heading_res_norm, d.heading_res = o["heading_residuals_normalized"].data_ptr(), o["heading_residuals"].data_ptr()
    d.size_scores, d.size_res_norm, d.size_res = o["size_scores"].data_ptr(), o["size_residuals_normalized"].data_ptr(), o["size_residuals"].data_ptr()
    if add is not None:
        d.add, d.add_stride = add.data_ptr(), add.stride(0)
    if want_box:
        o["_box"] = f(bs, 7)
        o["_cls"] = torch.empty((bs, 2), device=dev, dtype=torch.int32)
        d.box, d.cls = o["_box"].data_ptr(), o["_cls"].data_ptr()
        if base_heading is not None:
            d.base_heading, d.base_stride = base_heading.data_ptr(), base_heading.stride(0)
    _lib.check(_lib.lib().al3d_fc_chain(ctypes.byref(d), bs, _stream()), "fc_chain")
    return o


def decode_boxes(center, heading_scores, heading_residuals, size_scores, size_residuals, base_heading=None):
    """-> (bs,7) f32 boxes [centre, l, w, h, heading] and (bs,2) i32 [heading bin, size cluster].
    base_heading: 1-D strided view (e.g. init_box[:, 6])."""
    _need_cuda(center, heading_scores, heading_residuals, size_scores, size_residuals, base_heading)
    bs = center.shape[0]
    box = torch.empty((bs, 7), device=center.device, dtype=torch.float32)
    cls = torch.empty((bs, 2), device=center.device, dtype=torch.int32)
    for t in (center, heading_scores, heading_residuals, size_scores, size_residuals):
        assert t.is_contiguous()
    _lib.check(_lib.lib().al3d_decode_boxes(_p(center), _p(heading_scores), _p(heading_residuals), _p(size_scores),
                                            _p(size_residuals), _p(base_heading),
                                            base_heading.stride(0) if base_heading is not None else 0, bs, _p(box),
                                            _p(cls), _stream()), "decode_boxes")
    return box, cls


def twostage_retransform(obj_pts, init_box, box_one, bbox_gt):
    """obj_pts (bs,3,m) -> re-centred points, heading class label (i64), residual label (f32)."""
    _need_cuda(obj_pts, init_box, box_one, bbox_gt)
    bs, _, m = obj_pts.shape
    for t in (obj_pts, init_box, box_one, bbox_gt):
        assert t.is_contiguous()
    out = torch.empty_like(obj_pts)
    cls = torch.empty((bs,), device=obj_pts.device, dtype=torch.int64)
    res = torch.empty((bs,), device=obj_pts.device, dtype=torch.float32)
    _lib.check(_lib.lib().al3d_twostage_retransform(_p(obj_pts), bs, m, _p(init_box), _p(box_one), _p(bbox_gt), _p(out),
                                                    _p(cls), _p(res), _stream()), "twostage_retransform")
    return out, cls, res
