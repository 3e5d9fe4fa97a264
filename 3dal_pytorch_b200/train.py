"""Training mode of the auto-label models (BASELINE.json configs[4]): forward with batch-statistics BatchNorm and
Dropout, backward, fused loss, one flat gradient bucket (single NCCL all-reduce) and a fused Adam step.

Reference: the training loops tools/static_train.py:65-90 / tools/dynamic_train.py:37-133 (``model.train()``,
``output = model(...)``, ``criterion(...)``, ``total_loss.backward()``, ``optimizer.step()`` with Adam lr 1e-3,
weight_decay 1e-4, :220) over the modules of tools/static_model.py:241-339 and tools/dynamic_model.py:157-312.

Every layer is GEMM -> BatchNorm(batch stats) -> ReLU on row-major (M = bs*n, C) activations, all arithmetic in
libal3d.so (csrc/linear_f32.cu, csrc/train.cu).  As in the reference the foreground gather is not differentiable
(tools/static_model.py:33-47 builds the object points from numpy indices), so the box-head gradients stop at the
gathered points and the segmentation net learns from the mask loss only.

Two ways in:
  * ``TrainStep`` -- the fused path: forward, loss, backward, all-reduce of ONE flat gradient bucket, fused Adam; no
    autograd graph at all.
  * the models' own ``forward`` in ``.train()`` mode returns tensors connected to ``torch.autograd.Function``s whose
    backward runs the same kernels, so the reference's ``total_loss.backward(); optimizer.step()`` loop works unchanged.
"""
import ctypes
import os

import torch

from . import _lib, engine, ops, spec

BN_MOMENTUM = 0.1          # nn.BatchNorm1d default
_ws_cache = {}


def _ws(n_floats, dev):
    key = str(dev)
    t = _ws_cache.get(key)
    if t is None or t.numel() < n_floats:
        t = torch.empty((max(int(n_floats), 1 << 20),), device=dev, dtype=torch.float32)
        _ws_cache[key] = t
    return t


def _p(t):
    return None if t is None else t.data_ptr()


# ------------------------------------------------------------------------------------------------ primitives
def bn_forward(y, bn, drop=None, rows_per_group=0, relu=True, training_stats=True):
    """y (M,C) -> z (M,C), mean (C), rstd (C).  Updates bn.running_* / num_batches_tracked like nn.BatchNorm1d."""
    M, C = y.shape
    dev = y.device
    z = torch.empty_like(y)
    mean = torch.empty((C,), device=dev, dtype=torch.float32)
    rstd = torch.empty((C,), device=dev, dtype=torch.float32)
    ws = _ws(_lib.lib().al3d_train_ws_floats(M, C), dev)
    sg, sc, sr = drop.stride() if drop is not None else (0, 0, 0)
    _lib.check(_lib.lib().al3d_bn_train_forward(_p(y), M, C, _p(bn.weight), _p(bn.bias), float(bn.eps), BN_MOMENTUM,
                                                _p(bn.running_mean) if training_stats else None,
                                                _p(bn.running_var) if training_stats else None, _p(drop), sg, sc, sr,
                                                rows_per_group, int(relu), _p(ws), _p(mean), _p(rstd), _p(z), ops._stream()),
               "bn_train_forward")
    if training_stats:
        bn.num_batches_tracked += 1
    return z, mean, rstd


def bn_backward(dz, y, bn, mean, rstd, dgamma, dbeta, drop=None, rows_per_group=0, relu=True):
    """-> dy, written over dz's storage when dz is a temporary (in place); dgamma / dbeta are (C,) output views."""
    M, C = y.shape
    ws = _ws(_lib.lib().al3d_train_ws_floats(M, C), y.device)
    sg, sc, sr = drop.stride() if drop is not None else (0, 0, 0)
    _lib.check(_lib.lib().al3d_bn_train_backward(_p(dz), _p(y), M, C, _p(bn.weight), _p(bn.bias), _p(mean), _p(rstd), _p(drop),
                                                 sg, sc, sr, rows_per_group, int(relu), _p(ws), _p(dgamma), _p(dbeta), _p(dz),
                                                 ops._stream()), "bn_train_backward")
    return dz


def colsum(x, out, rows_per_group=0):
    M, C = x.shape
    ws = _ws(_lib.lib().al3d_train_ws_floats(M, C), x.device)
    _lib.check(_lib.lib().al3d_group_colsum(_p(x), M, C, rows_per_group, _p(ws), _p(out), ops._stream()), "group_colsum")
    return out


def group_max(z, G, n):
    C = z.shape[1]
    g = torch.empty((G, C), device=z.device, dtype=torch.float32)
    arg = torch.empty((G, C), device=z.device, dtype=torch.int32)
    _lib.check(_lib.lib().al3d_group_max_forward(_p(z), G, n, C, _p(g), _p(arg), ops._stream()), "group_max_forward")
    return g, arg


def group_max_backward(dg, arg, G, n):
    C = dg.shape[1]
    dz = torch.zeros((G * n, C), device=dg.device, dtype=torch.float32)
    _lib.check(_lib.lib().al3d_group_max_backward(_p(dg.contiguous()), _p(arg), G, n, C, _p(dz), ops._stream()), "group_max_backward")
    return dz


def wgrad(dy, x, dw, accumulate=False):
    """dw (N,K) view with unit column stride (+)= dy^T x;  dy (M,N), x (M,K) row-major."""
    M, N = dy.shape
    K = x.shape[1]
    assert dw.shape == (N, K) and dw.stride(1) == 1 and dy.stride(1) == 1 and x.stride(1) == 1
    if wgrad_split_fits(M, N, K) and dy.stride(0) % 4 == 0 and x.stride(0) % 4 == 0:
        return wgrad_split(dy, x, dw, accumulate=accumulate)
    ws = _ws(_lib.lib().al3d_wgrad_ws_floats(M, N, K), dy.device)
    _lib.check(_lib.lib().al3d_wgrad_f32(_p(dy), dy.stride(0), _p(x), x.stride(0), M, N, K, _p(ws), _p(dw), dw.stride(0),
                                         int(accumulate), ops._stream()), "wgrad_f32")


def dgrad(dy, w, out=None, accumulate=False):
    """dx (M,K) = dy (M,N) . w (N,K): the NT GEMM on the transposed weight."""
    if gemm_split_fits(dy.shape[0], w.shape[1], w.shape[0]) and dy.stride(0) % 4 == 0:
        return linear_split(dy, w, out=out, accumulate=accumulate, transposed=True)
    wt = w.t().contiguous()
    return ops.linear(dy, wt, None, act=ops.ACT_NONE, out=out, accumulate=accumulate)


# ---- split-precision tensor-core GEMMs (csrc/gemm_split.cu) for the layers that are big enough to feed them
# GEMM_MODE: "x6" (default) = bf16 hi + mid + lo, six MMAs per product: fp32-grade (~1e-7 relative), gradients at the fp32
#                             oracle's own noise level;
#            "x3"           = bf16 hi + lo, three MMAs: ~1e-5 relative; train-mode BatchNorm amplifies that where
#                             |mean| >> std (logits 3e-4, gradients 1e-3..3e-2 of float64) -- faster, opt-in;
#            "f32"          = every GEMM on the fp32 SIMT kernels.
# AL3D_TRAIN_GEMM or set_gemm_mode() selects; tests/test_gpu_train.py runs every training test in all three.
GEMM_MODE = os.environ.get("AL3D_TRAIN_GEMM", "x6")
_PARTS = {"x3": 2, "x6": 3}


def set_gemm_mode(mode):
    global GEMM_MODE
    assert mode in ("f32", "x3", "x6"), mode
    GEMM_MODE = mode


_gemm_ws_cache = {}


def _gemm_ws(nbytes, dev):
    t = _gemm_ws_cache.get(str(dev))
    if t is None or t.numel() < nbytes:
        t = torch.empty((max(int(nbytes), 1 << 21),), device=dev, dtype=torch.uint8)
        _gemm_ws_cache[str(dev)] = t
    return t


def gemm_split_fits(M, N, K):
    return GEMM_MODE in _PARTS and M >= 1024 and K >= 64 and K % 64 == 0 and (N in (64, 128) or (N % 256 == 0 and N <= 4096))


def linear_split(a, w, bias=None, rowbias=None, rows_per_group=0, out=None, accumulate=False, transposed=False, K=None, parts=None):
    """(M,N) (+)= a[:, :K] . B^T (+ bias | per-group row bias) on the tensor cores in split precision.
    w is B (N, K) -- or, with transposed=True, the (K, N) matrix whose transpose is B (dgrad takes the weight as it is)."""
    ops._need_cuda(a, w, bias, rowbias)
    parts = parts or _PARTS[GEMM_MODE]
    M = a.shape[0]
    if transposed:
        Kk, N = w.shape
    else:
        N, Kk = w.shape
    K = Kk if K is None else K
    assert K == Kk and a.stride(1) == 1 and w.stride(1) == 1 and a.shape[1] >= K
    y = out if out is not None else torch.empty((M, N), device=a.device, dtype=torch.float32)
    assert y.shape == (M, N) and y.stride(1) == 1
    ws = _gemm_ws(_lib.lib().al3d_gemm_split_ws_bytes(N, K, parts), a.device)
    _lib.check(_lib.lib().al3d_gemm_split_nt(_p(a), a.stride(0), M, K, _p(w), w.stride(0), int(transposed), _p(bias), _p(rowbias),
                                             int(rows_per_group), N, int(accumulate), _p(y), y.stride(0), parts, _p(ws), ops._stream()),
               "gemm_split_nt")
    return y


def wgrad_split_fits(M, N, K):
    if GEMM_MODE not in _PARTS:
        return False
    kc = 256 if GEMM_MODE == "x3" else 128
    return M >= 1024 and N % 8 == 0 and N >= 64 and K >= 64 and (K % 32 == 0 if K <= kc else K % kc == 0)


def wgrad_split(dy, x, dw, accumulate=False, parts=None):
    """dw (N,K) view with unit column stride (+)= dy^T x on the tensor cores in split precision; dy (M,N), x (M,>=K)."""
    parts = parts or _PARTS[GEMM_MODE]
    M, N = dy.shape
    K = dw.shape[1]
    assert dw.shape == (N, K) and dw.stride(1) == 1 and dy.stride(1) == 1 and x.stride(1) == 1 and x.shape[1] >= K
    ws = _gemm_ws(_lib.lib().al3d_gemm_split_tn_ws_bytes(M, N, K, parts), dy.device)
    _lib.check(_lib.lib().al3d_gemm_split_tn(_p(dy), dy.stride(0), _p(x), x.stride(0), M, N, K, parts, _p(ws), _p(dw), dw.stride(0),
                                             int(accumulate), ops._stream()), "gemm_split_tn")


def _w2(layer):
    return layer.weight.view(layer.weight.shape[0], -1)


# ------------------------------------------------------------------------------------------------ gradient bucket
class GradBucket:
    """One flat fp32 buffer holding the gradient of every parameter of a model (the single bucket of the NCCL
    all-reduce and of the fused Adam step); ``view(p)`` is the slice of parameter ``p`` in its own shape."""

    def __init__(self, model):
        self.params = [p for p in model.parameters()]
        dev = self.params[0].device
        self.offsets, off = {}, 0
        for p in self.params:
            self.offsets[id(p)] = (off, p.numel())
            off += p.numel()
        self.flat = torch.zeros((off,), device=dev, dtype=torch.float32)

    def view(self, p):
        off, n = self.offsets[id(p)]
        return self.flat[off:off + n].view(p.shape)

    def view2(self, layer):
        """(out, in) view of a Conv1d(k=1) / Linear weight gradient."""
        v = self.view(layer.weight)
        return v.view(v.shape[0], -1)

    def zero_(self):
        self.flat.zero_()

    def attach(self):
        """Make every ``p.grad`` a view of the bucket (so torch optimisers and DDP-style code see the gradients)."""
        for p in self.params:
            p.grad = self.view(p)


class FlatParams:
    """Re-homes the parameters of a model as views of one flat buffer (values preserved), so that one kernel can update
    all of them."""

    def __init__(self, model):
        self.params = [p for p in model.parameters()]
        flat = torch.cat([p.detach().reshape(-1).float() for p in self.params])
        off = 0
        for p in self.params:
            n = p.numel()
            p.data = flat[off:off + n].view(p.shape)
            off += n
        self.flat = flat


class FusedAdam:
    """torch.optim.Adam(lr, betas, eps, weight_decay) semantics (tools/static_train.py:220) in one launch over the flat
    parameter / gradient buckets."""

    def __init__(self, model, bucket, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4):
        self.flat_p = FlatParams(model)
        assert [id(p) for p in self.flat_p.params] == [id(p) for p in bucket.params]
        self.bucket = bucket
        self.lr, self.betas, self.eps, self.wd = lr, betas, eps, weight_decay
        self.m = torch.zeros_like(self.flat_p.flat)
        self.v = torch.zeros_like(self.flat_p.flat)
        self.t = 0

    def step(self, grad_scale=1.0):
        self.t += 1
        _lib.check(_lib.lib().al3d_adam_step(_p(self.flat_p.flat), _p(self.bucket.flat), _p(self.m), _p(self.v),
                                             self.flat_p.flat.numel(), self.lr, self.betas[0], self.betas[1], self.eps, self.wd,
                                             self.t, float(grad_scale), ops._stream()), "adam_step")


def allreduce_gradients(bucket, group=None):
    """One all-reduce (sum) of the whole flat bucket; returns the factor the optimiser applies (1 / world size)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return 1.0
    dist.all_reduce(bucket.flat, group=group)
    return 1.0 / dist.get_world_size(group)


# ------------------------------------------------------------------------------------------------ layer blocks
def _layer_fwd(x, lin, bn, tape, key, rowbias=None, rows_per_group=0, K=None, drop=None):
    """Linear (+ per-group bias) -> BatchNorm(batch stats) -> ReLU (-> Dropout multiplier).  bn None: plain linear."""
    W = _w2(lin)
    Wk = W if K is None else W[:, :K]
    if gemm_split_fits(x.shape[0], Wk.shape[0], Wk.shape[1]) and x.stride(0) % 4 == 0:
        y = linear_split(x, Wk, lin.bias if rowbias is None else None, rowbias=rowbias, rows_per_group=rows_per_group, K=Wk.shape[1])
    else:
        y = ops.linear(x, Wk, lin.bias if rowbias is None else None, rowbias=rowbias, rows_per_group=rows_per_group,
                       act=ops.ACT_NONE, K=Wk.shape[1])
    if bn is None:
        tape[key] = (x, None, None, None, None, rows_per_group)
        return y
    z, mean, rstd = bn_forward(y, bn, drop=drop, rows_per_group=rows_per_group)
    tape[key] = (x, y, mean, rstd, drop, rows_per_group)
    return z


def _layer_bwd(dz, lin, bn, tape, key, grads, need_dx=True, K=None, dx_out=None, dx_accumulate=False, bias_grad=True):
    """Backward of _layer_fwd.  Returns dx (or None) and dy (the gradient w.r.t. the pre-BN output: callers that own
    extra inputs of the layer, like dconv1's global feature, use it)."""
    x, y, mean, rstd, drop, rpg = tape.pop(key)
    W = _w2(lin)
    if bn is not None:
        dy = bn_backward(dz, y, bn, mean, rstd, grads.view(bn.weight), grads.view(bn.bias), drop=drop, rows_per_group=rpg)
    else:
        dy = dz
    gw = grads.view2(lin)
    Kx = x.shape[1] if K is None else K
    wgrad(dy, x, gw if K is None else gw[:, :Kx])
    if bias_grad:
        if bn is not None:
            # a bias in front of a BatchNorm: its gradient sum(dy) is exactly zero (the BatchNorm backward removes the batch
            # mean of dy; the reference's autograd returns ~1e-10 rounding noise here) -- written as zero, not reduced
            grads.view(lin.bias).zero_()
        else:
            colsum(dy, grads.view(lin.bias))
    dx = None
    if need_dx:
        dx = dgrad(dy, W if K is None else W[:, :Kx], out=dx_out, accumulate=dx_accumulate)
    return dx, dy


# ------------------------------------------------------------------------------------------------ segmentation net
def seg_forward(seg, pts, drop_mask=None):
    """PointNetInstanceSeg.forward in training mode (tools/static_model.py:271-296).  pts (bs,C,n) any strides;
    drop_mask: None (dropout off) or the (bs,128,n) multiplier nn.Dropout would apply (values 0 or 1/(1-p)).
    Returns logits (bs,n,2) and the tape for seg_backward."""
    bs, C, n = pts.shape
    M = bs * n
    tape = {"shape": (bs, C, n)}
    x0 = pts.transpose(1, 2).reshape(M, C).contiguous()
    z1 = _layer_fwd(x0, seg.conv1, seg.bn1, tape, "conv1")
    z2 = _layer_fwd(z1, seg.conv2, seg.bn2, tape, "conv2")
    z3 = _layer_fwd(z2, seg.conv3, seg.bn3, tape, "conv3")
    z4 = _layer_fwd(z3, seg.conv4, seg.bn4, tape, "conv4")
    z5 = _layer_fwd(z4, seg.conv5, seg.bn5, tape, "conv5")
    g, arg = group_max(z5, bs, n)
    del z5
    tape["pool"] = (g, arg)
    # dconv1 on cat[out2, global repeated]: the 1024-wide half is a per-object bias (+ the conv bias)
    Wd1 = _w2(seg.dconv1)
    gb = ops.linear(g, Wd1[:, 64:], seg.dconv1.bias, act=ops.ACT_NONE, K=1024)
    d1 = _layer_fwd(z2, seg.dconv1, seg.dbn1, tape, "dconv1", rowbias=gb, rows_per_group=n, K=64)
    d2 = _layer_fwd(d1, seg.dconv2, seg.dbn2, tape, "dconv2")
    d3 = _layer_fwd(d2, seg.dconv3, seg.dbn3, tape, "dconv3")
    d4 = _layer_fwd(d3, seg.dconv4, seg.dbn4, tape, "dconv4", rows_per_group=n, drop=drop_mask)
    logits = _layer_fwd(d4, seg.dconv5, None, tape, "dconv5")
    return logits.view(bs, n, 2), tape


def seg_backward(seg, tape, dlogits, grads):
    """dlogits (bs,n,2) -> parameter gradients of the segmentation net written into `grads` (a GradBucket)."""
    bs, C, n = tape["shape"]
    M = bs * n
    dl = dlogits.reshape(M, 2).contiguous()
    d, _ = _layer_bwd(dl, seg.dconv5, None, tape, "dconv5", grads)
    d, _ = _layer_bwd(d, seg.dconv4, seg.dbn4, tape, "dconv4", grads)
    d, _ = _layer_bwd(d, seg.dconv3, seg.dbn3, tape, "dconv3", grads)
    d, _ = _layer_bwd(d, seg.dconv2, seg.dbn2, tape, "dconv2", grads)
    # dconv1: per-point half (K = 64) through the generic block, the global-feature half by hand
    g, arg = tape.pop("pool")
    dz2_a, dy1 = _layer_bwd(d, seg.dconv1, seg.dbn1, tape, "dconv1", grads, K=64)
    Wd1 = _w2(seg.dconv1)
    S = torch.empty((bs, 512), device=dl.device, dtype=torch.float32)
    colsum(dy1, S, rows_per_group=n)                               # per-object sums of dY
    wgrad(S, g, grads.view2(seg.dconv1)[:, 64:])                    # dW[:, 64:] = S^T g
    dg = dgrad(S, Wd1[:, 64:])                                      # (bs,1024)
    del dy1, d
    dz5 = group_max_backward(dg, arg, bs, n)
    d, _ = _layer_bwd(dz5, seg.conv5, seg.bn5, tape, "conv5", grads)
    d, _ = _layer_bwd(d, seg.conv4, seg.bn4, tape, "conv4", grads)
    # out2 has two consumers (conv3 and dconv1): accumulate conv3's input gradient onto dconv1's
    d, _ = _layer_bwd(d, seg.conv3, seg.bn3, tape, "conv3", grads, dx_out=dz2_a, dx_accumulate=True)
    d, _ = _layer_bwd(d, seg.conv2, seg.bn2, tape, "conv2", grads)
    _layer_bwd(d, seg.conv1, seg.bn1, tape, "conv1", grads, need_dx=False)


# ------------------------------------------------------------------------------------------------ trunk + FC heads
def head_forward(mod, x, n_conv=4, fcs=("fc1", "fc2", "fc3"), extra=None):
    """conv1..conv4 (+BN+ReLU) -> max over points -> FC layers (BN+ReLU on all but a layer without BN).
    x (bs,C,m) any strides; with n_conv == 0 x is (bs, K) rows (the dynamic box head)."""
    tape = {}
    if n_conv:
        bs, C, m = x.shape
        h = x.transpose(1, 2).reshape(bs * m, C).contiguous()
        for i in range(1, n_conv + 1):
            h = _layer_fwd(h, getattr(mod, "conv%d" % i), getattr(mod, "bn%d" % i), tape, "conv%d" % i)
        g, arg = group_max(h, bs, m)
        tape["pool"] = (arg, bs, m)
        h = g
    else:
        h = x.contiguous()
    for name in fcs:
        bn = getattr(mod, "fcbn" + name[2:], None)
        h = _layer_fwd(h, getattr(mod, name), bn, tape, name)
    tape["meta"] = (n_conv, fcs)
    return h, tape


def head_backward(mod, tape, dout, grads, need_dx=False):
    n_conv, fcs = tape.pop("meta")
    d = dout.contiguous()
    for i, name in enumerate(reversed(fcs)):
        bn = getattr(mod, "fcbn" + name[2:], None)
        last = (i == len(fcs) - 1)
        d, _ = _layer_bwd(d, getattr(mod, name), bn, tape, name, grads, need_dx=(not last) or n_conv > 0 or need_dx)
    if not n_conv:
        return d
    arg, bs, m = tape.pop("pool")
    d = group_max_backward(d, arg, bs, m)
    for i in range(n_conv, 0, -1):
        d, _ = _layer_bwd(d, getattr(mod, "conv%d" % i), getattr(mod, "bn%d" % i), tape, "conv%d" % i, grads, need_dx=i > 1)
    return None


# ------------------------------------------------------------------------------------------------ loss
LOSS_WEIGHTS = (1.0, 10.0, 1.0, 1.0, 20.0, 20.0)     # mask, centre, heading class, size class, heading / size residual


def loss_forward_backward(logits, box_pred, init_center, labels, w_box=1.0, with_mask=True):
    """Fused loss of one head set (tools/static_model.py:348-425): returns (dict of loss tensors, dlogits or None,
    dbox (bs,39)).  labels = (mask_label, center_label, hcls, hres, scls, sres); centre = box_pred[:, :3] + init_center."""
    from . import losses
    mask_label, center_label, hcls, hres, scls, sres = labels
    dev = box_pred.device
    bs = box_pred.shape[0]
    out = ops.parse_heads(box_pred.contiguous(), add=init_center)
    six = losses._six({"center": out["center"], "heading_scores": out["heading_scores"],
                       "heading_residuals_normalized": out["heading_residuals_normalized"], "size_scores": out["size_scores"],
                       "size_residuals_normalized": out["size_residuals_normalized"]}, "",
                      logits if with_mask else None, mask_label if with_mask else None, center_label, hcls, hres, scls, sres)
    w6 = torch.tensor([LOSS_WEIGHTS[0] if with_mask else 0.0] + [w * w_box for w in LOSS_WEIGHTS[1:]], device=dev, dtype=torch.float32)
    f = lambda t: t.float().contiguous()
    dlogits = torch.empty_like(logits) if with_mask else None
    dbox = torch.empty((bs, spec.HEAD_WIDTH), device=dev, dtype=torch.float32)
    M = logits.shape[0] * logits.shape[1] if with_mask else 0
    _lib.check(_lib.lib().al3d_loss_backward(_p(f(logits)) if with_mask else None, _p(f(mask_label).view(-1)) if with_mask else None, M,
                                             _p(out["center"]), _p(f(center_label)), _p(out["heading_scores"]), _p(hcls.long().contiguous()),
                                             _p(out["heading_residuals_normalized"]), _p(f(hres)), _p(out["size_scores"]),
                                             _p(scls.long().contiguous()), _p(out["size_residuals_normalized"]), _p(f(sres)), bs, _p(w6),
                                             _p(dlogits), _p(dbox), ops._stream()), "loss_backward")
    return six, out, dlogits, dbox


def seg_accuracy_count(logits, mask_label):
    """Number of points whose arg-max class equals the label (tools/static_train.py:128-129), as a 0-dim int64 tensor."""
    cnt = torch.zeros((1,), device=logits.device, dtype=torch.int64)
    M = logits.shape[0] * logits.shape[1]
    _lib.check(_lib.lib().al3d_seg_correct(_p(logits.float().contiguous()), _p(mask_label.float().contiguous().view(-1)), M, _p(cnt),
                                           ops._stream()), "seg_correct")
    return cnt[0]


def dropout_multiplier(bs, n, p, device, generator=None):
    """The multiplier nn.Dropout(p) applies to dconv4's (bs,128,n) output (tools/static_model.py:264,293), drawn with
    torch's own RNG in the reference's layout: 0 or 1/(1-p)."""
    if p <= 0.0:
        return None
    keep = torch.rand((bs, 128, n), device=device, generator=generator) >= p
    return keep.float() / (1.0 - p)


# ------------------------------------------------------------------------------------------------ fused training step
class TrainStep:
    """One full training step of StaticModelOneBoxEst without autograd: forward (train-mode BN, dropout), fused loss,
    backward into ONE flat gradient bucket, NCCL all-reduce of that bucket, fused Adam."""

    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4, dropout_p=0.5, group=None):
        if model.name != "one_box_est":
            raise NotImplementedError("TrainStep is built for StaticModelOneBoxEst; the other models train through autograd")
        self.model = model
        self.grads = GradBucket(model)
        self.opt = FusedAdam(model, self.grads, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        self.dropout_p, self.group = dropout_p, group

    def forward_backward(self, pts, init_box, labels, drop_mask="auto", w_box=1.0):
        m = self.model
        bs, C, n = pts.shape
        if isinstance(drop_mask, str):
            drop_mask = dropout_multiplier(bs, n, self.dropout_p, pts.device)
        self.grads.zero_()
        logits, tape = seg_forward(m.ins_seg, pts, drop_mask)
        obj, mask, _ = engine.mask_and_gather(pts[:, :3, :], logits, spec.NUM_OBJECT_POINT, m.gather_policy)
        box_pred, htape = head_forward(m.box_est, obj)
        six, heads, dlogits, dbox = loss_forward_backward(logits, box_pred, init_box.float().contiguous(), labels, w_box=w_box)
        head_backward(m.box_est, htape, dbox, self.grads)
        seg_backward(m.ins_seg, tape, dlogits, self.grads)
        mk, c, h, s, hr, sr = six.unbind(0)
        total = mk + w_box * (c * 10 + h + s + hr * 20 + sr * 20)
        return {"total_loss": total, "mask_loss": mk, "center_loss": w_box * c * 10, "heading_class_loss": w_box * h,
                "size_class_loss": w_box * s, "heading_residuals_normalized_loss": w_box * hr * 20,
                "size_residuals_normalized_loss": w_box * sr * 20, "logits": logits, "mask": mask, "box_pred": box_pred}

    def step(self, pts, init_box, labels, drop_mask="auto", w_box=1.0):
        out = self.forward_backward(pts, init_box, labels, drop_mask, w_box)
        scale = allreduce_gradients(self.grads, self.group)
        self.opt.step(grad_scale=scale)
        return out


class AutogradTrainStep:
    """The same step for ANY of the three models, through their own ``forward`` in ``.train()`` mode -- the reference's
    loop (tools/static_train.py:65-90, tools/dynamic_train.py:37-133): ``out = model(*inputs)``,
    ``criterion(out, *labels)['total_loss'].backward()`` -- where the backward of the models' autograd Functions writes every
    parameter gradient into the model's flat bucket; then ONE all-reduce of that bucket and the fused Adam step
    (``p.grad`` is not used)."""

    def __init__(self, model, criterion, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4, group=None):
        self.model, self.crit, self.group = model.train(), criterion, group
        self.grads = GradBucket(model)
        self.opt = FusedAdam(model, self.grads, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        # the optimiser re-homed the parameters in one flat buffer: make the Functions' backward write into THIS bucket
        model._bucket, model._bucket_key = self.grads, tuple(p.data_ptr() for p in model.parameters())

    def step(self, inputs, labels):
        self.grads.zero_()
        out = self.model(*inputs)
        ls = self.crit(out, *labels)
        ls["total_loss"].backward()
        for p in self.model.parameters():
            p.grad = None                                  # copies of the bucket's views handed to autograd: not needed
        scale = allreduce_gradients(self.grads, self.group)
        self.opt.step(grad_scale=scale)
        return ls


# ------------------------------------------------------------------------------------------------ autograd wrappers
class _SegFn(torch.autograd.Function):
    """logits = ins_seg(pts) in training mode; backward fills a temporary bucket and returns its views."""

    @staticmethod
    def forward(ctx, owner, seg, pts, drop_mask, *params):
        logits, tape = seg_forward(seg, pts, drop_mask)
        ctx.owner, ctx.seg, ctx.tape = owner, seg, tape
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        grads = ctx.owner._grad_bucket()
        seg_backward(ctx.seg, ctx.tape, dlogits.contiguous(), grads)
        return (None, None, None, None) + tuple(grads.view(p).clone() for p in ctx.seg.parameters())


class _HeadFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, owner, mod, x, n_conv, fcs, need_dx, *params):
        out, tape = head_forward(mod, x, n_conv=n_conv, fcs=fcs)
        ctx.owner, ctx.mod, ctx.tape, ctx.need_dx = owner, mod, tape, need_dx
        return out

    @staticmethod
    def backward(ctx, dout):
        grads = ctx.owner._grad_bucket()
        dx = head_backward(ctx.mod, ctx.tape, dout.contiguous(), grads, need_dx=ctx.need_dx)
        return (None, None, dx if ctx.need_dx else None, None, None, None) + tuple(grads.view(p).clone() for p in ctx.mod.parameters())


def seg_apply(owner, seg, pts, drop_mask):
    return _SegFn.apply(owner, seg, pts, drop_mask, *list(seg.parameters()))


def head_apply(owner, mod, x, n_conv=4, fcs=("fc1", "fc2", "fc3"), need_dx=False):
    return _HeadFn.apply(owner, mod, x, n_conv, fcs, need_dx, *list(mod.parameters()))


def parse_heads_torch(box_pred):
    """parse_output_to_tensors (tools/static_model.py:64-96) as differentiable tensor slicing (training mode only: the
    five views autograd needs; the eval path uses the al3d_parse_heads kernel)."""
    bs = box_pred.shape[0]
    H, S = spec.NUM_HEADING_BIN, spec.NUM_SIZE_CLUSTER
    anchors = torch.tensor(spec.MEAN_SIZE_ARR, device=box_pred.device, dtype=torch.float32)
    center = box_pred[:, :3]
    hs = box_pred[:, 3:3 + H]
    hrn = box_pred[:, 3 + H:3 + 2 * H]
    ss = box_pred[:, 3 + 2 * H:3 + 2 * H + S]
    srn = box_pred[:, 3 + 2 * H + S:].contiguous().view(bs, S, 3)
    return {"center_boxnet": center, "heading_scores": hs, "heading_residuals_normalized": hrn,
            "heading_residuals": hrn * (3.141592653589793 / H), "size_scores": ss, "size_residuals_normalized": srn,
            "size_residuals": srn * anchors.unsqueeze(0)}
