"""Drop-in replacements for the reference's static auto-label models (tools/static_model.py).

Same class names, constructor arguments, attributes (``name``, ``n_classes``, ``n_channel``),
sub-module / ``state_dict`` layout and ``forward(pts, init_box, bbox_gt) -> dict`` contract as
StaticModelOneBoxEst (tools/static_model.py:108-146) and StaticModelTwoBoxEst (:148-239), so the
reference's static_eval.py / static_train.py can import them unchanged.  The
forward runs entirely in libal3d.so (sm_100a CUDA); the nn.Conv1d / nn.BatchNorm1d / nn.Linear
children only hold parameters (identical default initialisation and key names) and are never called.

Extra, non-reference attributes: ``precision`` ("fp32" SIMT | "bf16x3" split-precision tensor cores | "mixed" bf16x3 with conv5 / dconv2 in fp16 | "bf16" fast tensor cores) and ``gather_policy``
("strided" device rule | "numpy_legacy" reference RNG replay).
"""
import numpy as np
import torch
import torch.nn as nn

from . import engine, ops, spec

NUM_HEADING_BIN = spec.NUM_HEADING_BIN
NUM_SIZE_CLUSTER = spec.NUM_SIZE_CLUSTER
NUM_OBJECT_POINT = spec.NUM_OBJECT_POINT
NUM_POINT = spec.NUM_POINT_STATIC
MEAN_SIZE_ARR = np.array(spec.MEAN_SIZE_ARR)


class _ParamBlock(nn.Module):
    """A sub-network that only owns parameters, created from a spec table in the reference's
    registration order (so ``torch.manual_seed(s); Model()`` draws the same initial weights)."""

    def __init__(self, table, dropout_before=None):
        super().__init__()
        self._table = table
        for lname, bn, cin, cout, kind in table:
            if dropout_before == lname:
                self.dropout = nn.Dropout(p=0.5)
            setattr(self, lname, nn.Conv1d(cin, cout, 1) if kind == "conv" else nn.Linear(cin, cout))
        for lname, bn, cin, cout, kind in table:
            if bn is not None:
                setattr(self, bn, nn.BatchNorm1d(cout))

    def forward(self, *a, **k):
        raise RuntimeError("parameter container only; the forward runs in libal3d.so via the parent model")


class PointNetInstanceSeg(_ParamBlock):
    def __init__(self, n_classes=3, n_channel=3):
        super().__init__(spec.seg_layers(n_channel), dropout_before="dconv5")
        self.n_channel = n_channel


class PointNetEstimation(_ParamBlock):
    def __init__(self, n_classes=3):
        super().__init__(spec.static_est_layers())


class _AutoLabelBase(nn.Module):
    precision = engine.DEFAULT_PRECISION
    gather_policy = "strided"

    def _init_common(self):
        self._packs = engine.PackCache()

    def forward(self, *args, **kwargs):
        """Eval mode: the fused inference kernels (no autograd graph).  Training mode (``.train()``): batch-statistics
        BatchNorm, Dropout and outputs connected to autograd Functions whose backward runs in libal3d.so (train.py)."""
        if self.training:
            return self._forward_train(*args, **kwargs)
        with torch.no_grad():
            return self._forward_eval(*args, **kwargs)

    def forward_boxes(self, *args):
        """Eval forward that also returns the decoded (bs,7) boxes (tools/static_eval.py:270-288) from the same fused
        head launch: (output dict, boxes)."""
        with torch.no_grad():
            out = self._forward_eval(*args, want_box=True)
        return out, self._last_box

    def _grad_bucket(self):
        """Flat gradient bucket the training-mode backward writes into (rebuilt when the parameters are re-homed)."""
        from . import train
        key = tuple(p.data_ptr() for p in self.parameters())
        if getattr(self, "_bucket_key", None) != key:
            self._bucket, self._bucket_key = train.GradBucket(self), key
        return self._bucket

    def _seg_train(self, pts):
        from . import train
        bs, _, n = pts.shape
        drop = train.dropout_multiplier(bs, n, self.ins_seg.dropout.p, pts.device)
        return train.seg_apply(self, self.ins_seg, pts, drop)

    def _check_inputs(self, pts, C):
        if not pts.is_cuda:
            raise RuntimeError("the B200 build has no CPU path: inputs must be CUDA tensors")
        if pts.dim() != 3 or pts.shape[1] != C:
            raise ValueError("pts must be (bs,%d,n), got %s" % (C, tuple(pts.shape)))
        if pts.dtype != torch.float32:
            raise TypeError("pts must be float32")

    def _tc_engine(self):
        """Tensor-core engines: "bf16" (fast, csrc/chain_bf16.cu) and "bf16x3" (split precision, parity-grade,
        csrc/chain_split.cu)."""
        if self.precision == "bf16":
            from . import engine_bf16
            return engine_bf16
        if self.precision == "bf16x3":
            from . import engine_split
            return engine_split
        if self.precision == "mixed":
            from . import engine_split
            return engine_split.mixed
        raise ValueError("precision must be 'fp32', 'bf16x3', 'mixed' or 'bf16', got %r" % (self.precision,))

    def _seg(self, pts):
        fw = self._packs.get("seg_f32", self.ins_seg, lambda: engine.fold_block(self.ins_seg, self.ins_seg._table))
        if self.precision == "fp32":
            return engine.seg_forward_fp32(fw, pts), None
        eng = self._tc_engine()
        pk = self._packs.get("seg_" + self.precision, self.ins_seg, lambda: eng.pack_seg(fw, self.ins_seg.n_channel))
        return eng.seg_forward(pk, fw, pts)

    def _fc_t(self, key, module, fw, names):
        """Transposed FC weights of a head for the fused kernel, cached with the other packs."""
        return self._packs.get(key + "_fc_t", module, lambda: engine.fc_layers_t(fw, names))

    def _trunk(self, key, module, x):
        fw = self._packs.get(key + "_f32", module, lambda: engine.fold_block(module, module._table))
        if self.precision == "fp32":
            return fw, engine.trunk_maxpool_fp32(fw, x)
        eng = self._tc_engine()
        pk = self._packs.get(key + "_" + self.precision, module, lambda: eng.pack_trunk(fw))
        return fw, eng.trunk_maxpool(pk, fw, x)


class StaticModelOneBoxEst(_AutoLabelBase):
    def __init__(self, n_classes=3, n_channel=3):
        super().__init__()
        self.name = "one_box_est"
        self.n_classes = n_classes
        self.n_channel = n_channel
        self.ins_seg = PointNetInstanceSeg(n_classes=n_classes, n_channel=n_channel)
        self.box_est = PointNetEstimation(n_classes=n_classes)
        self._init_common()

    def _forward_train(self, pts, init_box, bbox_gt=None):
        """tools/static_model.py:117-146 under model.train() (tools/static_train.py:66,83)."""
        from . import train
        self._check_inputs(pts, self.n_channel)
        logits = self._seg_train(pts)
        with torch.no_grad():          # the gather is index work: not differentiable in the reference either (:33-47)
            obj, mask, _ = engine.mask_and_gather(pts[:, :3, :], logits.detach(), NUM_OBJECT_POINT, self.gather_policy)
        h = train.parse_heads_torch(train.head_apply(self, self.box_est, obj))
        out = {"logits": logits, "mask": mask}
        out.update(h)
        out["center"] = h["center_boxnet"] + init_box.float()[:, :3]
        return out

    def _forward_eval(self, pts, init_box, bbox_gt=None, want_box=False):
        self._check_inputs(pts, self.n_channel)
        logits, seg_mask = self._seg(pts)
        obj, mask, _ = engine.mask_and_gather(pts[:, :3, :], logits, NUM_OBJECT_POINT, self.gather_policy, mask=seg_mask)
        fw, g = self._trunk("box_est", self.box_est, obj)
        init_box = init_box.float()
        out = ops.fc_chain(g, self._fc_t("box_est", self.box_est, fw, ("fc1", "fc2", "fc3")), heads=True, add=init_box,
                           base_heading=init_box[:, 6], want_box=want_box)
        self._last_box = out.get("_box")
        return {
            "logits": logits, "mask": mask, "center_boxnet": out["center_boxnet"],
            "heading_scores": out["heading_scores"],
            "heading_residuals_normalized": out["heading_residuals_normalized"],
            "heading_residuals": out["heading_residuals"], "size_scores": out["size_scores"],
            "size_residuals_normalized": out["size_residuals_normalized"],
            "size_residuals": out["size_residuals"], "center": out["center"],
        }


class StaticModelTwoBoxEst(_AutoLabelBase):
    def __init__(self, n_classes=3, n_channel=3):
        super().__init__()
        self.name = "two_box_est"
        self.n_classes = n_classes
        self.n_channel = n_channel
        self.ins_seg = PointNetInstanceSeg(n_classes=n_classes, n_channel=n_channel)
        self.box_est_one = PointNetEstimation(n_classes=n_classes)
        self.box_est_two = PointNetEstimation(n_classes=n_classes)
        self._init_common()

    def _forward_train(self, pts, init_box, bbox_gt):
        """tools/static_model.py:158-239 under model.train(): the decode of box_one and the re-centering of the gathered
        points are index / label work on detached values, exactly as in the reference (numpy round trip :177-205)."""
        from . import train
        self._check_inputs(pts, self.n_channel)
        init_box = init_box.float().contiguous()
        bbox_gt = bbox_gt.float().contiguous()
        logits = self._seg_train(pts)
        with torch.no_grad():
            obj, mask, _ = engine.mask_and_gather(pts[:, :3, :], logits.detach(), NUM_OBJECT_POINT, self.gather_policy)
        one = train.parse_heads_torch(train.head_apply(self, self.box_est_one, obj))
        center_one = one["center_boxnet"] + init_box[:, :3]
        with torch.no_grad():
            c = lambda t: t.detach().contiguous()
            box_one, _ = ops.decode_boxes(c(center_one), c(one["heading_scores"]), c(one["heading_residuals"]), c(one["size_scores"]),
                                          c(one["size_residuals"]), base_heading=init_box[:, 6])
            obj2, cls2, res2 = ops.twostage_retransform(obj, init_box, box_one, bbox_gt)
        two = train.parse_heads_torch(train.head_apply(self, self.box_est_two, obj2))
        center_two = two["center_boxnet"] + center_one
        out = {"logits": logits, "mask": mask, "center_one": center_one, "box_one": box_one, "center_two": center_two,
               "heading_class_label_two": cls2, "heading_residuals_label_two": res2, "center": center_two}
        for k in ("heading_scores", "heading_residuals_normalized", "heading_residuals", "size_scores", "size_residuals_normalized",
                  "size_residuals"):
            out[k + "_one"], out[k + "_two"] = one[k], two[k]
        for k in ("heading_scores", "heading_residuals", "size_scores", "size_residuals"):
            out[k] = two[k]
        return out

    def _forward_eval(self, pts, init_box, bbox_gt, want_box=False):
        self._check_inputs(pts, self.n_channel)
        init_box = init_box.float().contiguous()
        bbox_gt = bbox_gt.float().contiguous()
        logits, seg_mask = self._seg(pts)
        obj, mask, _ = engine.mask_and_gather(pts[:, :3, :], logits, NUM_OBJECT_POINT, self.gather_policy, mask=seg_mask)
        fw1, g1 = self._trunk("box_est_one", self.box_est_one, obj)
        one = ops.fc_chain(g1, self._fc_t("box_est_one", self.box_est_one, fw1, ("fc1", "fc2", "fc3")), heads=True, add=init_box,
                           base_heading=init_box[:, 6], want_box=True)
        box_one = one["_box"]
        obj2, cls2, res2 = ops.twostage_retransform(obj, init_box, box_one, bbox_gt)
        fw2, g2 = self._trunk("box_est_two", self.box_est_two, obj2)
        two = ops.fc_chain(g2, self._fc_t("box_est_two", self.box_est_two, fw2, ("fc1", "fc2", "fc3")), heads=True, add=one["center"],
                           base_heading=box_one[:, 6], want_box=want_box)
        self._last_box = two.get("_box")
        return {
            "logits": logits, "mask": mask,
            "heading_scores_one": one["heading_scores"],
            "heading_residuals_normalized_one": one["heading_residuals_normalized"],
            "heading_residuals_one": one["heading_residuals"], "size_scores_one": one["size_scores"],
            "size_residuals_normalized_one": one["size_residuals_normalized"],
            "size_residuals_one": one["size_residuals"], "center_one": one["center"], "box_one": box_one,
            "heading_scores_two": two["heading_scores"],
            "heading_residuals_normalized_two": two["heading_residuals_normalized"],
            "heading_residuals_two": two["heading_residuals"], "size_scores_two": two["size_scores"],
            "size_residuals_normalized_two": two["size_residuals_normalized"],
            "size_residuals_two": two["size_residuals"], "center_two": two["center"],
            "heading_class_label_two": cls2, "heading_residuals_label_two": res2,
            "center": two["center"], "heading_scores": two["heading_scores"],
            "heading_residuals": two["heading_residuals"], "size_scores": two["size_scores"],
            "size_residuals": two["size_residuals"],
        }
