"""Drop-in replacement for the reference's dynamic-object model (tools/dynamic_model.py).

DynamicModel keeps the reference's constructor, attributes (``r``, ``s``, ``n_classes``,
``n_channel``), sub-module / ``state_dict`` layout and ``forward(pts, box, bbox_gt) -> dict``
contract (tools/dynamic_model.py:109-155): 4-channel segmentation net, foreground gather of
5*512 points, point embedding, the 101-step box-trajectory encoder and the FC box head.
"""
import numpy as np
import torch
import torch.nn as nn

from . import engine, ops, spec
from .static_model import _AutoLabelBase, _ParamBlock

NUM_HEADING_BIN = spec.NUM_HEADING_BIN
NUM_SIZE_CLUSTER = spec.NUM_SIZE_CLUSTER
NUM_OBJECT_POINT = spec.NUM_OBJECT_POINT
NUM_POINT = spec.NUM_POINT_DYNAMIC
NUM_FRAME = spec.NUM_FRAME
MEAN_SIZE_ARR = np.array(spec.MEAN_SIZE_ARR)


class PointNetInstanceSeg(_ParamBlock):
    def __init__(self, n_classes=3, n_channel=4):
        super().__init__(spec.seg_layers(n_channel), dropout_before="dconv5")
        self.n_channel = n_channel


class PointEmbedding(_ParamBlock):
    def __init__(self, n_classes=3):
        super().__init__(spec.point_emb_layers())


class BoxEmbedding(_ParamBlock):
    def __init__(self, n_classes=3):
        super().__init__(spec.box_emb_layers())


class PointNetEstimation(_ParamBlock):
    def __init__(self, n_classes=3):
        super().__init__(spec.dynamic_est_layers(n_classes))


class DynamicModel(_AutoLabelBase):
    def __init__(self, n_classes=3, n_channel=4):
        super().__init__()
        self.r = 2
        self.s = 50
        self.n_classes = n_classes
        self.n_channel = n_channel
        self.ins_seg = PointNetInstanceSeg(n_classes=n_classes, n_channel=n_channel)
        self.point_emb = PointEmbedding(n_classes=n_classes)
        self.box_emb = BoxEmbedding(n_classes=n_classes)
        self.box_est = PointNetEstimation(n_classes=n_classes)
        self._init_common()

    def _forward_train(self, pts, box, bbox_gt=None):
        """tools/dynamic_model.py:121-155 under model.train() (tools/dynamic_train.py:50,66)."""
        from . import train
        self._check_inputs(pts, self.n_channel)
        if box.dim() != 3 or box.shape[1] != 8:
            raise ValueError("box must be (bs,8,steps), got %s" % (tuple(box.shape),))
        logits = self._seg_train(pts)
        with torch.no_grad():
            obj, mask, _ = engine.mask_and_gather(pts[:, :4, :], logits.detach(), NUM_FRAME * NUM_OBJECT_POINT, self.gather_policy)
        pe = train.head_apply(self, self.point_emb, obj, n_conv=4, fcs=("fc1", "fc2"))
        be = train.head_apply(self, self.box_emb, box.float(), n_conv=4, fcs=("fc1", "fc2"))
        box_pred = train.head_apply(self, self.box_est, torch.cat([pe, be], dim=1), n_conv=0, fcs=("fc1", "fc2", "fc3"), need_dx=True)
        h = train.parse_heads_torch(box_pred)
        out = {"logits": logits, "mask": mask, "center": h["center_boxnet"]}
        for k in ("heading_scores", "heading_residuals_normalized", "heading_residuals", "size_scores", "size_residuals_normalized",
                  "size_residuals"):
            out[k] = h[k]
        return out

    def _forward_eval(self, pts, box, bbox_gt=None, want_box=False):
        self._check_inputs(pts, self.n_channel)
        if box.dim() != 3 or box.shape[1] != 8:
            raise ValueError("box must be (bs,8,steps), got %s" % (tuple(box.shape),))
        logits, seg_mask = self._seg(pts)
        obj, mask, _ = engine.mask_and_gather(pts[:, :4, :], logits, NUM_FRAME * NUM_OBJECT_POINT, self.gather_policy, mask=seg_mask)
        fwp, gp = self._trunk("point_emb", self.point_emb, obj)
        pe = ops.fc_chain(gp, self._fc_t("point_emb", self.point_emb, fwp, ("fc1", "fc2")))
        fwb, gb = self._trunk("box_emb", self.box_emb, box.float())
        be = ops.fc_chain(gb, self._fc_t("box_emb", self.box_emb, fwb, ("fc1", "fc2")))
        fwe = self._packs.get("box_est_f32", self.box_est, lambda: engine.fold_block(self.box_est, self.box_est._table))
        # cat[point_e, box_e] (tools/dynamic_model.py:137) is read in place as two inputs of the fused head
        out = ops.fc_chain(pe, self._fc_t("box_est", self.box_est, fwe, ("fc1", "fc2", "fc3")), x1=be, heads=True, want_box=want_box)
        self._last_box = out.get("_box")
        return {
            "logits": logits, "mask": mask, "center": out["center"], "heading_scores": out["heading_scores"],
            "heading_residuals_normalized": out["heading_residuals_normalized"],
            "heading_residuals": out["heading_residuals"], "size_scores": out["size_scores"],
            "size_residuals_normalized": out["size_residuals_normalized"], "size_residuals": out["size_residuals"],
        }
