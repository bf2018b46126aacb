"""Training-side target assignment - mirror of
/root/reference/engine/layers/detection.py: CalculateIOU (:378-422), AssignBoxes (:589-697),
/root/reference/engine/layers/instance.py: AssignMasks (:296-386) and
/root/reference/engine/metrics.py: DetectionIOUMetric (:109-165).
"""
import torch

from .. import runtime as rt
from .base import Layer, ctx_of, register


def _f32(ctx, x, what):
    return rt.as_device_f32(ctx, x, what).contiguous()


@register
class CalculateIOU(Layer):
    """[aa_boxes [Na,>=4], bb_boxes [Nb,>=4]] (cx,cy,w,h) -> IoU matrix float32 [Na,Nb]."""

    def call(self, inputs, **kwargs):
        ctx = ctx_of(inputs[0])
        aa, bb = _f32(ctx, inputs[0], "CalculateIOU aa_boxes"), _f32(ctx, inputs[1], "CalculateIOU bb_boxes")
        if aa.dim() != 2 or bb.dim() != 2 or aa.shape[1] < 4 or bb.shape[1] < 4:
            raise rt.InvalidArgumentError(rt.MLP_EINVAL, "CalculateIOU: expected [Na,>=4] and [Nb,>=4]")
        out = ctx.empty((int(aa.shape[0]), int(bb.shape[0])), torch.float32)
        if out.numel():
            rt.check(ctx.lib.mlp_calculate_iou(ctx.handle, ctx.view(aa), int(aa.shape[0]), int(aa.shape[1]),
                                               ctx.view(bb), int(bb.shape[0]), int(bb.shape[1]), ctx.view(out),
                                               ctx.stream()))
        return out


@register
class AssignBoxes(Layer):
    """[gt_boxes [B,G,6] (-1 padded), pr_boxes [B,N,4]] -> (cls_true [B,N,C] one-hot, loc_true [B,N,4],
    assign_mask [B,N,1]: 1 background, 0 assigned, -1 ignored)."""

    def __init__(self, num_classes, **kwargs):
        self.num_classes = num_classes
        super().__init__(**kwargs)

    def call(self, inputs, **kwargs):
        ctx = ctx_of(inputs[0])
        gt = _f32(ctx, inputs[0], "AssignBoxes gt_boxes")
        pr = inputs[1]
        if not isinstance(pr, torch.Tensor) or not pr.is_cuda:
            raise rt.InvalidArgumentError(rt.MLP_EDLPACK, "AssignBoxes: pr_boxes must be a CUDA tensor (no CPU path)")
        if pr.dtype not in (torch.float32, torch.int32):
            pr = pr.to(torch.float32)
        pr = pr.contiguous()
        B, G, N = int(gt.shape[0]), int(gt.shape[1]), int(pr.shape[1])
        if tuple(gt.shape) != (B, G, 6) or tuple(pr.shape) != (B, N, 4):
            raise rt.InvalidArgumentError(rt.MLP_EINVAL, f"AssignBoxes: gt {tuple(gt.shape)}, priors {tuple(pr.shape)}")
        C = int(self.num_classes)
        cls_true = ctx.empty((B, N, C), torch.float32)
        loc_true = ctx.empty((B, N, 4), torch.float32)
        mask = ctx.empty((B, N, 1), torch.float32)
        rt.check(ctx.lib.mlp_assign_boxes(
            ctx.handle, ctx.view(gt), ctx.view(pr), rt.MLP_F32 if pr.dtype == torch.float32 else rt.MLP_I32, B, G, N,
            C, ctx.view(cls_true), ctx.view(loc_true), ctx.view(mask), ctx.stream()))
        return cls_true, loc_true, mask

    def get_config(self):
        config = super().get_config()
        config.update({"num_classes": self.num_classes})
        return config


@register
class AssignMasks(Layer):
    """[roi_boxes [B,R,6], roi_masks [B,R,mh,mw,C] (shape only), gt_boxes [B,G,6], gt_masks [B,G,H,W]]
    -> int32 [B,R,mh,mw]: class id where the RoI's best ground truth mask (same class, IoU >=
    match_iou_threshold), cropped to the RoI, is > 0.5; num_classes elsewhere."""

    def __init__(self, match_iou_threshold=0.5, **kwargs):
        self.match_iou_threshold = match_iou_threshold
        kwargs.update({"trainable": False})
        super().__init__(**kwargs)

    def call(self, inputs, **kwargs):
        ctx = ctx_of(inputs[0])
        roi = _f32(ctx, inputs[0], "AssignMasks roi_boxes")
        roi_masks = inputs[1]
        shape = tuple(roi_masks.shape) if hasattr(roi_masks, "shape") else tuple(roi_masks)
        mh, mw, C = (int(d) for d in shape[-3:])              # only the shape of the mask head output is used
        gt = _f32(ctx, inputs[2], "AssignMasks gt_boxes")
        gm = _f32(ctx, inputs[3], "AssignMasks gt_masks")
        B, R, G = int(roi.shape[0]), int(roi.shape[1]), int(gt.shape[1])
        H, W = int(gm.shape[2]), int(gm.shape[3])
        if tuple(gm.shape[:2]) != (B, G) or tuple(roi.shape) != (B, R, 6) or tuple(gt.shape) != (B, G, 6):
            raise rt.InvalidArgumentError(rt.MLP_EINVAL, "AssignMasks: inconsistent shapes")
        out = ctx.empty((B, R, mh, mw), torch.int32)
        rt.check(ctx.lib.mlp_assign_masks(ctx.handle, ctx.view(roi), R, ctx.view(gt), G, ctx.view(gm), B, H, W, mh, mw,
                                          C, float(self.match_iou_threshold), ctx.view(out), ctx.stream()))
        return out

    def get_config(self):
        config = super().get_config()
        config.update({"match_iou_threshold": self.match_iou_threshold})
        return config


@register
class DetectionIOUMetric(Layer):
    """[pred_boxes [B,P,6], gt_boxes [B,G,6]] -> (precision, recall, fmeasure), float32 [B] each."""

    def call(self, inputs, **kwargs):
        ctx = ctx_of(inputs[0])
        pred, gt = _f32(ctx, inputs[0], "DetectionIOUMetric pred_boxes"), _f32(ctx, inputs[1], "DetectionIOUMetric gt_boxes")
        B, P, G = int(pred.shape[0]), int(pred.shape[1]), int(gt.shape[1])
        out = ctx.empty((B, 3), torch.float32)
        rt.check(ctx.lib.mlp_detection_iou_metric(ctx.handle, ctx.view(pred), P, ctx.view(gt), G, B, ctx.view(out),
                                                  ctx.stream()))
        return out[:, 0].contiguous(), out[:, 1].contiguous(), out[:, 2].contiguous()
