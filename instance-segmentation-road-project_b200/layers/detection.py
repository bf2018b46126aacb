"""Detection layers of the hot path — same names, constructor arguments, `call(inputs)`
list conventions and `get_config()` as /root/reference/engine/layers/detection.py:
PriorLayer (:236-306), RestoreBoxes (:309-344), NormalizeBoxes (:347-375),
DetectionProposal (:435-578).  Tensors are CUDA torch tensors; the arithmetic is the
kernels of libmasklab_b200.so reached through the C ABI (no torch math, no CPU path).
"""
import ctypes

import torch

from .. import runtime as rt
from ..prior import PriorBoxes
from .base import Layer, ctx_of, i32_scalar, null, register


@register
class PriorLayer(Layer):
    """images [B,H,W,3] (shape only) -> prior boxes int32 [B,N,4] (cx,cy,w,h).

    padding: 'same' (ResNet-style, ceil(H/stride)) or 'valid' (floor)."""

    def __init__(self, prior, padding="same", **kwargs):
        if isinstance(prior, dict):
            self.prior = PriorBoxes(**prior)
        elif isinstance(prior, PriorBoxes):
            self.prior = prior
        else:
            raise ValueError("prior must be a PriorBoxes instance or its config dict")
        self.padding = padding
        kwargs.update({"trainable": False})                       # detection.py:264-266
        super().__init__(**kwargs)

    def call(self, inputs, **kwargs):
        ctx = ctx_of(inputs)
        batch, height, width = int(inputs.shape[0]), int(inputs.shape[1]), int(inputs.shape[2])
        pc = self.prior.to_c(self.padding)
        n = int(ctx.lib.mlp_prior_count(ctypes.byref(pc), height, width))
        if n < 0:
            rt.check(n)
        out = ctx.empty((batch, n, 4), torch.int32)
        rt.check(ctx.lib.mlp_prior_layer(ctx.handle, ctypes.byref(pc), batch, height, width,
                                         ctx.view(out, torch.int32), ctx.stream()))
        return out

    def get_config(self):
        config = super().get_config()
        config.update({"prior": self.prior.config, "padding": self.padding})
        return config


@register
class RestoreBoxes(Layer):
    """[loc_pred [B,N,4], pr_boxes [B,N,4]] -> restored boxes f32 [B,N,4] (cx,cy,w,h)."""

    def call(self, inputs, **kwargs):
        loc_pred, pr_boxes = inputs[0], inputs[1]
        ctx = ctx_of(loc_pred)
        loc = rt.as_device_f32(ctx, loc_pred, "RestoreBoxes loc_pred")
        if pr_boxes.dtype == torch.int32:
            pr, is_f32 = pr_boxes.contiguous(), 0
        else:
            pr, is_f32 = rt.as_device_f32(ctx, pr_boxes, "RestoreBoxes pr_boxes"), 1
        if tuple(pr.shape) != tuple(loc.shape) or loc.shape[-1] != 4:
            raise rt.InvalidArgumentError(
                rt.MLP_EINVAL, f"RestoreBoxes: shapes {tuple(loc.shape)} vs {tuple(pr.shape)}")
        out = torch.empty_like(loc)
        rt.check(ctx.lib.mlp_restore_boxes(ctx.handle, ctx.view(loc), ctx.view(pr), is_f32,
                                           loc.numel() // 4, ctx.view(out), ctx.stream()))
        return out


@register
class NormalizeBoxes(Layer):
    """boxes [...,>=4] (cx,cy,w,h) -> (y1,x1,y2,x2) divided by `shape`=(H,W) (default ones)."""

    def call(self, inputs, **kwargs):
        boxes = inputs
        ctx = ctx_of(boxes)
        shape = kwargs.get("shape", None)
        if shape is None:
            ih, iw = 1.0, 1.0
        else:
            if isinstance(shape, torch.Tensor):
                shape = shape.tolist()
            ih, iw = float(shape[0]), float(shape[1])
        b = rt.as_device_f32(ctx, boxes, "NormalizeBoxes boxes")
        stride = int(b.shape[-1])
        rows = b.numel() // stride
        out = ctx.empty(tuple(b.shape[:-1]) + (4,), torch.float32)
        rt.check(ctx.lib.mlp_normalize_boxes(ctx.handle, ctx.view(b), rows, stride, ih, iw,
                                             ctx.view(out), ctx.stream()))
        return out


@register
class DetectionProposal(Layer):
    """[cls_pred [B,N,C], boxes [B,N,4] (cx,cy,w,h), images] -> proposed boxes f32 [B,M,6]
    (cx,cy,w,h,class id,confidence), -1 padded, M = max(1, max kept per image).

    Score threshold -> per-(image,class) NMS -> per-image cross-class NMS, exactly the
    reference's order; `images` is accepted and ignored like in the reference.  After a
    call, `last_keep` ([B,M,2] anchor index, class id) and `last_counts` ([B]) hold the
    kept indices on the device."""

    def __init__(self, min_confidence=0.05, nms_iou_threshold=0.4, post_iou_threshold=0.65,
                 nms_max_output_size=1000, max_batch_size=64, **kwargs):
        self.min_confidence = min_confidence
        self.nms_iou_threshold = nms_iou_threshold
        self.post_iou_threshold = post_iou_threshold
        self.nms_max_output_size = nms_max_output_size
        self.max_batch_size = max_batch_size
        self.last_keep = None
        self.last_counts = None
        super().__init__(**kwargs)

    def params_c(self):
        return rt.DetectionParamsC(float(self.min_confidence), float(self.nms_iou_threshold),
                                   float(self.post_iou_threshold), int(self.nms_max_output_size),
                                   # MoldBatch(max_batch_size=None) has no 32-image limit (misc.py:239)
                                   0 if self.max_batch_size is None else 1)

    def call(self, inputs, **kwargs):
        cls_pred, boxes = inputs[0], inputs[1]
        ctx = ctx_of(cls_pred)
        cls = rt.as_device_f32(ctx, cls_pred, "DetectionProposal cls_pred")
        box = rt.as_device_f32(ctx, boxes, "DetectionProposal boxes")
        if cls.dim() != 3 or box.dim() != 3 or box.shape[-1] != 4 or box.shape[:2] != cls.shape[:2]:
            raise rt.InvalidArgumentError(
                rt.MLP_EINVAL,
                f"DetectionProposal: cls_pred {tuple(cls.shape)} / boxes {tuple(box.shape)}")
        B, N, C = (int(v) for v in cls.shape)
        K = int(self.nms_max_output_size)
        det = ctx.empty((B, K, 6), torch.float32)
        keep = ctx.empty((B, K, 2), torch.int32)
        counts = i32_scalar(ctx, B)
        m_dev = i32_scalar(ctx, 1)
        p = self.params_c()
        rt.check(ctx.lib.mlp_detection_proposal(
            ctx.handle, ctx.view(cls), ctx.view(box), B, N, C, ctypes.byref(p), ctx.view(det),
            ctx.view(keep), ctx.view(counts), ctx.view(m_dev), ctx.stream()))
        M = int(m_dev.item())                  # the one D2H the dynamic output shape needs
        self.last_keep = keep[:, :M]
        self.last_counts = counts
        return det[:, :M].contiguous()

    def get_config(self):
        config = super().get_config()
        config.update({
            "min_confidence": self.min_confidence,
            "nms_iou_threshold": self.nms_iou_threshold,
            "post_iou_threshold": self.post_iou_threshold,
            "nms_max_output_size": self.nms_max_output_size,
            "max_batch_size": self.max_batch_size,
        })
        return config
