"""Overlay layers of the serving graph - mirror of
/root/reference/engine/layers/misc.py: DrawSegmentation (:404-429) and DrawInstance (:432-475);
wired in /root/reference/road_project/setup/serving.py:34-40.
"""
import ctypes

import torch

from .. import runtime as rt
from .base import Layer, ctx_of, null, register


def _image(ctx, images, what):
    if not isinstance(images, torch.Tensor) or not images.is_cuda:
        raise rt.InvalidArgumentError(rt.MLP_EDLPACK, f"{what}: images must be a CUDA tensor (no CPU path)")
    if images.dim() != 4 or int(images.shape[3]) != 3:
        raise rt.InvalidArgumentError(rt.MLP_EINVAL, f"{what}: images must be [B,PH,PW,3]")
    if images.dtype not in (torch.uint8, torch.float32):
        images = images.to(torch.float32)                  # tf.cast(inputs[0], tf.float32)
    return images.contiguous(), (rt.MLP_U8 if images.dtype == torch.uint8 else rt.MLP_F32)


def _seg(seg_outs, what):
    if not isinstance(seg_outs, torch.Tensor) or not seg_outs.is_cuda:
        raise rt.InvalidArgumentError(rt.MLP_EDLPACK, f"{what}: seg_outs must be a CUDA tensor (no CPU path)")
    if seg_outs.dtype not in (torch.int32, torch.float32):
        seg_outs = seg_outs.to(torch.float32)
    return seg_outs.contiguous(), (rt.MLP_I32 if seg_outs.dtype == torch.int32 else rt.MLP_F32)


@register
class DrawBoxes(Layer):
    """[images [B,PH,PW,3], det_outs int32 [B,M,6]] -> uint8 [B,PH,PW,3]: the one-pixel white rectangle
    of tf.image.draw_bounding_boxes for every detection row (misc.py:481-503)."""

    def call(self, inputs, **kwargs):
        ctx = ctx_of(inputs[1])
        img, it = _image(ctx, inputs[0], "DrawBoxes")
        det = inputs[1].to(torch.int32).contiguous()
        B, PH, PW = (int(d) for d in img.shape[:3])
        out = ctx.empty((B, PH, PW, 3), torch.uint8)
        rt.check(ctx.lib.mlp_draw_boxes(ctx.handle, ctx.view(img), it, ctx.view(det), B, int(det.shape[1]),
                                        int(det.shape[1]), null(), PH, PW, ctx.view(out), ctx.stream()))
        return out


@register
class DrawSegmentation(Layer):
    """[images [B,PH,PW,3], seg_outs [B,PH,PW,C]] -> uint8 [B,PH,PW,3]: every class colour times its
    mask, summed, scaled by alpha, added to the image, clipped to [0,255] and truncated."""

    def __init__(self, colors, alpha=.3, **kwargs):
        self.colors = colors
        self.alpha = alpha
        super().__init__(**kwargs)

    def call(self, inputs, **kwargs):
        ctx = ctx_of(inputs[0])
        img, it = _image(ctx, inputs[0], "DrawSegmentation")
        seg, st = _seg(inputs[1], "DrawSegmentation")
        B, PH, PW = (int(d) for d in img.shape[:3])
        col = rt.DrawColorsC.make(self.colors, self.alpha)
        if tuple(seg.shape) != (B, PH, PW, col.num_classes):
            raise rt.InvalidArgumentError(
                rt.MLP_EINVAL, f"DrawSegmentation: seg_outs {tuple(seg.shape)} != {(B, PH, PW, col.num_classes)}")
        out = ctx.empty((B, PH, PW, 3), torch.uint8)
        rt.check(ctx.lib.mlp_draw_segmentation(ctx.handle, ctx.view(img), it, ctx.view(seg), st, B, PH, PW,
                                               ctypes.byref(col), ctx.view(out), ctx.stream()))
        return out

    def get_config(self):
        config = super().get_config()
        config.update({"colors": self.colors, "alpha": self.alpha})
        return config


@register
class DrawInstance(Layer):
    """[images [B,PH,PW,3], det_outs int32 [B,M,6], crop_and_padded_masks [B,M,PH,PW]] -> uint8
    [B,PH,PW,3]: per class, the pasted masks of its instances are summed and thresholded at 0.5,
    then drawn like DrawSegmentation.  Masks may be CropAndPadMask's float32 tensor or its uint8
    binary form.  `from_tiles` draws the same image from CropAndPadMask's INPUTS."""

    def __init__(self, colors, alpha=.3, **kwargs):
        self.colors = colors
        self.alpha = alpha
        super().__init__(**kwargs)

    def call(self, inputs, **kwargs):
        ctx = ctx_of(inputs[1])
        img, it = _image(ctx, inputs[0], "DrawInstance")
        det = inputs[1].to(torch.int32).contiguous()
        masks = inputs[2]
        if not isinstance(masks, torch.Tensor) or not masks.is_cuda:
            raise rt.InvalidArgumentError(rt.MLP_EDLPACK, "DrawInstance: masks must be a CUDA tensor (no CPU path)")
        if masks.dtype not in (torch.float32, torch.uint8):
            masks = masks.to(torch.float32)
        masks = masks.contiguous()
        B, PH, PW = (int(d) for d in img.shape[:3])
        M = int(det.shape[1])
        if tuple(masks.shape) != (B, M, PH, PW):
            raise rt.InvalidArgumentError(rt.MLP_EINVAL, f"DrawInstance: masks {tuple(masks.shape)} != {(B, M, PH, PW)}")
        col = rt.DrawColorsC.make(self.colors, self.alpha)
        out = ctx.empty((B, PH, PW, 3), torch.uint8)
        rt.check(ctx.lib.mlp_draw_instance(
            ctx.handle, ctx.view(img), it, ctx.view(det), ctx.view(masks),
            rt.MLP_F32 if masks.dtype == torch.float32 else rt.MLP_U8, B, M, M, null(), PH, PW,
            ctypes.byref(col), ctx.view(out), ctx.stream()))
        return out

    def from_tiles(self, inputs, seg_outs=None, semantic_colors=None, semantic_alpha=.3, boxes=False):
        """[images, det_outs int32 [B,M,6], ins_outs int32 [B,M,mh,mw]] -> the same image without the
        [B,M,PH,PW] tensor.  With seg_outs and semantic_colors, DrawSegmentation(semantic_colors,
        semantic_alpha) over the result (serving.py:38-40) is applied in the same pass; with boxes=True
        DrawBoxes()([images, det_outs]) (serving.py:34) goes in front of it, also in the same pass."""
        ctx = ctx_of(inputs[1])
        img, it = _image(ctx, inputs[0], "DrawInstance")
        det = inputs[1].to(torch.int32).contiguous()
        ins = inputs[2].to(torch.int32).contiguous()
        B, PH, PW = (int(d) for d in img.shape[:3])
        M, mh, mw = int(det.shape[1]), int(ins.shape[2]), int(ins.shape[3])
        col = rt.DrawColorsC.make(self.colors, self.alpha)
        seg_ptr, st, sem = null(), rt.MLP_I32, None
        if seg_outs is not None:
            seg, st = _seg(seg_outs, "DrawInstance")
            sem = rt.DrawColorsC.make(semantic_colors, semantic_alpha)
            if tuple(seg.shape) != (B, PH, PW, sem.num_classes):
                raise rt.InvalidArgumentError(rt.MLP_EINVAL, f"DrawInstance: seg_outs {tuple(seg.shape)}")
            seg_ptr = ctx.view(seg)
        out = ctx.empty((B, PH, PW, 3), torch.uint8)
        fn = ctx.lib.mlp_draw_tiles_boxes if boxes else ctx.lib.mlp_draw_tiles
        rt.check(fn(
            ctx.handle, ctx.view(img), it, ctx.view(det), ctx.view(ins), null(), 0, null(), 0, null(), B, M, M,
            null(), mh, mw, PH, PW, ctypes.byref(col), seg_ptr, st,
            ctypes.byref(sem) if sem is not None else None, ctx.view(out), ctx.stream()))
        return out

    def get_config(self):
        config = super().get_config()
        config.update({"colors": self.colors, "alpha": self.alpha})
        return config
