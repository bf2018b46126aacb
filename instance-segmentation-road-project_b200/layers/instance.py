"""Instance layers of the hot path — mirror of /root/reference/engine/layers/instance.py:
MaskDistribute (:32-74), PyramidRoiAlign (:77-147), TrimInstances (:250-285).
"""
import ctypes

import torch

from .. import runtime as rt
from .base import Layer, ctx_of, i32_scalar, null, register


@register
class MaskDistribute(Layer):
    """proposed boxes [B,M,6] -> [B,M,7] with the FPN level k prepended
    (k = clip(floor(log2(sqrt(w*h)/base_size)), 0, max_k); -1 on padded rows)."""

    def __init__(self, max_k=2, base_size=64, **kwargs):
        self.max_k = max_k
        self.base_size = base_size
        super().__init__(**kwargs)

    def call(self, inputs, **kwargs):
        ctx = ctx_of(inputs)
        x = rt.as_device_f32(ctx, inputs, "MaskDistribute inputs")
        if x.shape[-1] != 6:
            raise rt.InvalidArgumentError(rt.MLP_EINVAL, f"MaskDistribute: last dim {x.shape[-1]} != 6")
        out = ctx.empty(tuple(x.shape[:-1]) + (7,), torch.float32)
        rt.check(ctx.lib.mlp_mask_distribute(ctx.handle, ctx.view(x), x.numel() // 6, int(self.max_k),
                                             float(self.base_size), ctx.view(out), ctx.stream()))
        return out

    def get_config(self):
        config = super().get_config()
        config.update({"max_k": self.max_k, "base_size": self.base_size})
        return config


@register
class PyramidRoiAlign(Layer):
    """[fmaps (list of [B,Hf,Wf,Cf] NHWC), dist_boxes [B,M,7], images [B,H,W,3]] ->
    [roi_fmaps (list of [B,Mf,ch,cw,Cf]), roi_boxes [B,sum Mf,6]]."""

    def __init__(self, crop_size=(14, 14), max_batch_size=64, **kwargs):
        self.crop_size = crop_size
        self.max_batch_size = max_batch_size
        super().__init__(**kwargs)

    def call(self, inputs, **kwargs):
        fmaps, dist_boxes, images = inputs[0], inputs[1], inputs[2]
        ctx = ctx_of(dist_boxes)
        dist = rt.as_device_f32(ctx, dist_boxes, "PyramidRoiAlign dist_boxes")
        fm = [rt.as_device_f32(ctx, f, "PyramidRoiAlign fmap") for f in fmaps]
        B, M = int(dist.shape[0]), int(dist.shape[1])
        if self.max_batch_size is not None and B > rt.MLP_MAX_BATCH:
            raise rt.InvalidArgumentError(
                rt.MLP_EBATCH, "PyramidRoiAlign: batch > 32 (tf.dynamic_partition(..,32), misc.py:275)")
        L = len(fm)
        Cf = int(fm[0].shape[-1])
        image_h, image_w = float(images.shape[1]), float(images.shape[2])
        ch, cw = int(self.crop_size[0]), int(self.crop_size[1])
        level_counts = i32_scalar(ctx, L * B)
        level_m = i32_scalar(ctx, L + 1)
        rt.check(ctx.lib.mlp_roi_align_plan(ctx.handle, ctx.view(dist), B, M, M, null(), L,
                                            ctx.view(level_counts), ctx.view(level_m), ctx.stream()))
        mf = level_m[:L].tolist()              # D2H of L ints: the dynamic output shapes
        crops = [ctx.empty((B, int(m), ch, cw, Cf), torch.float32) for m in mf]
        roi_boxes = ctx.empty((B, int(sum(mf)), 6), torch.float32)
        fmap_ptrs = (ctypes.c_void_p * L)(*[ctx.view(f).value for f in fm])
        crop_ptrs = (ctypes.c_void_p * L)(*[ctx.view(c).value for c in crops])
        fh = (ctypes.c_int32 * L)(*[int(f.shape[1]) for f in fm])
        fw = (ctypes.c_int32 * L)(*[int(f.shape[2]) for f in fm])
        rt.check(ctx.lib.mlp_roi_align_run(
            ctx.handle, fmap_ptrs, fh, fw, L, Cf, ctx.view(dist), B, M, M, null(), image_h, image_w,
            ch, cw, ctx.view(level_counts), ctx.view(level_m), crop_ptrs, ctx.view(roi_boxes),
            ctx.stream()))
        return [crops, roi_boxes]

    def get_config(self):
        config = super().get_config()
        config.update({"crop_size": self.crop_size, "max_batch_size": self.max_batch_size})
        return config


@register
class TrimInstances(Layer):
    """[roi_boxes [B,R,6], roi_masks [B,R,mh,mw,C]] -> (boxes [B,M,6], masks [B,M,mh,mw]):
    drops MoldBatch padding rows and picks every instance's own class channel."""

    def __init__(self, mold=True, max_batch_size=64, **kwargs):
        self.mold = mold
        self.max_batch_size = max_batch_size
        super().__init__(**kwargs)

    def call(self, inputs, **kwargs):
        roi_boxes, roi_masks = inputs[0], inputs[1]
        ctx = ctx_of(roi_boxes)
        rb = rt.as_device_f32(ctx, roi_boxes, "TrimInstances roi_boxes")
        rm = rt.as_device_f32(ctx, roi_masks, "TrimInstances roi_masks")
        B, R = int(rb.shape[0]), int(rb.shape[1])
        mh, mw, C = int(rm.shape[2]), int(rm.shape[3]), int(rm.shape[4])
        if self.mold and self.max_batch_size is not None and B > rt.MLP_MAX_BATCH:
            raise rt.InvalidArgumentError(
                rt.MLP_EBATCH, "TrimInstances: batch > 32 (tf.dynamic_partition(..,32), misc.py:275)")
        counts = i32_scalar(ctx, B)
        m_dev = i32_scalar(ctx, 1)
        rt.check(ctx.lib.mlp_trim_plan(ctx.handle, ctx.view(rb), B, R, null(), ctx.view(counts),
                                       ctx.view(m_dev), ctx.stream()))
        M = int(m_dev.item())
        out_boxes = ctx.empty((B, M, 6), torch.float32)
        out_masks = ctx.empty((B, M, mh, mw), torch.float32)
        rt.check(ctx.lib.mlp_trim_run(ctx.handle, ctx.view(rb), ctx.view(rm), B, R, null(), mh, mw, C,
                                      ctx.view(m_dev), ctx.view(out_boxes), ctx.view(out_masks),
                                      ctx.stream()))
        if self.mold:
            return out_boxes, out_masks
        # mold=False: flat [K,6] / [K,mh,mw] in row-major (b,j) order (instance.py:276-277);
        # compaction of the molded result is index plumbing only
        valid = (torch.arange(M, device=rb.device)[None, :] < counts[:, None])
        return out_boxes[valid], out_masks[valid]

    def get_config(self):
        config = super().get_config()
        config.update({"mold": self.mold, "max_batch_size": self.max_batch_size})
        return config
