"""Semantic post-processing next to the path - mirror of
/root/reference/engine/layers/semantic.py: SemanticSmoothing (:260-293).
"""
import torch

from .. import runtime as rt
from .base import Layer, ctx_of, register


@register
class SemanticSmoothing(Layer):
    """semantic probabilities [B,h,w,S] -> the same shape: grey erosion then dilation with a flat
    kernel_size x kernel_size window (padding SAME), times weight - specks narrower than the window
    disappear, everything else keeps its value."""

    def __init__(self, kernel_size=10, weight=1., **kwargs):
        self.kernel_size = kernel_size
        self.weight = weight
        super().__init__(**kwargs)

    def call(self, inputs, **kwargs):
        ctx = ctx_of(inputs)
        x = rt.as_device_f32(ctx, inputs, "SemanticSmoothing inputs")
        if x.dim() != 4:
            raise rt.InvalidArgumentError(rt.MLP_EINVAL, "SemanticSmoothing: expected [B,h,w,S]")
        B, h, w, S = (int(d) for d in x.shape)
        out = ctx.empty((B, h, w, S), torch.float32)
        rt.check(ctx.lib.mlp_semantic_smoothing(ctx.handle, ctx.view(x), B, h, w, S, int(self.kernel_size),
                                                float(self.weight), ctx.view(out), ctx.stream()))
        return out

    def get_config(self):
        config = super().get_config()
        config.update({"kernel_size": self.kernel_size, "weight": self.weight})
        return config
