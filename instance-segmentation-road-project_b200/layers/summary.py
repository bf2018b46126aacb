"""The first consumer of the pasted masks - mirror of
/root/reference/engine/layers/misc.py: CrackToInstance (:506-543), SummaryOutput (:546-591),
IncludeMyRoad (:594-625), CalculateInstanceSize (:628-724); wired in
/root/reference/road_project/setup/serving.py:45-48.

Floating point: sums accumulate in float64 and round to float32 once, the road-border fit is the
closed-form least squares in float64 (oracle/summary_oracle.py states the contract).
"""
import ctypes

import torch

from .. import runtime as rt
from .base import Layer, ctx_of, i32_scalar, null, register

ROAD_CHANNEL = 1        # seg_outs[..., 1] = my_road   (misc.py:605, :638)
CRACK_CHANNEL = 2       # seg_outs[..., 2] = crack     (misc.py:560)


def _seg_i32(ctx, seg_outs, what):
    if not isinstance(seg_outs, torch.Tensor) or not seg_outs.is_cuda:
        raise rt.InvalidArgumentError(rt.MLP_EDLPACK, f"{what}: seg_outs must be a CUDA tensor (no CPU path)")
    if seg_outs.dim() != 4:
        raise rt.InvalidArgumentError(rt.MLP_EINVAL, f"{what}: seg_outs must be [B,PH,PW,S]")
    return _integral_i32(seg_outs, what)


def _integral_i32(x, what):
    """The kernels take the semantic map as int32 and test `> 0` (road borders, misc.py:661), `> 0.5` after a
    float cast (IncludeMyRoad, misc.py:605-611) and `!= 0` (tf.where in CrackToInstance, misc.py:516).  On the
    integer map UpSampleOutput produces (misc.py:195) the three agree with the reference; on a FRACTIONAL float
    map they would not (0.7 truncates to 0), so such an input is rejected instead of silently summarised."""
    if x.dtype in (torch.int32, torch.int64, torch.int16, torch.int8, torch.uint8, torch.bool):
        return x.to(torch.int32).contiguous()
    if x.is_floating_point():
        if bool((x != x.trunc()).any()):
            raise rt.InvalidArgumentError(
                rt.MLP_EINVAL, f"{what}: the semantic map holds fractional values; pass UpSampleOutput's int32 "
                "{0,1} map (the reference thresholds it with > 0.5 first, engine/layers/misc.py:195)")
        return x.to(torch.int32).contiguous()
    raise rt.InvalidArgumentError(rt.MLP_EINVAL, f"{what}: unsupported semantic map dtype {x.dtype}")


def _masks(ctx, masks, what):
    if not isinstance(masks, torch.Tensor) or not masks.is_cuda:
        raise rt.InvalidArgumentError(rt.MLP_EDLPACK, f"{what}: masks must be a CUDA tensor (no CPU path)")
    if masks.dtype not in (torch.float32, torch.uint8):
        masks = masks.to(torch.float32)                     # tf.cast(crop_ins_outs, tf.float32)
    return masks.contiguous()


def road_scan(ctx, seg, default_road_size, with_crack):
    """seg int32 [B,PH,PW,S] -> (unit [B,PH] f32, road_bits [B,PH,ceil(PW/32)] i32,
    crack_bits (same layout)|None, crack_box [4] i32|None)."""
    B, PH, PW, S = (int(d) for d in seg.shape)
    if S <= ROAD_CHANNEL or (with_crack and S <= CRACK_CHANNEL):
        raise rt.InvalidArgumentError(rt.MLP_EINVAL, f"seg_outs has {S} channels; my_road is 1, crack is 2")
    unit = ctx.empty((B, PH), torch.float32)
    bits = ctx.empty((B, PH, (PW + 31) // 32), torch.int32)
    box = i32_scalar(ctx, 4 + 8 * B) if with_crack else None      # box + per-image crack reductions
    cbits = ctx.empty((B, PH, (PW + 31) // 32), torch.int32) if with_crack else None
    rt.check(ctx.lib.mlp_road_scan(
        ctx.handle, ctx.view(seg), B, PH, PW, S, ROAD_CHANNEL, CRACK_CHANNEL if with_crack else -1,
        float(default_road_size), ctx.view(unit), ctx.view(bits), ctx.view(cbits) if with_crack else null(),
        ctx.view(box) if with_crack else null(), ctx.stream()))
    return unit, bits, cbits, box


def summarize(ctx, det, masks, seg, default_road_size=3.25, threshold=0.1, with_crack=True):
    """det int32 [B,M,6], masks f32|u8 [B,M,PH,PW], seg int32 [B,PH,PW,S] -> [B,M',11] f32."""
    B, M = int(det.shape[0]), int(det.shape[1])
    PH, PW = int(masks.shape[2]), int(masks.shape[3])
    if tuple(seg.shape[:3]) != (B, PH, PW) or tuple(masks.shape[:2]) != (B, M):
        raise rt.InvalidArgumentError(
            rt.MLP_EINVAL, f"shape mismatch: det {tuple(det.shape)}, masks {tuple(masks.shape)}, seg {tuple(seg.shape)}")
    unit, bits, cbits, box = road_scan(ctx, seg, default_road_size, with_crack)
    out = ctx.empty((B * (M + 1) * 11,), torch.float32)
    m_out = i32_scalar(ctx, 1)
    rt.check(ctx.lib.mlp_summary_output(
        ctx.handle, ctx.view(det), ctx.view(masks), rt.MLP_F32 if masks.dtype == torch.float32 else rt.MLP_U8,
        ctx.view(unit), ctx.view(bits), ctx.view(box) if with_crack else null(), B, M, M, null(), PH, PW,
        float(threshold), ctx.view(out), ctx.view(m_out), ctx.stream()))
    Mo = int(m_out.item()) if with_crack else M              # one 4-byte D2H: the dynamic shape
    return out[:B * Mo * 11].view(B, Mo, 11)


def summarize_from_tiles(ctx, det, ins, seg, default_road_size=3.25, threshold=0.1, with_crack=True):
    """The same [B,M',11] summary without the pasted tensor: det int32 [B,M,6] and the int32
    {0,1} tiles ins [B,M,mh,mw] (UpSampleOutput's outputs, i.e. CropAndPadMask's inputs); every
    instance's float32 paste values are evaluated inside its box and reduced on the fly."""
    B, M = int(det.shape[0]), int(det.shape[1])
    mh, mw = int(ins.shape[2]), int(ins.shape[3])
    PH, PW, S = int(seg.shape[1]), int(seg.shape[2]), int(seg.shape[3])
    if int(seg.shape[0]) != B or tuple(ins.shape[:2]) != (B, M):
        raise rt.InvalidArgumentError(
            rt.MLP_EINVAL, f"shape mismatch: det {tuple(det.shape)}, ins {tuple(ins.shape)}, seg {tuple(seg.shape)}")
    unit, bits, cbits, box = road_scan(ctx, seg, default_road_size, with_crack)
    out = ctx.empty((B * (M + 1) * 11,), torch.float32)
    m_out = i32_scalar(ctx, 1)
    rt.check(ctx.lib.mlp_tile_summary(
        ctx.handle, ctx.view(det), ctx.view(ins), null(), 0, null(), 0, null(), B, M, M, null(), mh, mw,
        ctx.view(unit), ctx.view(bits), ctx.view(box) if with_crack else null(), PH, PW, float(threshold),
        ctx.view(out), ctx.view(m_out), null(), ctx.stream()))
    Mo = int(m_out.item()) if with_crack else M
    return out[:B * Mo * 11].view(B, Mo, 11)


@register
class CrackToInstance(Layer):
    """crack int [B,PH,PW] -> (crack_det_outs int32 [B,1,6], crack_seg_outs f32 [B,1,PH,PW]): the
    bounding box of the non-zero pixels of the whole batch as one pseudo-detection per image
    (class id is the literal 5 - the reference ignores crack_id, misc.py:529)."""

    def __init__(self, crack_id=5, **kwargs):
        self.crack_id = crack_id
        super().__init__(**kwargs)

    def call(self, inputs, **kwargs):
        ctx = ctx_of(inputs)
        crack = _integral_i32(inputs, "CrackToInstance")
        B, PH, PW = (int(d) for d in crack.shape)
        # the scan reads a [B,PH,PW,S] map: view the crack plane as S = 1 with both channels = 0
        unit = ctx.empty((B, PH), torch.float32)
        bits = ctx.empty((B, PH, (PW + 31) // 32), torch.int32)
        cbits = torch.empty_like(bits)
        box = i32_scalar(ctx, 4 + 8 * B)
        rt.check(ctx.lib.mlp_road_scan(ctx.handle, ctx.view(crack), B, PH, PW, 1, 0, 0, 3.25, ctx.view(unit),
                                       ctx.view(bits), ctx.view(cbits), ctx.view(box), ctx.stream()))
        y0, x0, y1, x1 = box[:4].tolist()
        if y1 < 0:
            y0 = x0 = y1 = x1 = 0
        h, w = y1 - y0, x1 - x0
        row = [x0 + w // 2, y0 + h // 2, w, h, 5, max(0, min(100, 100 * h * w))]
        det = torch.tensor(row, dtype=torch.int32, device=crack.device).repeat(B, 1, 1)
        return det, crack[:, None].to(torch.float32)

    def get_config(self):
        config = super().get_config()
        config.update({"crack_id": self.crack_id})
        return config


@register
class SummaryOutput(Layer):
    """[det_outs int32 [B,M,6], seg_outs int32 [B,PH,PW,S], crop_ins_outs [B,M,PH,PW]] ->
    float32 [B,M',11] = (class, cx, cy, w, h, conf, pixel_counts, instance_size,
    horizontal_size, vertical_size, include_my_road); M' = M + 1 when the batch holds a crack
    region of positive area (CrackToInstance).  crop_ins_outs may be the float32 tensor of
    CropAndPadMask or its uint8 binary form."""

    def __init__(self, default_road_size=3.25, **kwargs):
        self.default_road_size = default_road_size
        super().__init__(**kwargs)

    def call(self, inputs, **kwargs):
        det_outs, seg_outs, crop_ins_outs = inputs[0], inputs[1], inputs[2]
        ctx = ctx_of(det_outs)
        det = det_outs.to(torch.int32).contiguous()
        return summarize(ctx, det, _masks(ctx, crop_ins_outs, "SummaryOutput"),
                         _seg_i32(ctx, seg_outs, "SummaryOutput"), self.default_road_size)

    def from_tiles(self, inputs):
        """[det_outs int32 [B,M,6], seg_outs, ins_outs int32 [B,M,mh,mw]] -> the same summary, computed
        from CropAndPadMask's INPUTS: the [B,M,PH,PW] tensor is never materialised."""
        det_outs, seg_outs, ins_outs = inputs[0], inputs[1], inputs[2]
        ctx = ctx_of(det_outs)
        return summarize_from_tiles(ctx, det_outs.to(torch.int32).contiguous(),
                                    ins_outs.to(torch.int32).contiguous(),
                                    _seg_i32(ctx, seg_outs, "SummaryOutput"), self.default_road_size)

    def get_config(self):
        config = super().get_config()
        config.update({"default_road_size": self.default_road_size})
        return config


def _zero_det(ctx, masks):
    return torch.zeros((int(masks.shape[0]), int(masks.shape[1]), 6), dtype=torch.int32, device=masks.device)


@register
class IncludeMyRoad(Layer):
    """[seg_outs, crop_ins_outs] -> float32 [B,M]: 1 where more than `threshold` of the instance's
    pixels (> 0.5) lie on my_road."""

    def __init__(self, threshold=0.1, **kwargs):
        self.threshold = threshold
        super().__init__(**kwargs)

    def call(self, inputs, **kwargs):
        seg_outs, crop_ins_outs = inputs[0], inputs[1]
        ctx = ctx_of(crop_ins_outs)
        masks = _masks(ctx, crop_ins_outs, "IncludeMyRoad")
        out = summarize(ctx, _zero_det(ctx, masks), masks, _seg_i32(ctx, seg_outs, "IncludeMyRoad"),
                        threshold=self.threshold, with_crack=False)
        return out[..., 10].contiguous()

    def get_config(self):
        config = super().get_config()
        config.update({"threshold": self.threshold})
        return config


@register
class CalculateInstanceSize(Layer):
    """[seg_outs, pad_ins_outs] -> float32 [B,M,3] = (instance_size, horizontal_size, vertical_size)
    in metres, from the fitted road width on every frame row."""

    def __init__(self, default_road_size=3.25, **kwargs):
        self.default_road_size = default_road_size
        super().__init__(**kwargs)

    def call(self, inputs, **kwargs):
        seg_outs, pad_ins_outs = inputs[0], inputs[1]
        ctx = ctx_of(pad_ins_outs)
        masks = _masks(ctx, pad_ins_outs, "CalculateInstanceSize")
        out = summarize(ctx, _zero_det(ctx, masks), masks, _seg_i32(ctx, seg_outs, "CalculateInstanceSize"),
                        default_road_size=self.default_road_size, with_crack=False)
        return out[..., 7:10].contiguous()

    def get_config(self):
        config = super().get_config()
        config.update({"default_road_size": self.default_road_size})
        return config
