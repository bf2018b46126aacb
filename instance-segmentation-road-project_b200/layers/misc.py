"""Misc layers of the hot path — mirror of /root/reference/engine/layers/misc.py:
DownSampleInput (:133-161), MoldBatch (:213-293), UpSampleOutput (:164-196), CropAndPadMask (:354-401).
"""
import torch

from .. import runtime as rt
from .base import Layer, ctx_of, i32_scalar, null, register


@register
class MoldBatch(Layer):
    """x [K,...] + batch_indices [K] -> [B, max(1, max count), ...], -1 padded, input order
    kept per image.  With max_batch_size not None the reference partitions into exactly 32
    buckets (misc.py:275), so B > 32 is an error; None lifts the limit (misc.py:239-271)."""

    def __init__(self, max_batch_size=None, **kwargs):
        self.max_batch_size = max_batch_size
        super().__init__(**kwargs)

    def call(self, inputs, **kwargs):
        batch_indices = kwargs.get("batch_indices")
        batch_size = int(kwargs.get("batch_size"))
        ctx = ctx_of(inputs)
        if self.max_batch_size is not None and batch_size > rt.MLP_MAX_BATCH:
            raise rt.InvalidArgumentError(
                rt.MLP_EBATCH, "MoldBatch: batch > 32 (tf.dynamic_partition(..,32), misc.py:275)")
        x = inputs
        if x.dtype not in (torch.float32, torch.int32):
            x = x.to(torch.float32)
        x = x.contiguous()
        if not x.is_cuda:
            raise rt.InvalidArgumentError(rt.MLP_EDLPACK, "MoldBatch: tensor is not on a CUDA device")
        bi = batch_indices.to(device=x.device, dtype=torch.int32).contiguous()   # tf.cast(.., int32)
        K = int(x.shape[0])
        row = 1
        for d in x.shape[1:]:
            row *= int(d)
        counts = i32_scalar(ctx, batch_size)
        m_dev = i32_scalar(ctx, 1)
        rt.check(ctx.lib.mlp_mold_batch_plan(ctx.handle, ctx.view(bi), K, batch_size, ctx.view(counts),
                                             ctx.view(m_dev), ctx.stream()))
        M = int(m_dev.item())
        out = ctx.empty((batch_size, M) + tuple(x.shape[1:]), x.dtype)
        if row > 0:
            rt.check(ctx.lib.mlp_mold_batch_run(ctx.handle, ctx.view(x), ctx.view(counts), K, row,
                                                batch_size, 1 if x.dtype == torch.float32 else 0,
                                                ctx.view(m_dev), ctx.view(out), ctx.stream()))
        return out

    def get_config(self):
        config = super().get_config()
        config.update({"max_batch_size": self.max_batch_size})
        return config


def resize_bilinear(ctx, x, out_h, out_w, threshold=False, align_corners=True):
    """tf.compat.v1.image.resize_bilinear(x, (out_h, out_w), align_corners) on a CUDA NHWC tensor;
    threshold=True returns int32 (value > 0.5)."""
    if not isinstance(x, torch.Tensor) or not x.is_cuda:
        raise rt.InvalidArgumentError(rt.MLP_EDLPACK, "resize_bilinear: expected a CUDA tensor (no CPU path)")
    if x.dim() != 4:
        raise rt.InvalidArgumentError(rt.MLP_EINVAL, "resize_bilinear: expected [B,h,w,S]")
    if x.dtype not in (torch.float32, torch.uint8, torch.int32):
        x = x.to(torch.float32)
    x = x.contiguous()
    code = {torch.float32: rt.MLP_F32, torch.uint8: rt.MLP_U8, torch.int32: rt.MLP_I32}[x.dtype]
    B, h, w, S = (int(d) for d in x.shape)
    out = ctx.empty((B, int(out_h), int(out_w), S), torch.int32 if threshold else torch.float32)
    rt.check(ctx.lib.mlp_resize_bilinear(ctx.handle, ctx.view(x), code, B, h, w, S, int(out_h), int(out_w),
                                         (1 if threshold else 0) | (0 if align_corners else 2), ctx.view(out),
                                         ctx.stream()))
    return out


@register
class ResizeLike(Layer):
    """x [B,h,w,S], target=<tensor [B,H,W,..] or (H, W)> -> float32 [B,H,W,S] (bilinear)."""

    def __init__(self, align_corners=True, **kwargs):
        self.align_corners = align_corners
        super().__init__(**kwargs)

    def call(self, inputs, **kwargs):
        target = kwargs.get("target")
        th, tw = (target.shape[1], target.shape[2]) if hasattr(target, "shape") else target
        return resize_bilinear(ctx_of(inputs), inputs, int(th), int(tw), align_corners=bool(self.align_corners))

    def get_config(self):
        config = super().get_config()
        config.update({"align_corners": self.align_corners})
        return config


@register
class DownSampleInput(Layer):
    """frames [B,H,W,C] -> float32 [B,th,tw,C]: resized with the smaller of the two ratios to
    target_size (aspect kept; the output size is truncated to int), bilinear, align_corners=True."""

    def __init__(self, target_size=(540, 960), **kwargs):
        self.target_size = target_size
        super().__init__(**kwargs)

    def call(self, inputs, **kwargs):
        ctx = ctx_of(inputs)
        ih, iw = int(inputs.shape[1]), int(inputs.shape[2])
        f32 = torch.float32
        ratio = torch.minimum(torch.tensor(float(self.target_size[0]), dtype=f32) / torch.tensor(float(ih), dtype=f32),
                              torch.tensor(float(self.target_size[1]), dtype=f32) / torch.tensor(float(iw), dtype=f32))
        th = int((ratio * torch.tensor(float(ih), dtype=f32)).to(torch.int32))      # tf.cast(.., tf.int32)
        tw = int((ratio * torch.tensor(float(iw), dtype=f32)).to(torch.int32))
        return resize_bilinear(ctx, inputs, th, tw)

    def get_config(self):
        config = super().get_config()
        config.update({"target_size": self.target_size})
        return config


@register
class UpSampleOutput(Layer):
    """[roi_box [B,M,6] f32, roi_mask [B,M,mh,mw] f32, semantic [B,hs,ws,S]], target=frames
    -> (int32 boxes, int32 {0,1} masks, int32 {0,1} semantic [B,PH,PW,S]).

    Boxes are scaled to the target frame with the reference's (swapped) ratios
    (misc.py:180-183: cx,w by PH/hs and cy,h by PW/ws) and truncated to int32; masks are
    thresholded at 0.5; the semantic map is resized to the target frame (bilinear,
    align_corners=True) and thresholded at 0.5 (misc.py:190-195).  `semantic` may also be a shape
    tuple (hs, ws) when that branch is handled elsewhere - it is then returned unchanged, as it is
    with semantic=False."""

    def __init__(self, semantic=True, **kwargs):
        self.semantic = semantic
        super().__init__(**kwargs)

    def get_config(self):
        config = super().get_config()
        config.update({"semantic": self.semantic})
        return config

    def call(self, inputs, **kwargs):
        target = kwargs.get("target")
        roi_box, roi_mask, semantic = inputs[0], inputs[1], inputs[2]
        ctx = ctx_of(roi_box)
        box = rt.as_device_f32(ctx, roi_box, "UpSampleOutput roi_box")
        mask = rt.as_device_f32(ctx, roi_mask, "UpSampleOutput roi_mask")
        src_h, src_w = (semantic.shape[1], semantic.shape[2]) if hasattr(semantic, "shape") else semantic
        dst_h, dst_w = (target.shape[1], target.shape[2]) if hasattr(target, "shape") else target
        # float32 division like tf.cast(shape, f32) / tf.cast(shape, f32)
        ratio = (torch.tensor([float(dst_h), float(dst_w)], dtype=torch.float32)
                 / torch.tensor([float(src_h), float(src_w)], dtype=torch.float32))
        box_i = ctx.empty(tuple(box.shape), torch.int32)
        mask_i = ctx.empty(tuple(mask.shape), torch.int32)
        rt.check(ctx.lib.mlp_upsample_output(
            ctx.handle, ctx.view(box), box.numel() // 6, float(ratio[0]), float(ratio[1]),
            ctx.view(box_i), ctx.view(mask), mask.numel(), ctx.view(mask_i), ctx.stream()))
        if self.semantic and isinstance(semantic, torch.Tensor) and semantic.is_cuda:
            semantic = resize_bilinear(ctx, semantic, int(dst_h), int(dst_w), threshold=True)
        return box_i, mask_i, semantic


@register
class CropAndPadMask(Layer):
    """[images [B,PH,PW,3] (shape only), det_outs int32 [B,M,6], ins_outs int32 [B,M,mh,mw], ...]
    -> pasted masks [B,M,PH,PW].

    output='float32' (default) is the reference's tensor: bilinear values of the resized
    mask inside the clipped box, 0 elsewhere.  output='uint8' fuses the consumers'
    `> 0.5` (misc.py:457, :611-615) and writes the binary mask, a quarter of the bytes;
    output='bits' packs that binary mask 8 pixels per byte ([B,M,PH,PW/8], numpy packbits
    bitorder='little')."""

    def __init__(self, output="float32", **kwargs):
        if output not in ("float32", "uint8", "bits"):
            raise ValueError("output must be 'float32', 'uint8' or 'bits'")
        self.output = output
        super().__init__(**kwargs)

    def call(self, inputs, **kwargs):
        images, det_outs, ins_outs = inputs[0], inputs[1], inputs[2]
        ctx = ctx_of(det_outs)
        frame_h, frame_w = (images.shape[1], images.shape[2]) if hasattr(images, "shape") else images
        det = det_outs.to(torch.int32).contiguous()
        ins = ins_outs.to(torch.int32).contiguous()
        B, M = int(det.shape[0]), int(det.shape[1])
        mh, mw = int(ins.shape[2]), int(ins.shape[3])
        mode = {"float32": rt.MLP_PASTE_F32, "uint8": rt.MLP_PASTE_U8, "bits": rt.MLP_PASTE_BITS}[self.output]
        if mode == rt.MLP_PASTE_BITS:
            if int(frame_w) % 8:
                raise rt.InvalidArgumentError(rt.MLP_EINVAL, "bit-packed output needs frame width % 8 == 0")
            out = ctx.empty((B, M, int(frame_h), int(frame_w) // 8), torch.uint8)
        else:
            out = ctx.empty((B, M, int(frame_h), int(frame_w)),
                            torch.uint8 if mode == rt.MLP_PASTE_U8 else torch.float32)
        rt.check(ctx.lib.mlp_crop_and_pad_mask(
            ctx.handle, ctx.view(det), ctx.view(ins), B, M, M, null(), mh, mw, int(frame_h),
            int(frame_w), mode, ctx.view(out), ctx.stream()))
        return out

    def get_config(self):
        config = super().get_config()
        config.update({"output": self.output})
        return config


@register
class EncodeImageContent(Layer):
    """[images uint8 [B,H,W,3]] -> the JPEG file of images[0] as a one-element list of `bytes` (the
    reference returns a string tensor of shape [1]): misc.py:343-351 = tf.io.encode_jpeg(inputs[0]) with
    default attributes (quality 95, 4:2:0, baseline, standard Huffman tables, JFIF 300 dpi), wired in
    road_project/setup/serving.py:41.  The bytes equal libjpeg(-turbo)'s.  `quality` is tf.io.encode_jpeg's
    attribute of that name.  `encode_batch` compresses every frame of the batch in the same five launches and
    leaves the files on the device."""

    def __init__(self, quality=95, **kwargs):
        self.quality = int(quality)
        super().__init__(**kwargs)

    def encode_batch(self, images, out=None, lengths=None):
        """images uint8 [B,H,W,3] (CUDA) -> (files uint8 [B,stride], lengths int32 [B]) on the device; file b is
        files[b, :lengths[b]].  A negative length means the file did not fit `out`'s stride (-bytes needed)."""
        if not isinstance(images, torch.Tensor) or not images.is_cuda:
            raise rt.InvalidArgumentError(rt.MLP_EDLPACK, "EncodeImageContent: images must be a CUDA tensor (no CPU path)")
        if images.dim() != 4 or int(images.shape[3]) != 3:
            raise rt.InvalidArgumentError(rt.MLP_EINVAL, "EncodeImageContent: images must be [B,H,W,3]")
        if images.dtype != torch.uint8:
            raise rt.InvalidArgumentError(rt.MLP_EDLPACK, "EncodeImageContent: images must be uint8 (tf.io.encode_jpeg)")
        ctx = ctx_of(images)
        images = images.contiguous()
        B, H, W = (int(d) for d in images.shape[:3])
        if out is None:
            # room for the largest possible file when that is small, else 3 bytes per pixel (a quality-95 file of
            # pure noise takes about 1.3); a frame that would not fit reports a negative length
            worst = int(ctx.lib.mlp_jpeg_max_bytes(H, W))
            stride = worst if B * worst <= (64 << 20) else min(worst, 1024 + 3 * H * W)
            out = ctx.empty((B, (stride + 15) // 16 * 16), torch.uint8)
        if lengths is None:
            lengths = ctx.empty((B,), torch.int32)
        rt.check(ctx.lib.mlp_jpeg_encode(ctx.handle, ctx.view(images), B, H, W, self.quality, ctx.view(out),
                                         int(out.shape[1]), ctx.view(lengths), ctx.stream()))
        return out, lengths

    def call(self, inputs, **kwargs):
        images = inputs[0] if isinstance(inputs, (list, tuple)) else inputs
        if isinstance(images, torch.Tensor) and images.dim() == 3:
            images = images[None]
        first = images[:1]
        files, lengths = self.encode_batch(first)
        n = int(lengths[0])
        if n < 0:                                           # pathological frame: retry with the hard bound
            ctx = ctx_of(first)
            worst = int(ctx.lib.mlp_jpeg_max_bytes(int(first.shape[1]), int(first.shape[2])))
            files, lengths = self.encode_batch(first, out=ctx.empty((1, (worst + 15) // 16 * 16), torch.uint8))
            n = int(lengths[0])
        return [files[0, :n].cpu().numpy().tobytes()]

    def get_config(self):
        config = super().get_config()
        config.update({"quality": self.quality})
        return config
