"""NumPy restatement of the overlay layers of the serving graph:
DrawSegmentation and DrawInstance.  TEST INFRASTRUCTURE — see oracle/__init__.py ("parity unpinned").

Reference: /root/reference/engine/layers/misc.py:404-429 (DrawSegmentation), :432-475 (DrawInstance),
wired in /root/reference/road_project/setup/serving.py:34-40.  float32, one rounding per operation.
The per-class sum of DrawInstance (tf.reduce_sum over the gathered instances) is taken in
instance order; adding the zeros of non-overlapping instances is exact, so only pixels covered by
several instances of one class depend on that order at all.
"""
import numpy as np

F32 = np.float32


def draw_segmentation(images, seg_outs, colors, alpha=0.3):
    """misc.py:412-421: images [B,PH,PW,3] (any dtype), seg_outs [B,PH,PW,C], colors [C,3] ->
    uint8 [B,PH,PW,3] = cast(clip(images + (sum_c colors[c] * seg[..., c]) * alpha, 0, 255))."""
    images = np.asarray(images).astype(F32)
    seg = np.asarray(seg_outs).astype(F32)
    colors = np.asarray(colors, dtype=F32)
    color_seg = np.zeros(images.shape, dtype=F32)
    for c in range(colors.shape[0]):
        color_seg = color_seg + colors[c][None, None, None, :] * seg[..., c][..., None]
    vis = np.clip(images + color_seg * F32(alpha), F32(0), F32(255)).astype(F32)
    return vis.astype(np.uint8)                      # tf.cast truncates


def class_masks(det_outs, masks, num_classes):
    """misc.py:446-459: per image and class, (sum of the masks of the instances whose class
    column equals the class id) > 0.5 -> float32 [B,PH,PW,C]."""
    det_outs = np.asarray(det_outs)
    masks = np.asarray(masks).astype(F32)
    B, M, PH, PW = masks.shape
    out = np.zeros((B, PH, PW, num_classes), dtype=F32)
    for b in range(B):
        for c in range(num_classes):
            s = np.zeros((PH, PW), dtype=F32)
            for j in range(M):
                if det_outs[b, j, 4] == c:
                    s = s + masks[b, j]
            out[b, :, :, c] = (s > F32(0.5)).astype(F32)
    return out


def draw_instance(images, det_outs, masks, colors, alpha=0.3):
    """misc.py:440-463: DrawSegmentation over the per-class instance masks."""
    cm = class_masks(det_outs, masks, len(colors))
    return draw_segmentation(images, cm, colors, alpha)
