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


def draw_bounding_boxes(images, boxes, color=(255.0, 255.0, 255.0, 255.0)):
    """tf.image.draw_bounding_boxes with one colour (TensorFlow core/kernels/draw_bounding_box_op.cc,
    restated): images float [B,H,W,D], boxes [B,n,4] normalised (ymin,xmin,ymax,xmax).  Corner
    coordinates are box * (size - 1) converted to integers by C++ truncation (toward zero); boxes
    with min > max or entirely outside are skipped; each of the four one-pixel lines is drawn
    only if its own coordinate lies inside the image."""
    out = np.array(images, dtype=F32, copy=True)
    B, H, W, D = out.shape
    boxes = np.asarray(boxes, dtype=F32)
    col = np.asarray(color, dtype=F32)[:D]
    for b in range(B):
        for bb in range(boxes.shape[1]):
            y0 = int(np.trunc(boxes[b, bb, 0] * F32(H - 1)))
            x0 = int(np.trunc(boxes[b, bb, 1] * F32(W - 1)))
            y1 = int(np.trunc(boxes[b, bb, 2] * F32(H - 1)))
            x1 = int(np.trunc(boxes[b, bb, 3] * F32(W - 1)))
            y0c, y1c, x0c, x1c = max(y0, 0), min(y1, H - 1), max(x0, 0), min(x1, W - 1)
            if y0 > y1 or x0 > x1:
                continue
            if y0 >= H or y1 < 0 or x0 >= W or x1 < 0:
                continue
            if y0 >= 0:
                out[b, y0, x0c:x1c + 1] = col
            if y1 < H:
                out[b, y1, x0c:x1c + 1] = col
            if x0 >= 0:
                out[b, y0c:y1c + 1, x0] = col
            if x1 < W:
                out[b, y0c:y1c + 1, x1] = col
    return out


def draw_boxes(images, det_outs):
    """DrawBoxes.call, misc.py:481-503: boxes = max(det[..., :4], 0) as (cx,cy,w,h) -> normalised
    corners -> white one-pixel rectangles on the float image, clipped to [0,255], uint8."""
    images = np.asarray(images)
    B, H, W, _ = images.shape
    d = np.maximum(np.asarray(det_outs)[..., :4], 0).astype(F32)
    cx, cy, w, h = d[..., 0], d[..., 1], d[..., 2], d[..., 3]
    fh, fw = F32(H), F32(W)
    xmin, xmax = (cx - w / F32(2)) / fw, (cx + w / F32(2)) / fw
    ymin, ymax = (cy - h / F32(2)) / fh, (cy + h / F32(2)) / fh
    bboxes = np.stack([ymin, xmin, ymax, xmax], axis=-1).astype(F32)
    vis = draw_bounding_boxes(images.astype(F32), bboxes)
    return np.clip(vis, F32(0), F32(255)).astype(np.uint8)
