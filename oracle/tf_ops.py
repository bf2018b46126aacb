"""Restatement of the TensorFlow 1.14/1.15 CPU kernels the hot path calls.

TEST INFRASTRUCTURE (see oracle/__init__.py).  TensorFlow is a third-party
dependency of the reference that is not vendored under /root/reference and
not pinned by any requirements file (API evidence: TF 1.14-1.15, SURVEY.md
§8c).  Each function names the TF kernel it restates and the reference call
site that fixes its arguments.  All arithmetic is float32 with one rounding
per operation (NumPy never fuses multiply-add), in the operation order of the
TF C++ source.
"""
import numpy as np

F32 = np.float32


# --------------------------------------------------------------------------
# tf.image.non_max_suppression  (NonMaxSuppressionV3, score_threshold=-inf)
#   tensorflow/core/kernels/non_max_suppression_op.cc : IOU(), DoNonMaxSuppressionOp
#   call sites: engine/layers/detection.py:507-510 (per class), :542-545 (per image)
# --------------------------------------------------------------------------
def _corners_area(boxes):
    b = np.asarray(boxes, dtype=F32).reshape(-1, 4)
    ymin = np.minimum(b[:, 0], b[:, 2])
    xmin = np.minimum(b[:, 1], b[:, 3])
    ymax = np.maximum(b[:, 0], b[:, 2])
    xmax = np.maximum(b[:, 1], b[:, 3])
    area = (ymax - ymin) * (xmax - xmin)
    return ymin, xmin, ymax, xmax, area


def iou_one_to_many(ymin, xmin, ymax, xmax, area, i, js):
    """IOU(boxes, i, j) for all j in js, float32, TF operation order."""
    iy0 = np.maximum(ymin[i], ymin[js])
    ix0 = np.maximum(xmin[i], xmin[js])
    iy1 = np.minimum(ymax[i], ymax[js])
    ix1 = np.minimum(xmax[i], xmax[js])
    inter = np.maximum(iy1 - iy0, F32(0)) * np.maximum(ix1 - ix0, F32(0))
    with np.errstate(divide="ignore", invalid="ignore"):
        iou = inter / (area[i] + area[js] - inter)
    bad = (area[i] <= 0) | (area[js] <= 0)
    return np.where(bad, F32(0), iou).astype(F32)


def nms_order(scores):
    """Pop order of the candidate queue: descending score, ties -> lower index."""
    s = np.asarray(scores, dtype=F32)
    return np.lexsort((np.arange(s.size), -s.astype(np.float64)))


def non_max_suppression(boxes, scores, max_output_size, iou_threshold, return_margin=False):
    """Greedy NMS; returns selected indices in selection order (int64).

    A candidate is suppressed iff IoU > iou_threshold (strict) with any box
    selected before it.  Vectorised "suppress forward" form; identical result
    to the literal pop-and-check loop (`non_max_suppression_literal`), which
    the tests verify.
    """
    ymin, xmin, ymax, xmax, area = _corners_area(boxes)
    thr = F32(iou_threshold)
    order = nms_order(scores)
    n = order.size
    alive = np.ones(n, dtype=bool)          # indexed by rank in `order`
    selected = []
    margin = np.inf
    for r in range(n):
        if len(selected) >= max_output_size:
            break
        if not alive[r]:
            continue
        i = order[r]
        selected.append(i)
        rest = np.nonzero(alive[r + 1:])[0] + (r + 1)
        if rest.size:
            iou = iou_one_to_many(ymin, xmin, ymax, xmax, area, i, order[rest])
            if return_margin:
                margin = min(margin, float(np.min(np.abs(iou.astype(np.float64) - float(thr)))))
            alive[rest[iou > thr]] = False
    sel = np.asarray(selected, dtype=np.int64)
    return (sel, margin) if return_margin else sel


def non_max_suppression_literal(boxes, scores, max_output_size, iou_threshold):
    """The TF loop as written: pop best, compare against selected (newest first)."""
    b = np.asarray(boxes, dtype=F32).reshape(-1, 4)
    thr = F32(iou_threshold)

    def iou(i, j):
        ymin_i, xmin_i = min(b[i, 0], b[i, 2]), min(b[i, 1], b[i, 3])
        ymax_i, xmax_i = max(b[i, 0], b[i, 2]), max(b[i, 1], b[i, 3])
        ymin_j, xmin_j = min(b[j, 0], b[j, 2]), min(b[j, 1], b[j, 3])
        ymax_j, xmax_j = max(b[j, 0], b[j, 2]), max(b[j, 1], b[j, 3])
        area_i = F32(ymax_i - ymin_i) * F32(xmax_i - xmin_i)
        area_j = F32(ymax_j - ymin_j) * F32(xmax_j - xmin_j)
        if area_i <= 0 or area_j <= 0:
            return F32(0)
        ih = max(F32(min(ymax_i, ymax_j) - max(ymin_i, ymin_j)), F32(0))
        iw = max(F32(min(xmax_i, xmax_j) - max(xmin_i, xmin_j)), F32(0))
        inter = F32(ih * iw)
        return F32(inter / F32(F32(area_i + area_j) - inter))

    selected = []
    for i in nms_order(scores):
        if len(selected) >= max_output_size:
            break
        if not any(iou(i, j) > thr for j in reversed(selected)):
            selected.append(int(i))
    return np.asarray(selected, dtype=np.int64)


# --------------------------------------------------------------------------
# tf.image.crop_and_resize  (bilinear, extrapolation_value=0)
#   tensorflow/core/kernels/crop_and_resize_op.cc : CropAndResize<CPUDevice,float>
#   call site: engine/layers/instance.py:125-126
# --------------------------------------------------------------------------
def crop_and_resize(image, boxes, box_ind, crop_size, extrapolation_value=0.0):
    """image [B,Hf,Wf,D] f32, boxes [n,4] normalised (y1,x1,y2,x2), box_ind [n]."""
    image = np.asarray(image, dtype=F32)
    boxes = np.asarray(boxes, dtype=F32).reshape(-1, 4)
    ch, cw = int(crop_size[0]), int(crop_size[1])
    _, Hf, Wf, D = image.shape
    n = boxes.shape[0]
    out = np.full((n, ch, cw, D), F32(extrapolation_value), dtype=F32)
    hm1, wm1 = F32(Hf - 1), F32(Wf - 1)
    ys = np.arange(ch, dtype=F32)
    xs = np.arange(cw, dtype=F32)
    for k in range(n):
        y1, x1, y2, x2 = boxes[k]
        if ch > 1:
            hs = (y2 - y1) * hm1 / F32(ch - 1)
            in_y = y1 * hm1 + ys * hs
        else:
            in_y = np.full(1, F32(0.5) * (y1 + y2) * hm1, dtype=F32)
        if cw > 1:
            ws = (x2 - x1) * wm1 / F32(cw - 1)
            in_x = x1 * wm1 + xs * ws
        else:
            in_x = np.full(1, F32(0.5) * (x1 + x2) * wm1, dtype=F32)
        # NaN coordinates fail both comparisons in C++ and would index garbage;
        # the oracle treats them as out of range.
        oky = ~((in_y < 0) | (in_y > hm1)) & np.isfinite(in_y)
        okx = ~((in_x < 0) | (in_x > wm1)) & np.isfinite(in_x)
        if not oky.any() or not okx.any():
            continue
        yy, xx = in_y[oky], in_x[okx]
        top = np.floor(yy).astype(np.int64)
        bot = np.ceil(yy).astype(np.int64)
        ly = (yy - np.floor(yy)).astype(F32)[:, None, None]
        left = np.floor(xx).astype(np.int64)
        right = np.ceil(xx).astype(np.int64)
        lx = (xx - np.floor(xx)).astype(F32)[None, :, None]
        img = image[int(box_ind[k])]
        tl = img[top][:, left]
        tr = img[top][:, right]
        bl = img[bot][:, left]
        br = img[bot][:, right]
        t = tl + (tr - tl) * lx
        b = bl + (br - bl) * lx
        val = t + (b - t) * ly
        sub = out[k]
        sub[np.ix_(np.nonzero(oky)[0], np.nonzero(okx)[0])] = val
    return out


# --------------------------------------------------------------------------
# tf.image.resize(..., align_corners=True)  -> legacy ResizeBilinear
#   tensorflow/core/kernels/resize_bilinear_op.cc + image_resizer_state.h
#   (half_pixel_centers=False); call site engine/layers/misc.py:387-391
# --------------------------------------------------------------------------
def resize_interp_weights(out_size, in_size, align_corners=True):
    """compute_interpolation_weights(): (lower, upper, lerp) per output index."""
    if align_corners and out_size > 1:
        scale = F32(in_size - 1) / F32(out_size - 1)
    else:
        scale = F32(in_size) / F32(out_size)
    p = np.arange(out_size, dtype=F32) * scale
    fl = np.floor(p)
    lower = np.maximum(fl.astype(np.int64), 0)
    upper = np.minimum(np.ceil(p).astype(np.int64), in_size - 1)
    lerp = (p - fl).astype(F32)
    return lower, upper, lerp


def resize_bilinear_align_corners(x, out_h, out_w):
    """x [h,w] (any numeric, cast to f32) -> [out_h,out_w] f32."""
    x = np.asarray(x).astype(F32)
    h, w = x.shape
    if out_h <= 0 or out_w <= 0:
        raise ValueError("output dimensions must be positive")   # TF: InvalidArgument
    ylo, yhi, yl = resize_interp_weights(out_h, h)
    xlo, xhi, xl = resize_interp_weights(out_w, w)
    tl = x[ylo][:, xlo]
    tr = x[ylo][:, xhi]
    bl = x[yhi][:, xlo]
    br = x[yhi][:, xhi]
    xl = xl[None, :]
    yl = yl[:, None]
    top = tl + (tr - tl) * xl
    bot = bl + (br - bl) * xl
    return (top + (bot - top) * yl).astype(F32)
