"""NumPy restatement of the reference's post-backbone layers (rows a1-a14 of
SURVEY.md §8).  TEST INFRASTRUCTURE — see oracle/__init__.py ("parity
unpinned": the reference has no golden vectors for this path).

Every function cites the reference file:line it follows (paths relative to
/root/reference).  float32 throughout, one rounding per operation, operation
order as in the reference source.
"""
import numpy as np

from . import tf_ops

F32 = np.float32
K_EPSILON = F32(1e-7)            # tf.keras.backend.epsilon()


# ----------------------------------------------------------------- a1 -----
def prior_table(strides, sizes, pr_scales, pr_ratios):
    """engine/prior.py:55-67  PriorBoxes.setup -> rows (stride, w, h) int64."""
    assert len(strides) == len(sizes)            # prior.py:40
    rows = []
    for size, stride in zip(sizes, strides):
        for wh_size in pr_scales:
            for wh_ratio in pr_ratios:
                w = int(np.round(size * wh_size * np.sqrt(wh_ratio)))
                h = int(np.round(size * wh_size / np.sqrt(wh_ratio)))
                rows.append((int(stride), w, h))
    return np.asarray(rows, dtype=np.int64).reshape(-1, 3)


def default_prior_config(strides=(8, 16, 32, 64, 128),
                         pr_scales=(2 ** 0, 2 ** (1 / 3), 2 ** (2 / 3)),
                         pr_ratios=(1 / 3, 1 / 2, 1, 2, 3)):
    """engine/retinamasklab.py:46-53 (sizes = 4*stride) + engine/config.py:60-61."""
    strides = [int(s) for s in strides]
    return dict(strides=strides, sizes=[4 * s for s in strides],
                pr_scales=list(pr_scales), pr_ratios=list(pr_ratios))


# ----------------------------------------------------------------- a2 -----
def prior_layer(table, height, width, padding="same"):
    """engine/layers/detection.py:269-298  PriorLayer.call (one image) ->
    [N,4] int32 (cx,cy,w,h), order (stride asc, y, x, anchor)."""
    table = np.asarray(table, dtype=np.int64)
    out = []
    for stride in sorted(set(table[:, 0].tolist())):       # groupby('stride') sorts keys
        rows = table[table[:, 0] == stride]
        if padding == "same":
            th = int(np.ceil(height / stride)) * stride
            tw = int(np.ceil(width / stride)) * stride
        else:
            th = int(np.floor(height / stride)) * stride
            tw = int(np.floor(width / stride)) * stride
        ys = np.arange(stride // 2, th, stride)
        xs = np.arange(stride // 2, tw, stride)
        xg, yg = np.meshgrid(xs, ys)
        per_anchor = [np.stack((xg, yg, np.ones_like(xg) * bw, np.ones_like(yg) * bh), axis=-1)
                      for _, bw, bh in rows]
        out.append(np.stack(per_anchor, axis=2).reshape(-1, 4))
    return np.concatenate(out, axis=0).astype(np.int32)


# ----------------------------------------------------------------- a3 -----
def exp_f32(x):
    """Correctly rounded float32 exp (documented deviation, oracle/__init__.py)."""
    return np.exp(np.asarray(x, dtype=F32).astype(np.float64)).astype(F32)


def log_f32(x):
    """Correctly rounded float32 natural log."""
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.log(np.asarray(x, dtype=F32).astype(np.float64)).astype(F32)


def restore_boxes(loc_pred, pr_boxes):
    """engine/layers/detection.py:325-344  RestoreBoxes.call -> [...,4] f32 cxcywh."""
    loc = np.asarray(loc_pred).astype(F32)
    pr = np.asarray(pr_boxes).astype(F32)
    cx = loc[..., 0] * pr[..., 2] + pr[..., 0]
    cy = loc[..., 1] * pr[..., 3] + pr[..., 1]
    w = exp_f32(loc[..., 2]) * pr[..., 2]
    h = exp_f32(loc[..., 3]) * pr[..., 3]
    return np.stack([cx, cy, w, h], axis=-1).astype(F32)


# ----------------------------------------------------------------- a4 -----
def normalize_boxes(boxes, shape=(1.0, 1.0)):
    """engine/layers/detection.py:360-375  NormalizeBoxes.call -> (y1,x1,y2,x2)."""
    b = np.asarray(boxes, dtype=F32)
    ih, iw = F32(shape[0]), F32(shape[1])
    cx, cy, w, h = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    x1 = (cx - w / F32(2)) / iw
    y1 = (cy - h / F32(2)) / ih
    x2 = (cx + w / F32(2)) / iw
    y2 = (cy + h / F32(2)) / ih
    return np.stack([y1, x1, y2, x2], axis=-1).astype(F32)


# ----------------------------------------------------------------- a8 -----
def mold_batch(x, batch_indices, batch_size, max_batch_size=64):
    """engine/layers/misc.py:231-286  MoldBatch.call.

    Rows of `x` are grouped by image id keeping input order, each image padded
    with -1 to max(1, max count).  With max_batch_size not None the reference
    uses tf.dynamic_partition(..., 32) (misc.py:275): ids >= 32 are an
    InvalidArgument error, modelled here as ValueError.
    """
    x = np.asarray(x)
    bi = np.asarray(batch_indices).astype(np.int64).reshape(-1)
    B = int(batch_size)
    if max_batch_size is not None and (B > 32 or (bi.size and bi.max() >= 32)):
        raise ValueError("MoldBatch: tf.dynamic_partition has 32 partitions (misc.py:275)")
    counts = np.bincount(bi, minlength=B) if bi.size else np.zeros(B, dtype=np.int64)
    M = max(1, int(counts.max()) if counts.size else 0)
    out = np.full((B, M) + x.shape[1:], -1, dtype=x.dtype)
    for b in range(B):
        rows = x[bi == b]
        out[b, :rows.shape[0]] = rows
    return out


# ------------------------------------------------------------ a5-a8 -------
def detection_proposal(cls_pred, boxes, min_confidence=0.05, nms_iou_threshold=0.4,
                       post_iou_threshold=0.65, nms_max_output_size=1000,
                       max_batch_size=64, return_debug=False):
    """engine/layers/detection.py:482-567  DetectionProposal.call.

    cls_pred [B,N,C] f32, boxes [B,N,4] f32 (cx,cy,w,h) -> [B,M,6] f32
    (cx,cy,w,h,class,score) padded with -1, M = max(1, max_b kept_b).
    """
    cls_pred = np.asarray(cls_pred, dtype=F32)
    boxes = np.asarray(boxes, dtype=F32)
    B, N, C = cls_pred.shape
    norm = normalize_boxes(boxes)                                   # :488 (shape = ones)
    keep = np.argwhere(cls_pred >= F32(min_confidence))             # :491 row-major (b,n,c)
    k_img, k_cls = keep[:, 0], keep[:, 2]
    k_conf = cls_pred[keep[:, 0], keep[:, 1], keep[:, 2]]           # :494
    k_box = norm[keep[:, 0], keep[:, 1]]                            # :495
    gid = k_img * (C + 1) + k_cls                                   # :519
    _, first = np.unique(gid, return_index=True)
    uniq = gid[np.sort(first)]                                      # tf.unique: first appearance (:520)

    per_class = []                                                  # rows (b,n,c), :522-526
    margins = []
    for g in uniq:
        ixs = np.nonzero(gid == g)[0]                               # :506 ascending
        sel = tf_ops.non_max_suppression(k_box[ixs], k_conf[ixs], nms_max_output_size,
                                         nms_iou_threshold, return_margin=return_debug)
        if return_debug:
            sel, m = sel
            margins.append(m)
        per_class.append(keep[ixs[sel]])
    per_class = (np.concatenate(per_class, axis=0) if per_class
                 else np.zeros((0, 3), dtype=np.int64))
    pc_img = per_class[:, 0]
    pc_conf = cls_pred[per_class[:, 0], per_class[:, 1], per_class[:, 2]]   # :528
    pc_box = norm[per_class[:, 0], per_class[:, 1]]                 # :529

    final = []                                                      # :531-555
    for b in range(B):
        ixs = np.nonzero(pc_img == b)[0]
        sel = tf_ops.non_max_suppression(pc_box[ixs], pc_conf[ixs], nms_max_output_size,
                                         post_iou_threshold, return_margin=return_debug)
        if return_debug:
            sel, m = sel
            margins.append(m)
        final.append(per_class[ixs[sel]])
    final = np.concatenate(final, axis=0) if final else np.zeros((0, 3), dtype=np.int64)

    conf = cls_pred[final[:, 0], final[:, 1], final[:, 2]][:, None]         # :557
    fbox = boxes[final[:, 0], final[:, 1]]                          # :558 un-normalised cxcywh
    fcls = final[:, 2].astype(F32)[:, None]                         # :559-562
    rows = np.concatenate([fbox, fcls, conf], axis=1).astype(F32).reshape(-1, 6)
    out = mold_batch(rows, final[:, 0], B, max_batch_size)          # :564-565
    if return_debug:
        return out, dict(candidates=keep, per_class_keep=per_class, keep=final,
                         min_iou_margin=min(margins) if margins else np.inf)
    return out


# ----------------------------------------------------------------- a9 -----
def mask_distribute(proposed, max_k=2, base_size=64, return_margin=False):
    """engine/layers/instance.py:52-66  MaskDistribute.call: [B,M,6] -> [B,M,7]."""
    x = np.asarray(proposed, dtype=F32)
    Hh, Ww = x[..., 2], x[..., 3]
    with np.errstate(invalid="ignore", divide="ignore"):
        size = np.sqrt(Hh * Ww)
        ratio = (size + K_EPSILON) / F32(float(base_size) + 1e-7)
        delta_k = log_f32(ratio) / log_f32(F32(2.0))
        k = np.floor(delta_k)
        k = np.minimum(np.maximum(k, F32(0)), F32(max_k))           # clip_by_value
    k = np.where(x[..., 0] == F32(-1.0), x[..., 0], k).astype(F32)
    out = np.concatenate([k[..., None], x], axis=-1).astype(F32)
    if return_margin:
        valid = x[..., 0] != F32(-1.0)
        with np.errstate(invalid="ignore"):
            d = np.abs(delta_k - np.round(delta_k))[valid]
        return out, (float(d.min()) if d.size else np.inf)
    return out


# ----------------------------------------------------------------- a10 ----
def pyramid_roi_align(fmaps, dist_boxes, image_hw, crop_size=(14, 14), max_batch_size=64):
    """engine/layers/instance.py:109-139  PyramidRoiAlign.call.

    fmaps: list of [B,Hf,Wf,Cf] f32 (levels 0..max_k); dist_boxes [B,M,7];
    image_hw = (H,W) of the model-input images.  Returns
    ([per level [B,Mf,ch,cw,Cf]], roi_boxes [B,sum Mf,6]).
    """
    dist = np.asarray(dist_boxes, dtype=F32)
    B = dist.shape[0]
    norm = normalize_boxes(dist[..., 1:5], shape=image_hw)          # :115-116
    roi_fmaps, roi_boxes = [], []
    for fmap_id, fmap in enumerate(fmaps):                          # :120
        idx = np.argwhere(dist[..., 0] == F32(fmap_id))             # :121 row-major (b,j)
        tnorm = norm[idx[:, 0], idx[:, 1]]
        bind = idx[:, 0].astype(np.int32)
        crops = tf_ops.crop_and_resize(fmap, tnorm, bind, crop_size)        # :125-126
        if crops.shape[0] == 0:
            crops = np.zeros((0, crop_size[0], crop_size[1], np.asarray(fmap).shape[-1]), F32)
        roi_fmaps.append(mold_batch(crops, bind, B, max_batch_size))        # :127-128
        tboxes = dist[idx[:, 0], idx[:, 1], 1:].reshape(-1, 6)              # :131
        roi_boxes.append(mold_batch(tboxes, bind, B, max_batch_size))       # :132-133
    return roi_fmaps, np.concatenate(roi_boxes, axis=1)             # :135-138


# ----------------------------------------------------------------- a11 ----
def trim_instances(roi_boxes, roi_masks, mold=True, max_batch_size=64):
    """engine/layers/instance.py:258-277  TrimInstances.call.

    roi_boxes [B,R,6], roi_masks [B,R,mh,mw,C] -> ([B,M,6], [B,M,mh,mw])."""
    rb = np.asarray(roi_boxes, dtype=F32)
    rm = np.asarray(roi_masks, dtype=F32)
    B = rb.shape[0]
    idx = np.argwhere(rb[:, :, -2] != F32(-1))                      # :262-263
    cls = rb[idx[:, 0], idx[:, 1], -2].astype(np.int64)             # :264-265
    tmask = rm[idx[:, 0], idx[:, 1], :, :, cls] if idx.shape[0] else \
        np.zeros((0,) + rm.shape[2:4], F32)                         # :267 transpose + gather_nd
    tbox = rb[idx[:, 0], idx[:, 1]].reshape(-1, rb.shape[-1])       # :268
    if not mold:
        return tbox, tmask
    return (mold_batch(tbox, idx[:, 0], B, max_batch_size),
            mold_batch(tmask, idx[:, 0], B, max_batch_size))


# ----------------------------------------------------------------- a12 ----
def upsample_output(roi_box, roi_mask, src_hw, dst_hw):
    """engine/layers/misc.py:169-188  UpSampleOutput.call (instance part).

    src_hw = semantic_output spatial shape (model resolution), dst_hw = target
    frame shape.  NB the reference multiplies cx,w by the HEIGHT ratio and cy,h
    by the WIDTH ratio (misc.py:180-183); reproduced as is.
    """
    rb = np.asarray(roi_box, dtype=F32)
    ratio = np.asarray(dst_hw, dtype=F32) / np.asarray(src_hw, dtype=F32)   # :177
    with np.errstate(invalid="ignore"):
        cx = (rb[..., 0] * ratio[0]).astype(np.int32)               # tf.cast truncates
        cy = (rb[..., 1] * ratio[1]).astype(np.int32)
        w = (rb[..., 2] * ratio[0]).astype(np.int32)
        h = (rb[..., 3] * ratio[1]).astype(np.int32)
        label = rb[..., 4].astype(np.int32)
        confs = (rb[..., 5] * F32(100)).astype(np.int32)
    box = np.stack([cx, cy, w, h, label, confs], axis=-1).astype(np.int32)
    mask = (np.asarray(roi_mask, dtype=F32) > F32(0.5)).astype(np.int32)    # :188
    return box, mask


# ------------------------------------------------------------ a13-a14 -----
def paste_geometry(det_row, image_h, image_w):
    """engine/layers/misc.py:372-386: clamp, ceil, clip -> (xmin,xmax,ymin,ymax)."""
    box = np.maximum(np.asarray(det_row, dtype=np.int32), 1).astype(F32)    # :373, :376
    cx, cy, w, h = box[0], box[1], box[2], box[3]
    xmin = int(np.clip(np.int32(np.ceil(cx - w / F32(2))), 0, image_w))
    xmax = int(np.clip(np.int32(np.ceil(cx + w / F32(2))), 0, image_w))
    ymin = int(np.clip(np.int32(np.ceil(cy - h / F32(2))), 0, image_h))
    ymax = int(np.clip(np.int32(np.ceil(cy + h / F32(2))), 0, image_h))
    return xmin, xmax, ymin, ymax


def crop_and_pad_mask(image_hw, det_outs, ins_outs):
    """engine/layers/misc.py:358-401  CropAndPadMask.call -> [B,M,PH,PW] f32.

    A box clipped to zero height or width makes tf.image.resize raise
    InvalidArgument in the reference; this restatement (and the CUDA path)
    defines that instance's mask as all zeros.
    """
    det = np.asarray(det_outs, dtype=np.int32)
    ins = np.asarray(ins_outs)
    PH, PW = int(image_hw[0]), int(image_hw[1])
    B, M, _ = det.shape
    mx = int(det[..., -1].max()) if det.size else -2 ** 31
    thr = 50 if mx > 50 else -100                                   # :366-369
    out = np.zeros((B, M, PH, PW), dtype=F32)                       # scatter_nd default
    for b, j in np.argwhere(det[..., -1] >= thr):                   # :370
        xmin, xmax, ymin, ymax = paste_geometry(det[b, j], PH, PW)
        if ymax - ymin <= 0 or xmax - xmin <= 0:
            continue
        out[b, j, ymin:ymax, xmin:xmax] = tf_ops.resize_bilinear_align_corners(
            ins[b, j], ymax - ymin, xmax - xmin)                    # :387-391
    return out


def binary_masks(pasted):
    """engine/layers/misc.py:457 / :611-615: consumers threshold with > 0.5."""
    return (np.asarray(pasted, dtype=F32) > F32(0.5)).astype(np.uint8)


def packed_masks(pasted):
    """Binary masks packed 8 pixels per byte, bit k of byte i = pixel 8*i+k."""
    return np.packbits(binary_masks(pasted), axis=-1, bitorder="little")


# ------------------------------------------------------- whole path -------
def synth_mask_head(roi_boxes, num_classes, mask_hw=(28, 28), seed=0):
    """Stand-in for MaskSubNet (not on the path): U(0,1) probs."""
    B, R, _ = np.asarray(roi_boxes).shape
    rng = np.random.default_rng(seed)
    return rng.random((B, R, mask_hw[0], mask_hw[1], num_classes), dtype=F32)


def full_path(loc_pred, cls_pred, fmaps, mask_head, prior_cfg, image_hw, frame_hw,
              min_confidence=0.05, nms_iou_threshold=0.4, post_iou_threshold=0.65,
              nms_max_output_size=1000, max_k=2, base_size=64, crop_size=(14, 14),
              padding="same"):
    """The chain of engine/retinamasklab.py:458-470, :615-616, :635-636 and
    road_project/setup/serving.py:30.  `mask_head(roi_fmaps, roi_boxes)` stands
    in for MaskSubNet.  Returns a dict of every intermediate."""
    B = np.asarray(cls_pred).shape[0]
    H, W = image_hw
    table = prior_table(**prior_cfg)
    pr = prior_layer(table, H, W, padding)
    pr_b = np.broadcast_to(pr[None], (B,) + pr.shape)
    restored = restore_boxes(loc_pred, pr_b)
    proposed = detection_proposal(cls_pred, restored, min_confidence, nms_iou_threshold,
                                  post_iou_threshold, nms_max_output_size)
    dist = mask_distribute(proposed, max_k, base_size)
    roi_fmaps, roi_boxes = pyramid_roi_align(fmaps[:max_k + 1], dist, image_hw, crop_size)
    roi_masks = mask_head(roi_fmaps, roi_boxes)
    det, ins = trim_instances(roi_boxes, roi_masks)
    det_i, ins_i = upsample_output(det, ins, image_hw, frame_hw)
    pasted = crop_and_pad_mask(frame_hw, det_i, ins_i)
    return dict(priors=pr, restored=restored, proposed=proposed, dist=dist, roi_fmaps=roi_fmaps,
                roi_boxes=roi_boxes, roi_masks=roi_masks, det=det, ins=ins, det_i=det_i,
                ins_i=ins_i, pasted=pasted, binary=binary_masks(pasted))
