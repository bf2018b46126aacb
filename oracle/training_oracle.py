"""NumPy restatement of the training-side target assignment: CalculateIOU,
AssignBoxes, AssignMasks and DetectionIOUMetric.  TEST INFRASTRUCTURE — see oracle/__init__.py
("parity unpinned").

Reference: /root/reference/engine/layers/detection.py:378-422 (CalculateIOU), :589-697 (AssignBoxes),
/root/reference/engine/layers/instance.py:296-386 (AssignMasks), /root/reference/engine/metrics.py:109-165
(DetectionIOUMetric).  float32, one rounding per operation, reference operation order.

Two places where TensorFlow's result depends on the execution order are fixed the way the CPU
kernels behave: tf.tensor_scatter_nd_update with repeated indices keeps the LAST update of the
index list, tf.scatter_nd adds repeated updates in list order (AssignBoxes scatters the regression
targets with scatter_nd, so a prior matched both by IoU >= 0.5 and as a ground truth's best prior
receives its target twice - reproduced as is).  log() is the correctly rounded float32 value, like
exp()/log() elsewhere in the oracle.
"""
import numpy as np

from . import masklab_oracle as mo
from . import tf_ops

F32 = np.float32


def calculate_iou(aa_boxes, bb_boxes):
    """detection.py:391-422: IoU matrix [Na, Nb] of (cx,cy,w,h) boxes; union + 1e-5 in the divisor."""
    aa = np.asarray(aa_boxes).astype(F32).reshape(-1, 4)
    bb = np.asarray(bb_boxes).astype(F32).reshape(-1, 4)
    area_b = bb[:, 2] * bb[:, 3]                 # "aa_area" in the reference
    area_a = aa[:, 2] * aa[:, 3]                 # "bb_area"
    areas = area_b[None, :] + area_a[:, None]
    an = mo.normalize_boxes(aa)
    bn = mo.normalize_boxes(bb)
    ay1, ax1, ay2, ax2 = [an[:, None, k] for k in range(4)]
    by1, bx1, by2, bx2 = [bn[None, :, k] for k in range(4)]
    in_w = np.maximum(F32(0), np.minimum(bx2, ax2) - np.maximum(bx1, ax1))
    in_h = np.maximum(F32(0), np.minimum(by2, ay2) - np.maximum(by1, ay1))
    inter = (in_w * in_h).astype(F32)
    union = areas - inter
    return (inter / (union + F32(1e-5))).astype(F32)


def assign_boxes(gt_boxes, pr_boxes, num_classes):
    """detection.py:619-690 -> (cls_true [B,N,C], loc_true [B,N,4], assign_mask [B,N,1])."""
    gt = np.asarray(gt_boxes).astype(F32)
    pr = np.asarray(pr_boxes).astype(F32)
    B, G, _ = gt.shape
    N = pr.shape[1]
    labels, conf = gt[..., 4], gt[..., 5]
    iou = calculate_iou(gt[..., :4].reshape(-1, 4), pr[0]).reshape(B, G, N)
    iou = iou * (gt[..., 0] != F32(-1)).astype(F32)[..., None]
    match = np.argwhere(iou >= F32(0.5))                                   # row-major (b, g, n)
    top = iou.reshape(-1, N).argmax(axis=1)                                # first maximum
    bs, gs = np.divmod(np.arange(B * G), G)
    best = np.stack([bs, gs, top], axis=1)[conf.reshape(-1) > 0]
    match = np.concatenate([match.reshape(-1, 3), best.reshape(-1, 3)], axis=0).astype(np.int64)
    mb, mg, mn = match[:, 0], match[:, 1], match[:, 2]
    cls_true = np.full((B, N), F32(-1))
    cls_true[mb, mn] = labels[mb, mg]                                      # repeated index: last one wins
    cls_true = np.where(cls_true != F32(-1), cls_true, F32(num_classes))
    one_hot = np.zeros((B, N, num_classes + 1), dtype=F32)
    ci = cls_true.astype(np.int32)
    ok = (ci >= 0) & (ci <= num_classes)
    bb_, nn_ = np.nonzero(ok)
    one_hot[bb_, nn_, ci[ok]] = 1
    assign = one_hot[..., -1].copy()
    ig = np.argwhere((iou < F32(0.5)) & (iou >= F32(0.4)))
    ignore = np.zeros((B, N), dtype=np.int64)
    np.add.at(ignore, (ig[:, 0], ig[:, 2]), 1)
    assign = np.where(ignore > 0, F32(-1), assign).astype(F32)
    p = pr[mb, mn]
    g = gt[mb, mg, :4]
    hat = np.stack([(g[:, 0] - p[:, 0]) / p[:, 2], (g[:, 1] - p[:, 1]) / p[:, 3],
                    mo.log_f32(g[:, 2] / p[:, 2]), mo.log_f32(g[:, 3] / p[:, 3])], axis=1).astype(F32)
    loc_true = np.zeros((B, N, 4), dtype=F32)
    for k in range(4):
        np.add.at(loc_true[..., k], (mb, mn), hat[:, k])                   # repeated index: summed in order
    return one_hot[..., :num_classes].copy(), loc_true, assign[..., None]


def assign_masks(roi_boxes, roi_masks_shape, gt_boxes, gt_masks, match_iou_threshold=0.5):
    """instance.py:330-380 -> match_gt_masks int32 [B,R,mh,mw]: per RoI the mask of its best ground
    truth (same class, both rows valid) cropped to the RoI and thresholded, holding the class id
    where the crop is > 0.5 and num_classes elsewhere (also everywhere when the best IoU is below
    the threshold).  roi_masks_shape = (mh, mw, C) of the mask head output."""
    rb = np.asarray(roi_boxes).astype(F32)
    gb = np.asarray(gt_boxes).astype(F32)
    gm = np.asarray(gt_masks).astype(F32)
    mh, mw, C = roi_masks_shape
    B, R = rb.shape[:2]
    H, W = gm.shape[2], gm.shape[3]
    out = np.zeros((B, R, mh, mw), dtype=np.int32)
    for b in range(B):
        norm = mo.normalize_boxes(rb[b], shape=(H, W))
        iou = calculate_iou(gb[b, :, :4], rb[b, :, :4])
        valid = ((gb[b, :, None, -1] != F32(-1)) & (rb[b, None, :, -1] != F32(-1))).astype(F32)
        same = (gb[b, :, None, -2] == rb[b, None, :, -2]).astype(F32)
        iou = iou * valid * same
        matched = iou.max(axis=0) >= F32(match_iou_threshold)
        gi = iou.argmax(axis=0)
        cls = np.where(matched, gb[b, gi, 4], F32(C)).astype(F32)
        crops = tf_ops.crop_and_resize(gm[b][..., None], norm, gi, (mh, mw))[..., 0]
        out[b] = np.where(crops > F32(0.5), cls[:, None, None], F32(C)).astype(np.int32)
    return out


def detection_iou_metric(proposed_boxes, gt_boxes):
    """metrics.py:117-160 -> (precision, recall, fmeasure), float32 [B] each."""
    pb = np.asarray(proposed_boxes).astype(F32)
    gb = np.asarray(gt_boxes).astype(F32)
    B = pb.shape[0]
    eps = F32(1e-7)                                           # K.epsilon()
    prec, rec, fm = [], [], []
    for b in range(B):
        iou = calculate_iou(pb[b, :, :4], gb[b, :, :4])
        keep = ((pb[b, :, 0] != F32(-1))[:, None] | (gb[b, :, 0] != F32(-1))[None, :]).astype(F32)
        iou = iou * keep
        num_pos = F32((iou.max(axis=1) > F32(0.5)).sum())
        num_true = F32((iou.max(axis=0) > F32(0.5)).sum())
        num_pred = F32((pb[b, :, 0] != F32(-1)).sum())
        num_gt = F32((gb[b, :, 0] != F32(-1)).sum())
        p = num_pos / (num_pred + eps)
        r = num_true / (num_gt + eps)
        prec.append(p)
        rec.append(r)
        fm.append(F32(2) * (p * r) / (p + r + eps))
    return np.asarray(prec, F32), np.asarray(rec, F32), np.asarray(fm, F32)
