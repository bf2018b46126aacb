"""ctypes front end of oracle/c/masklab_oracle.c (TEST INFRASTRUCTURE, see oracle/__init__.py).

The C restatement is a second, independently written CPU implementation of rows a2-a14; the
tests require it to agree bit for bit with the NumPy oracle, and bench.py uses it as the
multi-threaded CPU baseline ("port": restated reference path, not TensorFlow)."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libmasklab_oracle.so")
F32, I32 = np.float32, np.int32
_P, _I, _L, _F = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float


class PriorC(ctypes.Structure):
    _fields_ = [("num_levels", ctypes.c_int32), ("padding_same", ctypes.c_int32),
                ("stride", ctypes.c_int32 * 8), ("num_anchors", ctypes.c_int32 * 8),
                ("anchor_w", (ctypes.c_int32 * 32) * 8), ("anchor_h", (ctypes.c_int32 * 32) * 8)]


_lib = None


def build():
    subprocess.run(["make", "-s", "-C", os.path.join(HERE, "c")], check=True)
    return LIB


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        L = ctypes.CDLL(LIB)
        L.mlo_prior_count.restype = _L
        L.mlo_prior_count.argtypes = [ctypes.POINTER(PriorC), _I, _I]
        L.mlo_prior_layer.argtypes = [ctypes.POINTER(PriorC), _I, _I, _P]
        L.mlo_restore_boxes.argtypes = [_P, _P, _L, _P]
        L.mlo_detection_proposal.restype = _I
        L.mlo_detection_proposal.argtypes = [_P, _P, _I, _I, _I, _F, _F, _F, _I, _P, _P, _P]
        L.mlo_mask_distribute.argtypes = [_P, _L, _I, _F, _P]
        L.mlo_roi_plan.argtypes = [_P, _I, _I, _I, _P, _P]
        L.mlo_roi_run.argtypes = [ctypes.POINTER(_P), _P, _P, _I, _I, _P, _I, _I, _F, _F, _I, _I, _P,
                                  ctypes.POINTER(_P), _P]
        L.mlo_trim_plan.restype = _I
        L.mlo_trim_plan.argtypes = [_P, _I, _I, _P]
        L.mlo_trim_run.argtypes = [_P, _P, _I, _I, _I, _I, _I, _I, _P, _P]
        L.mlo_upsample.argtypes = [_P, _L, _F, _F, _P, _P, _L, _P]
        L.mlo_paste.argtypes = [_P, _P, _I, _I, _I, _I, _I, _I, _P, _P, _P]
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def prior_struct(prior_cfg, padding="same"):
    """prior_cfg: dict(strides,sizes,pr_scales,pr_ratios) -> PriorC (engine/prior.py:55-67 table,
    grouped by stride ascending like engine/layers/detection.py:260-262)."""
    rows = []
    for size, stride in zip(prior_cfg["sizes"], prior_cfg["strides"]):
        for s in prior_cfg["pr_scales"]:
            for r in prior_cfg["pr_ratios"]:
                rows.append((int(stride), int(np.round(size * s * np.sqrt(r))), int(np.round(size * s / np.sqrt(r)))))
    p = PriorC()
    strides = sorted(set(r[0] for r in rows))
    p.num_levels = len(strides)
    p.padding_same = 1 if padding == "same" else 0
    for l, st in enumerate(strides):
        whs = [(w, h) for s_, w, h in rows if s_ == st]
        p.stride[l] = st
        p.num_anchors[l] = len(whs)
        for a, (w, h) in enumerate(whs):
            p.anchor_w[l][a] = w
            p.anchor_h[l][a] = h
    return p


def prior_layer(prior_cfg, H, W, padding="same"):
    p = prior_struct(prior_cfg, padding)
    n = lib().mlo_prior_count(ctypes.byref(p), H, W)
    out = np.empty((n, 4), I32)
    lib().mlo_prior_layer(ctypes.byref(p), H, W, _ptr(out))
    return out


def restore_boxes(loc, prior):
    loc = np.ascontiguousarray(loc, F32)
    # the C loop walks loc and prior row by row: a [N,4] (or [1,N,4]) prior is broadcast over the batch here, the way
    # RestoreBoxes' `loc * prior` broadcasts it (detection.py:325-344)
    prior = np.ascontiguousarray(np.broadcast_to(np.asarray(prior, I32), loc.shape))
    out = np.empty_like(loc)
    lib().mlo_restore_boxes(_ptr(loc), _ptr(prior), loc.size // 4, _ptr(out))
    return out


def detection_proposal(cls, boxes, min_confidence=0.05, nms_iou_threshold=0.4, post_iou_threshold=0.65,
                       nms_max_output_size=1000, return_keep=False):
    cls = np.ascontiguousarray(cls, F32)
    boxes = np.ascontiguousarray(boxes, F32)
    B, N, C = cls.shape
    K = int(nms_max_output_size)
    det = np.empty((B, K, 6), F32)
    keep = np.empty((B, K, 2), I32)
    counts = np.empty((B,), I32)
    M = lib().mlo_detection_proposal(_ptr(cls), _ptr(boxes), B, N, C, min_confidence, nms_iou_threshold,
                                     post_iou_threshold, K, _ptr(det), _ptr(keep), _ptr(counts))
    out = np.ascontiguousarray(det[:, :M])
    return (out, keep[:, :M], counts) if return_keep else out


def mask_distribute(det, max_k=2, base_size=64):
    det = np.ascontiguousarray(det, F32)
    out = np.empty(det.shape[:-1] + (7,), F32)
    lib().mlo_mask_distribute(_ptr(det), det.size // 6, max_k, float(base_size), _ptr(out))
    return out


def pyramid_roi_align(fmaps, dist, image_hw, crop_size=(14, 14)):
    dist = np.ascontiguousarray(dist, F32)
    fm = [np.ascontiguousarray(f, F32) for f in fmaps]
    B, M, _ = dist.shape
    L, Cf = len(fm), fm[0].shape[-1]
    counts = np.empty((L, B), I32)
    level_m = np.empty((L + 1,), I32)
    lib().mlo_roi_plan(_ptr(dist), B, M, L, _ptr(counts), _ptr(level_m))
    ch, cw = crop_size
    crops = [np.empty((B, int(level_m[f]), ch, cw, Cf), F32) for f in range(L)]
    roi_boxes = np.empty((B, int(level_m[L]), 6), F32)
    fptr = (ctypes.c_void_p * L)(*[f.ctypes.data for f in fm])
    cptr = (ctypes.c_void_p * L)(*[c.ctypes.data for c in crops])
    fh = np.array([f.shape[1] for f in fm], I32)
    fw = np.array([f.shape[2] for f in fm], I32)
    lib().mlo_roi_run(fptr, _ptr(fh), _ptr(fw), L, Cf, _ptr(dist), B, M, float(image_hw[0]), float(image_hw[1]),
                      ch, cw, _ptr(level_m), cptr, _ptr(roi_boxes))
    return crops, roi_boxes


def trim_instances(roi_boxes, roi_masks):
    rb = np.ascontiguousarray(roi_boxes, F32)
    rm = np.ascontiguousarray(roi_masks, F32)
    B, R, _ = rb.shape
    mh, mw, C = rm.shape[2:]
    counts = np.empty((B,), I32)
    M = lib().mlo_trim_plan(_ptr(rb), B, R, _ptr(counts))
    ob = np.empty((B, M, 6), F32)
    om = np.empty((B, M, mh, mw), F32)
    lib().mlo_trim_run(_ptr(rb), _ptr(rm), B, R, mh, mw, C, M, _ptr(ob), _ptr(om))
    return ob, om


def upsample_output(det, masks, src_hw, dst_hw):
    det = np.ascontiguousarray(det, F32)
    masks = np.ascontiguousarray(masks, F32)
    ratio = np.asarray(dst_hw, F32) / np.asarray(src_hw, F32)
    di = np.empty(det.shape, I32)
    mi = np.empty(masks.shape, I32)
    lib().mlo_upsample(_ptr(det), det.size // 6, float(ratio[0]), float(ratio[1]), _ptr(di), _ptr(masks),
                       masks.size, _ptr(mi))
    return di, mi


def crop_and_pad_mask(frame_hw, det_i, mask_i, binary=False, bits=False):
    det_i = np.ascontiguousarray(det_i, I32)
    mask_i = np.ascontiguousarray(mask_i, I32)
    B, M, _ = det_i.shape
    mh, mw = mask_i.shape[2:]
    PH, PW = int(frame_hw[0]), int(frame_hw[1])
    if bits:
        assert PW % 8 == 0
        out = np.empty((B, M, PH, PW // 8), np.uint8)
        lib().mlo_paste(_ptr(det_i), _ptr(mask_i), B, M, mh, mw, PH, PW, None, None, _ptr(out))
    elif binary:
        out = np.empty((B, M, PH, PW), np.uint8)
        lib().mlo_paste(_ptr(det_i), _ptr(mask_i), B, M, mh, mw, PH, PW, None, _ptr(out), None)
    else:
        out = np.empty((B, M, PH, PW), F32)
        lib().mlo_paste(_ptr(det_i), _ptr(mask_i), B, M, mh, mw, PH, PW, _ptr(out), None, None)
    return out


def full_path(loc_pred, cls_pred, fmaps, mask_head, prior_cfg, image_hw, frame_hw, min_confidence=0.05,
              nms_iou_threshold=0.4, post_iou_threshold=0.65, nms_max_output_size=1000, max_k=2,
              base_size=64, crop_size=(14, 14), padding="same", binary=True, bits=False):
    """Same chain as masklab_oracle.full_path, every stage in C."""
    B = cls_pred.shape[0]
    H, W = image_hw
    pr = prior_layer(prior_cfg, H, W, padding)
    restored = restore_boxes(loc_pred, np.broadcast_to(pr[None], (B,) + pr.shape))
    proposed = detection_proposal(cls_pred, restored, min_confidence, nms_iou_threshold, post_iou_threshold,
                                  nms_max_output_size)
    dist = mask_distribute(proposed, max_k, base_size)
    roi_fmaps, roi_boxes = pyramid_roi_align(fmaps[:max_k + 1], dist, image_hw, crop_size)
    roi_masks = mask_head(roi_fmaps, roi_boxes)
    det, ins = trim_instances(roi_boxes, roi_masks)
    det_i, ins_i = upsample_output(det, ins, image_hw, frame_hw)
    pasted = crop_and_pad_mask(frame_hw, det_i, ins_i, binary=binary, bits=bits)
    return dict(priors=pr, restored=restored, proposed=proposed, dist=dist, roi_fmaps=roi_fmaps,
                roi_boxes=roi_boxes, det=det, ins=ins, det_i=det_i, ins_i=ins_i,
                **({"bits": pasted} if bits else ({"binary": pasted} if binary else {"pasted": pasted})))
