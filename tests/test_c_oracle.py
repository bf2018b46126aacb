"""The C restatement (oracle/c) must agree bit for bit with the NumPy oracle: two independently
written CPU implementations of the same reference semantics (both are test infrastructure)."""
import numpy as np
import pytest

import synth
from oracle import c_oracle as co
from oracle import masklab_oracle as mo

F32 = np.float32


@pytest.mark.parametrize("mu,max_out,C", [(-5.0, 60, 4), (-4.0, 25, 3), (-5.8, 100, 5)])
def test_full_path_c_equals_numpy(mu, max_out, C):
    B, H, W, Cf = 3, 128, 256, 8
    cfgp = synth.prior_config()
    N = synth.num_anchors(cfgp, H, W)
    loc, cls = synth.head_tensors(B, N, C, mu=mu, seed=5)
    cls[1] = 0
    fmaps = synth.fpn_maps(B, H, W, Cf, seed=6)
    head = lambda f, b: synth.mask_probs(B, b.shape[1], C, seed=7)
    kw = dict(nms_max_output_size=max_out, base_size=36, nms_iou_threshold=0.4, post_iou_threshold=0.6)
    a = mo.full_path(loc, cls, fmaps, head, cfgp, (H, W), (192, 320), **kw)
    b = co.full_path(loc, cls, fmaps, head, cfgp, (H, W), (192, 320), binary=False, **kw)
    for k in ("priors", "restored", "proposed", "dist", "roi_boxes", "det", "ins", "det_i", "ins_i", "pasted"):
        assert np.array_equal(a[k], b[k]), k
    for x, y in zip(a["roi_fmaps"], b["roi_fmaps"]):
        assert np.array_equal(x, y)
    c = co.full_path(loc, cls, fmaps, head, cfgp, (H, W), (192, 320), binary=True, **kw)
    assert np.array_equal(c["binary"], a["binary"])
    d = co.full_path(loc, cls, fmaps, head, cfgp, (H, W), (192, 320), bits=True, **kw)
    assert np.array_equal(d["bits"], mo.packed_masks(a["pasted"]))


def test_detection_ties_and_keep_indices():
    cfgp = synth.prior_config()
    H = W = 64
    pr = mo.prior_layer(mo.prior_table(**cfgp), H, W)
    loc, cls = synth.head_tensors(2, pr.shape[0], 3, mu=-4.5, seed=9)
    cls = np.round(cls * 20).astype(F32) / 20
    boxes = mo.restore_boxes(loc, np.broadcast_to(pr[None], (2,) + pr.shape))
    want, dbg = mo.detection_proposal(cls, boxes, nms_max_output_size=80, return_debug=True)
    got, keep, counts = co.detection_proposal(cls, boxes, nms_max_output_size=80, return_keep=True)
    assert np.array_equal(got, want)
    for b in range(2):
        rows = dbg["keep"][dbg["keep"][:, 0] == b]
        assert counts[b] == rows.shape[0]
        assert np.array_equal(keep[b, :counts[b]], rows[:, 1:].astype(np.int32))


def test_paste_threshold_branches_and_edge_boxes():
    PH, PW = 40, 56
    det = synth.detections(2, 12, 3, PH, PW, seed=11, lo=2.0, hi=40.0, pad_tail=2)
    masks = synth.mask_probs(2, 12, 1, seed=12)[..., 0]
    di, mi = mo.upsample_output(det, masks, (PH, PW), (PH, PW))
    di[0, 0, :4] = [PW + 30, 5, 10, 10]
    di[0, 1, :4] = [2, 2, 30, 30]
    for cap in (None, 40):
        d = di.copy()
        if cap:
            d[..., 5] = np.minimum(d[..., 5], cap)
        assert np.array_equal(co.crop_and_pad_mask((PH, PW), d, mi), mo.crop_and_pad_mask((PH, PW), d, mi))
