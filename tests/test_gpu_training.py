"""GPU parity of the training-side target assignment against
oracle/training_oracle.py through the C ABI: everything bit-exact (float32, same operation order)."""
import numpy as np
import pytest
import torch

import synth
from oracle import masklab_oracle as mo
from oracle import training_oracle as to

pytestmark = pytest.mark.gpu

F32 = np.float32


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def ground_truth(B, G, C, H, W, seed, pad=2):
    rng = np.random.default_rng(seed)
    gt = synth.detections(B, G, C, H, W, seed=seed, lo=12.0, hi=200.0, pad_tail=pad)
    gt[..., 5] = np.where(gt[..., 0] == -1, -1, 1).astype(F32)          # confidence 1 on valid rows
    return gt.astype(F32)


def test_calculate_iou():
    import masklab_b200 as ml
    a = synth.detections(1, 37, 3, 200, 300, seed=1)[0]
    b = synth.detections(1, 53, 3, 200, 300, seed=2)[0]
    b[5, :4] = a[3, :4]                                                  # an identical pair
    b[6, 2:4] = 0                                                        # zero area
    want = to.calculate_iou(a[:, :4], b[:, :4])
    got = ml.CalculateIOU()([dev(a), dev(b)]).cpu().numpy()              # 6-column rows: only [:4] is read
    assert got.shape == (37, 53) and np.array_equal(got, want)


@pytest.mark.parametrize("H,W,G,int_priors", [(128, 256, 12, True), (96, 160, 5, False), (540, 960, 20, True)])
def test_assign_boxes(H, W, G, int_priors):
    import masklab_b200 as ml
    B, C = 3, 5
    cfgp = synth.prior_config()
    pr = mo.prior_layer(mo.prior_table(**cfgp), H, W)                    # [N,4] int32
    prb = np.broadcast_to(pr[None], (B,) + pr.shape).copy()
    gt = ground_truth(B, G, C, H, W, seed=H)
    gt[0, 0, :4] = pr[len(pr) // 3].astype(F32)                          # exactly a prior: matched twice
    gt[0, 1, :4] = gt[0, 0, :4]                                          # same box, next class: last update wins
    gt[0, 1, 4] = (gt[0, 0, 4] + 1) % C
    gt[1, 0, :4] = [3 * W, 3 * H, 8, 8]                                  # overlaps nothing: argmax of zeros = prior 0
    want = to.assign_boxes(gt, prb, C)
    layer = ml.AssignBoxes(num_classes=C)
    got = layer([dev(gt), dev(prb if int_priors else prb.astype(F32))])
    for g, w, name in zip(got, want, ("cls_true", "loc_true", "assign_mask")):
        assert tuple(g.shape) == w.shape, name
        assert np.array_equal(g.cpu().numpy(), w), name
    assert want[0].sum() > 0 and (want[2] == -1).any() and (want[2] == 1).any()
    assert layer.get_config()["num_classes"] == C


def test_assign_masks():
    import masklab_b200 as ml
    B, G, R, C, H, W, mh, mw = 2, 6, 15, 4, 96, 128, 28, 28
    gt = ground_truth(B, G, C, H, W, seed=7)
    rng = np.random.default_rng(8)
    gm = np.zeros((B, G, H, W), F32)
    for b in range(B):
        for g in range(G):
            if gt[b, g, 0] == -1:
                continue
            cx, cy, w, h = gt[b, g, :4]
            yy, xx = np.mgrid[0:H, 0:W]
            gm[b, g] = ((((xx - cx) / (w / 2 + 1)) ** 2 + ((yy - cy) / (h / 2 + 1)) ** 2) < 1).astype(F32)   # ellipse
    roi = -np.ones((B, R, 6), F32)
    for b in range(B):
        for r in range(R - 2):
            g = r % G
            roi[b, r] = gt[b, g]
            roi[b, r, :4] += rng.normal(0, 4, 4).astype(F32)            # jittered copies of the ground truth
            if r % 5 == 4:
                roi[b, r, 4] = (roi[b, r, 4] + 1) % C                    # wrong class: never matched
    want = to.assign_masks(roi, (mh, mw, C), gt, gm, 0.5)
    got = ml.AssignMasks(0.5)([dev(roi), torch.empty((B, R, mh, mw, C), device="cuda"), dev(gt), dev(gm)])
    assert got.dtype == torch.int32 and np.array_equal(got.cpu().numpy(), want)
    assert (want != C).any() and (want == C).any()
    want7 = to.assign_masks(roi, (7, 9, C), gt, gm, 0.7)
    got7 = ml.AssignMasks(0.7)([dev(roi), (B, R, 7, 9, C), dev(gt), dev(gm)])
    assert np.array_equal(got7.cpu().numpy(), want7)


def test_detection_iou_metric():
    import masklab_b200 as ml
    B, P, G, C, H, W = 4, 30, 9, 3, 200, 300
    gt = ground_truth(B, G, C, H, W, seed=11)
    pred = synth.detections(B, P, C, H, W, seed=12, lo=12.0, hi=200.0, pad_tail=6)
    pred[:, :5, :4] = gt[:, :5, :4] + np.float32(1.5)                    # some good predictions
    pred[3] = -1                                                         # an image without predictions
    want = to.detection_iou_metric(pred, gt)
    got = ml.DetectionIOUMetric()([dev(pred), dev(gt)])
    for g, w in zip(got, want):
        assert np.array_equal(g.cpu().numpy(), w)
    assert want[0][0] > 0 and want[1][0] > 0
