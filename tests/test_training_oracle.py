"""CPU pins of oracle/training_oracle.py (training-side target assignment): hand-derived answers."""
import numpy as np

from oracle import training_oracle as to

F32 = np.float32


def test_calculate_iou_known_answers():
    a = np.array([[5, 5, 10, 10], [5, 5, 10, 10], [100, 100, 4, 4]], F32)
    b = np.array([[5, 5, 10, 10], [10, 5, 10, 10], [5, 5, 0, 0]], F32)
    iou = to.calculate_iou(a, b)
    assert iou.shape == (3, 3)
    assert np.isclose(iou[0, 0], 100 / (100 + 1e-5), rtol=1e-6)            # identical boxes
    assert np.isclose(iou[0, 1], 50 / (150 + 1e-5), rtol=1e-6)             # half overlap
    assert iou[2, 0] == 0 and iou[0, 2] == 0                               # disjoint / zero area


def test_assign_boxes_small_case():
    pr = np.array([[[8, 8, 16, 16], [24, 8, 16, 16], [8, 24, 16, 16], [24, 24, 16, 16]]] * 2, F32)
    gt = -np.ones((2, 2, 6), F32)
    gt[0, 0] = [8, 8, 16, 16, 2, 1]            # exactly prior 0: IoU ~1 -> matched by IoU AND as the best prior
    gt[0, 1] = [24, 25, 16, 14, 1, 1]          # best prior 3, IoU 0.8125 >= 0.5
    gt[1, 0] = [40, 40, 4, 4, 0, 1]            # overlaps nothing: argmax of zeros = prior 0
    cls, loc, mask = to.assign_boxes(gt, pr, 3)
    assert cls.shape == (2, 4, 3) and loc.shape == (2, 4, 4) and mask.shape == (2, 4, 1)
    assert cls[0, 0].tolist() == [0, 0, 1] and cls[0, 3].tolist() == [0, 1, 0]
    assert cls[0, 1].sum() == 0 and cls[0, 2].sum() == 0
    assert cls[1, 0].tolist() == [1, 0, 0]                                  # the quirk: prior 0 gets the far box
    assert mask[0, :, 0].tolist() == [0, 1, 1, 0] and mask[1, :, 0].tolist() == [0, 1, 1, 1]
    assert np.all(loc[0, 0] == 0)                                           # exact match: zero offsets (twice)
    # prior 3 receives its target twice (scatter_nd adds repeated updates)
    want = np.array([0.0, (25 - 24) / 16, 0.0, np.log(14 / 16)], F32)
    assert np.allclose(loc[0, 3], 2 * want, rtol=1e-6)
    assert np.allclose(loc[1, 0], [2 * 1.0 * (40 - 8) / 16 / 2, (40 - 8) / 16, np.log(4 / 16), np.log(4 / 16)], rtol=1e-6)


def test_assign_boxes_ignore_band_and_last_update_wins():
    pr = np.array([[[10, 10, 20, 20], [100, 100, 20, 20]]], F32)
    gt = -np.ones((1, 3, 6), F32)
    gt[0, 0] = [10, 10, 20, 20, 0, 1]
    gt[0, 1] = [10, 10, 20, 20, 1, 1]          # same box, other class: the later update wins
    gt[0, 2] = [100, 106, 20, 20, 2, 1]        # IoU with prior 1 = 14/26 = 0.538 -> matched
    cls, loc, mask = to.assign_boxes(gt, pr, 3)
    assert cls[0, 0].tolist() == [0, 1, 0] and cls[0, 1].tolist() == [0, 0, 1]
    gt[0, 2] = [100, 108, 20, 20, 2, 1]        # IoU = 12/28 = 0.4286: ignore band, but still its best prior
    cls, loc, mask = to.assign_boxes(gt, pr, 3)
    assert cls[0, 1].tolist() == [0, 0, 1] and mask[0, 1, 0] == -1


def test_detection_iou_metric_small_case():
    pred = -np.ones((1, 3, 6), F32)
    gt = -np.ones((1, 2, 6), F32)
    pred[0, 0] = [10, 10, 20, 20, 0, .9]
    pred[0, 1] = [60, 60, 20, 20, 0, .8]
    gt[0, 0] = [10, 12, 20, 20, 0, 1]          # IoU 0.82 with pred 0
    p, r, f = to.detection_iou_metric(pred, gt)
    assert np.isclose(p[0], 1 / 2, rtol=1e-6) and np.isclose(r[0], 1.0, rtol=1e-6)
    assert np.isclose(f[0], 2 * 0.5 / 1.5, rtol=1e-5)


def test_assign_masks_small_case():
    H, W = 32, 32
    gt_boxes = -np.ones((1, 2, 6), F32)
    gt_boxes[0, 0] = [8, 8, 16, 16, 1, 1]
    gt_masks = np.zeros((1, 2, H, W), F32)
    gt_masks[0, 0, 0:16, 0:8] = 1              # left half of the box
    roi = -np.ones((1, 3, 6), F32)
    roi[0, 0] = [8, 8, 16, 16, 1, .9]          # same box, same class -> matched
    roi[0, 1] = [8, 8, 16, 16, 2, .9]          # other class -> unmatched
    out = to.assign_masks(roi, (4, 4, 3), gt_boxes, gt_masks)
    assert out.shape == (1, 3, 4, 4) and out.dtype == np.int32
    # samples at 0, 5.17, 10.33, 15.5 of the 32-pixel map: the last row sits between mask row 15 (1) and
    # row 16 (0) -> 0.5, which is not > 0.5
    assert np.all(out[0, 0, :3, :2] == 1) and np.all(out[0, 0, :, 2:] == 3) and np.all(out[0, 0, 3] == 3)
    assert np.all(out[0, 1] == 3) and np.all(out[0, 2] == 3)
