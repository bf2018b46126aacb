"""CPU pins of the SummaryOutput restatement (oracle/summary_oracle.py): hand-derived known answers
and a brute-force cross-check.  The reference has no tests for these layers ("parity unpinned")."""
import numpy as np

import synth
from oracle import masklab_oracle as mo
from oracle import summary_oracle as so

F32 = np.float32


def test_fit_line_exact_and_degenerate():
    ys = np.arange(10, 60)
    t0, t1 = so.fit_line(np.stack([ys, 2 * ys + 3], axis=1))
    assert (t0, t1) == (F32(2), F32(3))
    assert so.fit_line(np.zeros((0, 2))) == (F32(0), F32(0))          # det = 0 -> zeros (misc.py:715-717)
    assert so.fit_line(np.array([[5, 9]])) == (F32(0), F32(0))         # one point: det = 0


def test_road_marginals_drop_rules():
    road = np.zeros((40, 50), dtype=np.int32)
    for y in range(10, 30):
        road[y, 10:20 + y] = 1
    road[12, :] = 0                       # a gap: segment_min/max give (0, 0) -> dropped
    road[15, :] = 0
    road[15, 7] = 1                       # single pixel: x_min == x_max -> dropped
    left, right = so.road_marginals(road)
    # 18 valid rows, drop = int(18 * 0.15) = 2 at both ends -> 14 rows
    assert left.shape == (14, 2) and right.shape == (14, 2)
    valid = [y for y in range(10, 30) if y not in (12, 15)]
    assert left[:, 0].tolist() == valid[2:-2]
    assert np.all(left[:, 1] == 10) and np.array_equal(right[:, 1], 19 + left[:, 0])
    # fewer than 7 valid rows: int(n * 0.15) = 0 -> clipped to 1
    road2 = np.zeros((10, 10), dtype=np.int32)
    road2[2:7, 1:4] = 1
    l2, _ = so.road_marginals(road2)
    assert l2[:, 0].tolist() == [3, 4, 5]
    e, _ = so.road_marginals(np.zeros((4, 4), dtype=np.int32))
    assert e.shape == (0, 2)


def test_unit_lengths_known_answer():
    PH, PW = 64, 200
    road = np.zeros((PH, PW), dtype=np.int32)
    for y in range(PH):
        road[y, 50 - y // 2: 150 + y // 2 + 1] = 1      # not a line (floor), but close
    unit = so.road_unit_lengths(road, 3.25)
    assert unit.dtype == F32 and unit.shape == (PH,)
    width = 100 + 2 * (np.arange(PH) // 2)
    assert np.allclose(F32(3.25) / unit, width, atol=1.0)
    # no road at all: theta = 0 -> width clipped to 1 -> unit = default_road_size
    assert np.all(so.road_unit_lengths(np.zeros((8, 8), dtype=np.int32), 3.25) == F32(3.25))


def test_crack_to_instance():
    crack = np.zeros((2, 20, 30), dtype=np.int32)
    det, seg = so.crack_to_instance(crack)
    assert det.shape == (2, 1, 6) and np.array_equal(det[0, 0], [0, 0, 0, 0, 5, 0])
    crack[0, 4, 7] = 1
    crack[1, 11, 21] = 1                  # the box spans the whole batch
    det, seg = so.crack_to_instance(crack)
    assert np.array_equal(det[1, 0], [7 + 7, 4 + 3, 14, 7, 5, 100])
    assert seg.shape == (2, 1, 20, 30) and seg.dtype == F32 and seg.sum() == 2
    crack[:] = 0
    crack[0, 5, 3:9] = 1                  # zero height -> conf 0 -> not appended by SummaryOutput
    assert so.crack_to_instance(crack)[0][0, 0, 5] == 0


def test_summary_known_answer():
    B, PH, PW = 1, 32, 48
    seg = np.zeros((B, PH, PW, 3), dtype=np.int32)
    seg[0, :, 8:40, 1] = 1                # road 32 px wide on every row -> 31 between borders
    masks = np.zeros((B, 2, PH, PW), dtype=F32)
    masks[0, 0, 10:14, 12:20] = 1.0       # 4 x 8 block on the road
    masks[0, 1, 2:5, 42:46] = 0.75        # 3 x 4 block off the road
    det = np.array([[[16, 12, 8, 4, 2, 90], [44, 3, 4, 3, 1, 60]]], dtype=np.int32)
    out = so.summary_output(det, seg, masks)
    assert out.shape == (1, 2, 11)                       # no crack -> nothing appended
    u = F32(3.25) / F32(31.0)
    assert np.array_equal(out[0, :, :6], [[2, 16, 12, 8, 4, 90], [1, 44, 3, 4, 3, 60]])
    assert out[0, 0, 6] == 32 and out[0, 1, 6] == 9
    assert np.isclose(out[0, 0, 7], 32 * float(u) ** 2, rtol=1e-6)
    assert np.isclose(out[0, 0, 8], 4 * float(u), rtol=1e-6)          # column sums: 4 rows
    assert np.isclose(out[0, 0, 9], 4 * float(u), rtol=1e-6)          # rows with a pixel > 0.5
    assert np.isclose(out[0, 1, 8], 3 * 0.75 * float(u), rtol=1e-6)
    assert out[0, 0, 10] == 1 and out[0, 1, 10] == 0
    # with a crack region the pseudo-instance is appended as the last row
    seg[0, 20:24, 10:30, 2] = 1
    out2 = so.summary_output(det, seg, masks)
    assert out2.shape == (1, 3, 11)
    assert np.array_equal(out2[0, 2, :6], [5, 10 + 9, 20 + 1, 19, 3, 100])
    assert out2[0, 2, 6] == 80 and out2[0, 2, 10] == 1
    assert np.array_equal(out2[:, :2], out)


def test_reductions_match_brute_force_on_pasted_masks():
    B, M, PH, PW, C = 2, 5, 48, 80, 4
    seg = synth.semantic_map(B, PH, PW, seed=11)
    det = synth.int_detections(B, M, C, PH, PW, seed=12, pad_tail=1)
    rng = np.random.default_rng(13)
    ins = (rng.random((B, M, 28, 28)) > 0.4).astype(np.int32)
    masks = mo.crop_and_pad_mask((PH, PW), det, ins)
    red = so.instance_reductions(seg, masks)
    for b in range(B):
        unit = so.road_unit_lengths(seg[b, :, :, 1]).astype(np.float64)
        for j in range(M):
            m = masks[b, j].astype(np.float64)
            pix = sum(float(v) for v in m.ravel())
            assert np.isclose(red[b, j, 0], pix, rtol=1e-6)
            size = sum(unit[y] ** 2 * m[y].sum() for y in range(PH))
            assert np.isclose(red[b, j, 1], size, rtol=1e-5)
            horiz = max(sum(unit[y] * m[y, x] for y in range(PH)) for x in range(PW))
            assert np.isclose(red[b, j, 2], horiz, rtol=1e-6)
            vert = sum(unit[y] for y in range(PH) if (masks[b, j, y] > 0.5).any())
            assert np.isclose(red[b, j, 3], vert, rtol=1e-6)
            on = masks[b, j] > 0.5
            inter = np.logical_and(on, seg[b, :, :, 1] > 0).sum()
            want = 1.0 if inter / (on.sum() + 1e-5) > 0.1 else 0.0
            assert red[b, j, 4] == want
