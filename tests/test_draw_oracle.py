"""CPU pins of the overlay restatement (oracle/draw_oracle.py): hand-derived known answers."""
import numpy as np

from oracle import draw_oracle as do

F32 = np.float32


def test_draw_segmentation_known_answer():
    img = np.zeros((1, 2, 2, 3), dtype=np.uint8)
    img[0, 0, 0] = [10, 250, 100]
    img[0, 1, 1] = [200, 200, 200]
    seg = np.zeros((1, 2, 2, 2), dtype=np.int32)
    seg[0, 0, 0, 1] = 1                    # colour 1 on pixel (0,0)
    seg[0, 1, 1, :] = 1                    # both colours on pixel (1,1)
    colors = [[64, 0, 128], [128, 96, 0]]
    out = do.draw_segmentation(img, seg, colors, 0.3)
    assert out.dtype == np.uint8
    a = F32(0.3)
    assert out[0, 0, 0].tolist() == [int(F32(10) + F32(128) * a), 255, int(F32(100) + F32(0) * a)]
    assert out[0, 1, 1].tolist() == [min(255, int(F32(200) + F32(192) * a)), int(F32(200) + F32(96) * a),
                                     int(F32(200) + F32(128) * a)]
    assert out[0, 0, 1].tolist() == [0, 0, 0]


def test_class_masks_sum_then_threshold():
    det = np.array([[[0, 0, 0, 0, 1, 90], [0, 0, 0, 0, 1, 80], [0, 0, 0, 0, 0, 70], [-1, -1, -1, -1, -1, -100]]],
                   dtype=np.int32)
    masks = np.zeros((1, 4, 1, 3), dtype=F32)
    masks[0, 0, 0] = [0.3, 0.6, 0.0]
    masks[0, 1, 0] = [0.3, 0.0, 0.2]       # 0.3 + 0.3 > 0.5: two weak instances of one class add up
    masks[0, 2, 0] = [0.0, 0.0, 1.0]
    masks[0, 3, 0] = [1.0, 1.0, 1.0]       # padding row: class -1 matches no class id
    cm = do.class_masks(det, masks, 2)
    assert cm.shape == (1, 1, 3, 2)
    assert cm[0, 0, :, 1].tolist() == [1.0, 1.0, 0.0]
    assert cm[0, 0, :, 0].tolist() == [0.0, 0.0, 1.0]
    img = np.full((1, 1, 3, 3), 100, dtype=np.uint8)
    out = do.draw_instance(img, det, masks, [[192, 32, 128], [160, 96, 0]], 0.3)
    assert out[0, 0, 0].tolist() == [int(F32(100) + F32(160) * F32(0.3)), int(F32(100) + F32(96) * F32(0.3)), 100]
    assert out[0, 0, 2].tolist() == [int(F32(100) + F32(192) * F32(0.3)), int(F32(100) + F32(32) * F32(0.3)),
                                     int(F32(100) + F32(128) * F32(0.3))]


def test_draw_boxes_known_answer():
    img = np.full((1, 10, 12, 3), 7, dtype=np.uint8)
    # cx=6, cy=5, w=4, h=6 -> x in [4,8]/12, y in [2,8]/10 -> rows trunc(.2*9)=1 .. trunc(.8*9)=7,
    # cols trunc(4/12*11)=3 .. trunc(8/12*11)=7
    det = np.array([[[6, 5, 4, 6, 1, 90], [-1, -1, -1, -1, -1, -100]]], dtype=np.int32)
    out = do.draw_boxes(img, det)
    want = img.copy()
    want[0, 1, 3:8] = 255
    want[0, 7, 3:8] = 255
    want[0, 1:8, 3] = 255
    want[0, 1:8, 7] = 255
    want[0, 0, 0] = 255            # the padding row: max(-1, 0) = 0 -> a one-pixel box at the origin
    assert np.array_equal(out, want)
    # a corner in (-1, 0) pixels: C++ truncation toward zero makes it 0, so the top/left lines ARE drawn
    det2 = np.array([[[1, 1, 3, 3, 0, 90]]], dtype=np.int32)
    out2 = do.draw_boxes(img, det2)
    want2 = img.copy()
    want2[0, 0, 0:3] = 255
    want2[0, 2, 0:3] = 255
    want2[0, 0:3, 0] = 255
    want2[0, 0:3, 2] = 255
    assert np.array_equal(out2, want2)
    # a corner beyond -1 pixel stays negative: its line is outside and not drawn
    det3 = np.array([[[1, 1, 6, 6, 0, 90]]], dtype=np.int32)
    out3 = do.draw_boxes(img, det3)
    want3 = img.copy()
    want3[0, 3, 0:4] = 255
    want3[0, 0:4, 3] = 255
    assert np.array_equal(out3, want3)
