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
