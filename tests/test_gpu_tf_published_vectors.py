"""The known-answer vectors published with TensorFlow's own kernel tests (see
tests/test_tf_published_vectors.py for the sources) driven through the CUDA kernels via the drop-in
layers: NonMaxSuppressionV3 behind DetectionProposal, CropAndResize behind PyramidRoiAlign,
ResizeBilinear(align_corners) behind CropAndPadMask."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

F32 = np.float32


def dev(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


CLUSTERS = np.array([[0, 0, 1, 1], [0, 0.1, 1, 1.1], [0, -0.1, 1, 0.9],
                     [0, 10, 1, 11], [0, 10.1, 1, 11.1], [0, 100, 1, 101]], F32)
SCORES = np.array([.9, .75, .6, .95, .5, .3], F32)


def cxcywh(corners):
    y1, x1, y2, x2 = corners.T
    return np.stack([(x1 + x2) / 2, (y1 + y2) / 2, x2 - x1, y2 - y1], 1).astype(F32)


@pytest.mark.parametrize("max_out,want", [(3, [3, 0, 5]), (2, [3, 0]), (30, [3, 0, 5])])
def test_nms_three_clusters_through_detection_proposal(max_out, want):
    import masklab_b200 as ml
    boxes = cxcywh(CLUSTERS)[None]
    cls = SCORES[None, :, None]
    layer = ml.DetectionProposal(min_confidence=0.05, nms_iou_threshold=0.5, post_iou_threshold=1.0,
                                 nms_max_output_size=max_out)
    out = layer([dev(cls), dev(boxes), None]).cpu().numpy()
    assert out.shape == (1, len(want), 6)
    assert np.array_equal(out[0, :, 5], SCORES[want]) and np.all(out[0, :, 4] == 0)
    assert np.array_equal(out[0, :, :4], boxes[0][want])


def roi_align(img, corners, crop, channels):
    import masklab_b200 as ml
    img = np.asarray(img, F32)
    fmap = np.repeat(img[None, :, :, None], channels, axis=3)
    boxes = cxcywh(np.asarray(corners, F32))
    n = boxes.shape[0]
    dist = np.concatenate([np.zeros((n, 1), F32), boxes, np.zeros((n, 1), F32), np.ones((n, 1), F32)], 1)[None]
    images = torch.zeros((1, 1, 1, 3), device="cuda")                 # boxes are already normalised: H = W = 1
    crops, _ = ml.PyramidRoiAlign(crop_size=crop)([[dev(fmap)], dev(dist), images])
    out = crops[0].cpu().numpy()[0]                                    # [n, ch, cw, channels]
    assert all(np.array_equal(out[..., 0], out[..., c]) for c in range(channels))
    return out[..., 0]


@pytest.mark.parametrize("channels", [1, 4, 128])
def test_crop_and_resize_published_vectors(channels):
    two = [[1, 2], [3, 4]]
    assert roi_align(two, [[0, 0, 1, 1]], (1, 1), channels).ravel().tolist() == [2.5]
    assert roi_align(two, [[0, 0, 1, 1]], (3, 3), channels).ravel().tolist() == [1, 1.5, 2, 2, 2.5, 3, 3, 3.5, 4]
    assert roi_align(two, [[1, 1, 0, 0]], (3, 3), channels).ravel().tolist() == [4, 3.5, 3, 3, 2.5, 2, 2, 1.5, 1]
    three = np.arange(1, 10, dtype=F32).reshape(3, 3)
    got = roi_align(three, [[0, 0, 1, 1], [0, 0, .5, .5]], (2, 2), channels)
    assert got[0].ravel().tolist() == [1, 3, 7, 9] and got[1].ravel().tolist() == [1, 2, 4, 5]
    got = roi_align(three, [[1, 1, 0, 0], [.5, .5, 0, 0]], (2, 2), channels)
    assert got[0].ravel().tolist() == [9, 7, 3, 1] and got[1].ravel().tolist() == [5, 4, 2, 1]
    # extrapolation: the layer's extrapolation_value is 0 (instance.py:125) - the published vector with 0
    got = roi_align(two, [[-1, -1, 1, 1]], (3, 3), channels)
    assert got.ravel().tolist() == [0, 0, 0, 0, 1, 2, 0, 3, 4]


def test_resize_bilinear_align_corners_published_vectors_through_paste():
    import masklab_b200 as ml
    for tile, want in (([[1, 2], [3, 4]], [1, 1.5, 2, 2, 2.5, 3, 3, 3.5, 4]),
                       (np.arange(1, 17).reshape(4, 4), [1, 2.5, 4, 7, 8.5, 10, 13, 14.5, 16])):
        ins = np.asarray(tile, np.int32)[None, None]
        det = np.array([[[2, 2, 3, 3, 0, 90]]], np.int32)            # box rows/cols [1,4): 3 x 3 pixels
        out = ml.CropAndPadMask()([(6, 8), dev(det), dev(ins)]).cpu().numpy()[0, 0]
        assert out[1:4, 1:4].ravel().tolist() == want
        out[1:4, 1:4] = 0
        assert not out.any()
