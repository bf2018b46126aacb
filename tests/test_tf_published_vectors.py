"""Pins of oracle/tf_ops.py against known-answer vectors PUBLISHED with TensorFlow itself - the unit
tests of the kernels the reference calls on this path (TensorFlow 1.14/1.15 source tree, not
available offline; values transcribed):

  tensorflow/core/kernels/non_max_suppression_op_test.cc   (NonMaxSuppressionV3, detection.py:507,542)
  tensorflow/core/kernels/crop_and_resize_op_test.cc       (CropAndResize, instance.py:125)
  tensorflow/core/kernels/resize_bilinear_op_test.cc       (ResizeBilinear align_corners, misc.py:389)

The reference repository itself holds no golden vectors for the path ("parity unpinned");
these are the closest published fixtures of the third-party arithmetic it delegates to.
"""
import numpy as np

from oracle import tf_ops

F32 = np.float32

# --- NonMaxSuppressionOpTest: six boxes in three clusters -----------------------------------
CLUSTERS = np.array([[0, 0, 1, 1], [0, 0.1, 1, 1.1], [0, -0.1, 1, 0.9],
                     [0, 10, 1, 11], [0, 10.1, 1, 11.1], [0, 100, 1, 101]], F32)
SCORES = np.array([.9, .75, .6, .95, .5, .3], F32)


def test_nms_select_from_three_clusters():
    assert tf_ops.non_max_suppression(CLUSTERS, SCORES, 3, 0.5).tolist() == [3, 0, 5]


def test_nms_select_from_three_clusters_flipped_coordinates():
    flipped = np.array([[1, 1, 0, 0], [0, 0.1, 1, 1.1], [0, .9, 1, -0.1],
                        [0, 10, 1, 11], [1, 10.1, 0, 11.1], [1, 101, 0, 100]], F32)
    assert tf_ops.non_max_suppression(flipped, SCORES, 3, 0.5).tolist() == [3, 0, 5]


def test_nms_select_at_most_two_boxes_from_three_clusters():
    assert tf_ops.non_max_suppression(CLUSTERS, SCORES, 2, 0.5).tolist() == [3, 0]


def test_nms_select_with_negative_scores():
    assert tf_ops.non_max_suppression(CLUSTERS, SCORES - F32(10.0), 6, 0.5).tolist() == [3, 0, 5]


def test_nms_first_box_degenerate():
    boxes = np.array([[0, 0, 0, 0], [1, 1, 2, 2], [2, 2, 3, 3]], F32)
    assert tf_ops.non_max_suppression(boxes, np.array([.9, .75, .6], F32), 3, 0.5).tolist() == [0, 1, 2]


def test_nms_select_at_most_thirty_boxes_from_three_clusters():
    assert tf_ops.non_max_suppression(CLUSTERS, SCORES, 30, 0.5).tolist() == [3, 0, 5]


def test_nms_select_single_box_and_ten_identical_boxes():
    assert tf_ops.non_max_suppression(np.array([[0, 0, 1, 1]], F32), np.array([.9], F32), 3, 0.5).tolist() == [0]
    same = np.tile(np.array([[0, 0, 1, 1]], F32), (10, 1))
    assert tf_ops.non_max_suppression(same, np.full(10, .9, F32), 3, 0.5).tolist() == [0]


def test_nms_empty_input():
    assert tf_ops.non_max_suppression(np.zeros((0, 4), F32), np.zeros((0,), F32), 30, 0.5).tolist() == []


# --- CropAndResizeOpTest ---------------------------------------------------------------------
def crop(img, boxes, size, extrapolation=0.0):
    img = np.asarray(img, F32)
    out = tf_ops.crop_and_resize(img.reshape(1, img.shape[0], img.shape[1], 1), np.asarray(boxes, F32),
                                 [0] * len(boxes), size, extrapolation)
    return out[..., 0]


def test_crop_and_resize_2x2_to_1x1():
    assert crop([[1, 2], [3, 4]], [[0, 0, 1, 1]], (1, 1)).ravel().tolist() == [2.5]


def test_crop_and_resize_2x2_to_3x3_and_flipped():
    assert crop([[1, 2], [3, 4]], [[0, 0, 1, 1]], (3, 3)).ravel().tolist() == [1, 1.5, 2, 2, 2.5, 3, 3, 3.5, 4]
    assert crop([[1, 2], [3, 4]], [[1, 1, 0, 0]], (3, 3)).ravel().tolist() == [4, 3.5, 3, 3, 2.5, 2, 2, 1.5, 1]


def test_crop_and_resize_3x3_to_2x2_and_flipped():
    img = np.arange(1, 10, dtype=F32).reshape(3, 3)
    got = crop(img, [[0, 0, 1, 1], [0, 0, .5, .5]], (2, 2))
    assert got[0].ravel().tolist() == [1, 3, 7, 9] and got[1].ravel().tolist() == [1, 2, 4, 5]
    got = crop(img, [[1, 1, 0, 0], [.5, .5, 0, 0]], (2, 2))
    assert got[0].ravel().tolist() == [9, 7, 3, 1] and got[1].ravel().tolist() == [5, 4, 2, 1]


def test_crop_and_resize_2x2_to_3x3_extrapolated():
    got = crop([[1, 2], [3, 4]], [[-1, -1, 1, 1]], (3, 3), extrapolation=-1.0)
    assert got.ravel().tolist() == [-1, -1, -1, -1, 1, 2, -1, 3, 4]


def test_crop_and_resize_no_boxes():
    out = tf_ops.crop_and_resize(np.zeros((1, 2, 2, 1), F32), np.zeros((0, 4), F32), [], (3, 3))
    assert out.shape == (0, 3, 3, 1)


# --- ResizeBilinearOpAlignCornersTest --------------------------------------------------------
def test_resize_bilinear_align_corners_published():
    two = np.array([[1, 2], [3, 4]], F32)
    assert tf_ops.resize_bilinear_align_corners(two, 1, 1).ravel().tolist() == [1]
    assert tf_ops.resize_bilinear_align_corners(two, 3, 3).ravel().tolist() == [1, 1.5, 2, 2, 2.5, 3, 3, 3.5, 4]
    three = np.arange(1, 10, dtype=F32).reshape(3, 3)
    assert tf_ops.resize_bilinear_align_corners(three, 2, 2).ravel().tolist() == [1, 3, 7, 9]
    four = np.arange(1, 17, dtype=F32).reshape(4, 4)
    assert tf_ops.resize_bilinear_align_corners(four, 3, 3).ravel().tolist() == [1, 2.5, 4, 7, 8.5, 10, 13, 14.5, 16]
