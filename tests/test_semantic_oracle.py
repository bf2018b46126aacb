"""CPU pins of oracle/semantic_oracle.py (resize steps either side of the path)."""
import numpy as np

from oracle import semantic_oracle as so
from oracle import tf_ops

F32 = np.float32


def test_resize_nhwc_reproduces_tf_published_vectors_per_channel():
    two = np.array([[1, 2], [3, 4]], F32)
    four = np.arange(1, 17, dtype=F32).reshape(4, 4)
    x = np.stack([np.pad(two, ((0, 2), (0, 2))), four], axis=-1)[None]          # [1,4,4,2]
    out = so.resize_bilinear_nhwc(x[:, :, :, 1:], 3, 3)
    assert out[0, :, :, 0].ravel().tolist() == [1, 2.5, 4, 7, 8.5, 10, 13, 14.5, 16]
    out2 = so.resize_bilinear_nhwc(two[None, :, :, None], 3, 3)
    assert out2[0, :, :, 0].ravel().tolist() == [1, 1.5, 2, 2, 2.5, 3, 3, 3.5, 4]
    rng = np.random.default_rng(0)
    y = rng.random((2, 7, 9, 3)).astype(F32)
    got = so.resize_bilinear_nhwc(y, 13, 5)
    for b in range(2):
        for c in range(3):
            assert np.array_equal(got[b, :, :, c], tf_ops.resize_bilinear_align_corners(y[b, :, :, c], 13, 5))


def test_downsample_input_size_rule():
    x = np.zeros((1, 1080, 1920, 3), np.uint8)
    assert so.downsample_input(x, (540, 960)).shape == (1, 540, 960, 3)
    assert so.downsample_input(np.zeros((1, 100, 100, 3), np.uint8), (54, 96)).shape == (1, 54, 54, 3)   # min ratio
    assert so.downsample_input(np.zeros((1, 50, 333, 1), np.uint8), (54, 96)).shape == (1, 14, 96, 1)    # truncated
    ramp = np.tile(np.arange(5, dtype=np.uint8)[None, None, :, None] * 10, (1, 3, 1, 1))
    assert so.downsample_input(ramp, (3, 3))[0, 0, :, 0].tolist() == [0, 20, 40]                        # corners kept


def test_upsample_semantic_thresholds_after_the_resize():
    sem = np.zeros((1, 2, 2, 1), F32)
    sem[0, 0, 0, 0] = 1.0
    up = so.upsample_semantic(sem, (3, 3))
    assert up.dtype == np.int32
    assert up[0, :, :, 0].tolist() == [[1, 0, 0], [0, 0, 0], [0, 0, 0]]        # 0.5 is not > 0.5
    sem[0, 0, 1, 0] = 0.2
    assert so.upsample_semantic(sem, (3, 3))[0, 0, :, 0].tolist() == [1, 1, 0]  # (1 + 0.2) / 2 = 0.6


def test_semantic_smoothing_known_answers():
    # an isolated speck is erased by the erosion; a block at least k wide survives the opening.
    # Erosion and dilation read the SAME even-sized window p-4 .. p+5 (the duality only reverses the
    # kernel values, all zero here), so the surviving block comes back shifted by one pixel up/left.
    x = np.zeros((1, 30, 30, 1), F32)
    x[0, 5, 5, 0] = 1.0                     # noise
    x[0, 10:22, 8:20, 0] = 0.8              # 12 x 12 block
    out = so.semantic_smoothing(x, 10, 1.0)
    assert out[0, 5, 5, 0] == 0 and out[0, :9].sum() == 0
    assert np.array_equal(out[0, 9:21, 7:19, 0], np.full((12, 12), F32(0.8)))
    assert np.isclose(out.sum(), 0.8 * 144)  # same size, nothing else
    # window alignment of the even-sized SAME kernel: rows p-4 .. p+5.  Erosion of a step edge
    # (ones from row 10 on) keeps ones where the whole window is inside the ones: p-4 >= 10.
    step = np.zeros((1, 30, 3, 1), F32)
    step[0, 10:, :, 0] = 1.0
    e = so._window_reduce(step, 10, 1, np.minimum, F32(np.inf))
    assert e[0, :, 0, 0].tolist() == [0.0] * 14 + [1.0] * 16
    d = so._window_reduce(e, 10, 1, np.maximum, F32(-np.inf))
    assert d[0, :, 0, 0].tolist() == [0.0] * 9 + [1.0] * 21    # dilation reaches back to p+5 >= 14 -> p >= 9
    # kernel_size 0: only the weight (semantic.py:283-284)
    y = np.random.default_rng(0).random((1, 4, 5, 2)).astype(F32)
    assert np.array_equal(so.semantic_smoothing(y, 0, 0.5), y * F32(0.5))
    # brute force on a small random map
    z = np.random.default_rng(1).random((2, 9, 11, 3)).astype(F32)
    got = so.semantic_smoothing(z, 4, 2.0)
    pt = 1
    er = np.empty_like(z)
    for yy in range(9):
        for xx in range(11):
            er[:, yy, xx] = z[:, max(0, yy - pt):yy - pt + 4, max(0, xx - pt):xx - pt + 4].min(axis=(1, 2))
    di = np.empty_like(z)
    for yy in range(9):
        for xx in range(11):
            di[:, yy, xx] = er[:, max(0, yy - pt):yy - pt + 4, max(0, xx - pt):xx - pt + 4].max(axis=(1, 2))
    assert np.array_equal(got, di * F32(2.0))
