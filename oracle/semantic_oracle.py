"""NumPy restatement of the resize steps either side of the path (resize
part): DownSampleInput and the semantic half of UpSampleOutput.  TEST INFRASTRUCTURE — see
oracle/__init__.py ("parity unpinned").

Reference: /root/reference/engine/layers/misc.py:133-161 (DownSampleInput), :190-195 (UpSampleOutput,
semantic branch); both call tf.compat.v1.image.resize_bilinear(align_corners=True), the legacy
ResizeBilinear kernel restated in oracle/tf_ops.py (pinned there against TensorFlow's published
resize_bilinear_op_test vectors).
"""
import numpy as np

from . import tf_ops

F32 = np.float32


def resize_bilinear_nhwc(x, out_h, out_w, align_corners=True):
    """x [B,h,w,S] (any numeric, cast to f32) -> [B,out_h,out_w,S] f32 (ResizeLike.call, misc.py:302-307,
    with the layer's align_corners; True everywhere else on the path)."""
    x = np.asarray(x).astype(F32)
    B, h, w, S = x.shape
    ylo, yhi, yl = tf_ops.resize_interp_weights(out_h, h, align_corners)
    xlo, xhi, xl = tf_ops.resize_interp_weights(out_w, w, align_corners)
    tl = x[:, ylo][:, :, xlo]
    tr = x[:, ylo][:, :, xhi]
    bl = x[:, yhi][:, :, xlo]
    br = x[:, yhi][:, :, xhi]
    xl = xl[None, None, :, None]
    yl = yl[None, :, None, None]
    top = tl + (tr - tl) * xl
    bot = bl + (br - bl) * xl
    return (top + (bot - top) * yl).astype(F32)


def downsample_input(inputs, target_size=(540, 960)):
    """DownSampleInput.call, misc.py:143-154: the frame resized with the smaller of the two ratios
    (aspect kept), output size truncated to int32."""
    x = np.asarray(inputs).astype(F32)
    ih, iw = F32(x.shape[1]), F32(x.shape[2])
    ratio = min(F32(target_size[0]) / ih, F32(target_size[1]) / iw)
    th, tw = int(np.int32(ratio * ih)), int(np.int32(ratio * iw))
    return resize_bilinear_nhwc(x, th, tw)


def upsample_semantic(semantic_output, dst_hw):
    """UpSampleOutput.call, misc.py:190-195: resize to the target frame, then > 0.5 -> int32."""
    up = resize_bilinear_nhwc(semantic_output, int(dst_hw[0]), int(dst_hw[1]))
    return (up > F32(0.5)).astype(np.int32)


def _window_reduce(x, k, axis, op, fill):
    """TF dilation2d window along one axis with padding='SAME', stride 1, rate 1: output position p
    reads input positions p - (k-1)//2 ... p - (k-1)//2 + k - 1; positions outside are skipped."""
    pt = (k - 1) // 2
    n = x.shape[axis]
    out = np.full_like(x, fill)
    for d in range(k):
        off = d - pt
        lo, hi = max(0, -off), min(n, n - off)               # output range whose source p+off is inside
        if hi <= lo:
            continue
        dst = [slice(None)] * x.ndim
        src = [slice(None)] * x.ndim
        dst[axis] = slice(lo, hi)
        src[axis] = slice(lo + off, hi + off)
        out[tuple(dst)] = op(out[tuple(dst)], x[tuple(src)])
    return out


def semantic_smoothing(inputs, kernel_size=10, weight=1.0):
    """SemanticSmoothing.call, engine/layers/semantic.py:270-284: grey erosion then dilation with a
    flat (all-zero) kernel_size x kernel_size structuring element, padding 'SAME', times weight.
    tf.nn.erosion2d(v, k) = -dilation2d(-v, reverse(k)) - with a zero kernel both read the SAME
    window, rows/cols [p - (k-1)//2, p - (k-1)//2 + k - 1] (for the default k = 10: p-4 .. p+5), and
    positions outside the map are skipped.  min/max are exact, so the 2-D windows are evaluated
    separably."""
    x = np.asarray(inputs).astype(F32)
    if kernel_size > 0:
        e = _window_reduce(x, kernel_size, 1, np.minimum, F32(np.inf))
        e = _window_reduce(e, kernel_size, 2, np.minimum, F32(np.inf))
        d = _window_reduce(e, kernel_size, 1, np.maximum, F32(-np.inf))
        x = _window_reduce(d, kernel_size, 2, np.maximum, F32(-np.inf))
    return (x * F32(weight)).astype(F32)
