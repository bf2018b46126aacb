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


def resize_bilinear_nhwc(x, out_h, out_w):
    """x [B,h,w,S] (any numeric, cast to f32) -> [B,out_h,out_w,S] f32, align_corners=True."""
    x = np.asarray(x).astype(F32)
    B, h, w, S = x.shape
    ylo, yhi, yl = tf_ops.resize_interp_weights(out_h, h)
    xlo, xhi, xl = tf_ops.resize_interp_weights(out_w, w)
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
