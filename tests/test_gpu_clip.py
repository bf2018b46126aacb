"""Box-clipped masks (csrc/clip.cu, PostProcessPipeline.trim_and_clip): per instance its clipped box and the bit rows
of (CropAndPadMask value > 0.5) inside it.  The host-side expander must rebuild the dense [B,M,PH,PW] binary masks of
the C restatement of the reference path bit for bit - also for boxes hanging over the frame edges, frame-sized boxes,
filtered rows, images without detections, odd frame widths and the planar mask-head layout."""
import numpy as np
import pytest
import torch

import synth
from oracle import c_oracle as co

pytestmark = pytest.mark.gpu


def _d(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _run(cfgp, loc, cls, fmaps, hw, frame, C, Cf, kw, planar=False, mask_hw=(28, 28), seed=9, pool_bytes=None):
    import masklab_b200 as ml
    B = loc.shape[0]
    cfg = ml.DetectionConfig(mask_size=mask_hw, mask_layout="planar" if planar else "interleaved", **kw)
    pipe = ml.PostProcessPipeline(cfgp, hw, frame, C, Cf, B, cfg)
    rois = pipe.detect_and_align(_d(loc), _d(cls), [_d(f) for f in fmaps])
    _, R = rois.shapes()
    probs = synth.mask_probs(B, R, C, mask_hw=mask_hw, seed=seed)
    det_i, geom, pool, used = pipe.trim_and_clip(rois, _d(probs.transpose(0, 1, 4, 2, 3) if planar else probs),
                                                 pool_bytes=pool_bytes)
    torch.cuda.synchronize()
    M = int(pipe.trim_m.item())
    want = co.full_path(loc, cls, fmaps, lambda f, b: probs, cfgp, hw, frame, binary=True, **kw)
    return pipe, ml, M, det_i.cpu().numpy(), geom.cpu().numpy(), pool.cpu().numpy(), used.cpu().numpy(), want


@pytest.mark.parametrize("case", [
    dict(B=3, H=96, W=160, PH=192, PW=320, C=4, Cf=8, mu=-3.0, K=40, planar=False),
    dict(B=2, H=96, W=160, PH=96, PW=163, C=3, Cf=8, mu=-3.0, K=40, planar=True),       # odd frame width
    dict(B=2, H=128, W=256, PH=256, PW=512, C=5, Cf=16, mu=-2.0, K=300, planar=False),   # no bit tiles (K > 256)
    dict(B=1, H=64, W=96, PH=64, PW=96, C=2, Cf=4, mu=-6.5, K=10, planar=False),          # hardly any detection
])
def test_expanded_clip_equals_dense_masks(case):
    c = case
    cfgp = synth.prior_config(strides=(8, 16, 32, 64))
    N = synth.num_anchors(cfgp, c["H"], c["W"])
    loc, cls = synth.head_tensors(c["B"], N, c["C"], mu=c["mu"], seed=11)
    if c["B"] > 1:
        cls[1] = 0
    fmaps = synth.fpn_maps(c["B"], c["H"], c["W"], c["Cf"], seed=12)
    kw = dict(min_confidence=0.05, nms_iou_threshold=0.4, post_iou_threshold=0.65, nms_max_output_size=c["K"],
              max_k=2, base_size=36)
    pipe, ml, M, det_i, geom, pool, used, want = _run(cfgp, loc, cls, fmaps, (c["H"], c["W"]), (c["PH"], c["PW"]),
                                                      c["C"], c["Cf"], kw, planar=c["planar"])
    assert np.array_equal(det_i.reshape(c["B"], c["K"], 6)[:, :M], want["det_i"])
    dense = ml.expand_clipped(geom, pool, (c["PH"], c["PW"]), m_rows=M)
    assert dense.shape == want["binary"].shape and np.array_equal(dense, want["binary"])
    # the pool holds exactly the clipped boxes: sum of h * ceil(w/8), far below the dense tensor
    need = int(sum(int(g[3]) * ((int(g[2]) + 7) // 8) for g in geom.reshape(-1, 8)))
    assert int(used[0]) == need and need <= want["binary"].size // 8 + geom.shape[0] * geom.shape[1]
    assert np.all(geom[:, M:, 2:4] == 0)                       # capacity rows past M hold nothing


def test_clip_pool_overflow_is_reported_not_written():
    cfgp = synth.prior_config(strides=(8, 16, 32))
    B, H, W, C, Cf = 2, 96, 160, 3, 8
    N = synth.num_anchors(cfgp, H, W)
    loc, cls = synth.head_tensors(B, N, C, mu=-3.0, seed=21)
    fmaps = synth.fpn_maps(B, H, W, Cf, seed=22)
    kw = dict(min_confidence=0.05, nms_iou_threshold=0.4, post_iou_threshold=0.65, nms_max_output_size=30,
              max_k=2, base_size=36)
    pipe, ml, M, det_i, geom, pool, used, want = _run(cfgp, loc, cls, fmaps, (H, W), (H, W), C, Cf, kw, pool_bytes=64)
    assert int(used[0]) > 64                                   # the caller sees that 64 bytes were not enough
    assert pool.shape[0] >= 64


def test_clip_full_size_cfg2():
    """The whole cfg-2 batch: 32 frames, 100 instances each - 2.4 MB of clipped rows against 1.68 GB dense."""
    import masklab_b200 as ml
    B, H, W, C, Cf = 32, 512, 1024, 5, 128
    cfgp = synth.prior_config()
    N = synth.num_anchors(cfgp, H, W)
    loc, cls = synth.head_tensors(B, N, C, mu=-5.8, seed=4321)
    fmaps = synth.fpn_maps(B, H, W, Cf, seed=4322)
    kw = dict(min_confidence=0.05, nms_iou_threshold=0.4, post_iou_threshold=0.6, nms_max_output_size=100,
              max_k=2, base_size=36)
    pipe = ml.PostProcessPipeline(cfgp, (H, W), (H, W), C, Cf, B, ml.DetectionConfig(**kw))
    ins = (_d(loc), _d(cls), [_d(f) for f in fmaps])
    rois = pipe.detect_and_align(*ins)
    _, R = rois.shapes()
    probs = _d(synth.mask_probs(B, R, C, seed=4323))
    pipe.trim_and_paste(rois, probs)
    _, dense = pipe.result_views()
    dense = dense.cpu().numpy()
    det_i, geom, pool, used = pipe.trim_and_clip(rois, probs)
    torch.cuda.synchronize()
    n = int(used[0])
    assert n < 8 << 20                                         # a few MB
    M = int(pipe.trim_m.item())
    got = ml.expand_clipped(geom.cpu().numpy(), pool[:n].cpu().numpy(), (H, W), m_rows=M)
    assert np.array_equal(got, dense)
