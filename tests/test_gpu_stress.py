"""Stress-size parity on the GPU (BASELINE.json configs[2] and [3]): ~3 000 candidates per
(image, class) out of N = 163,680 anchors, 6 classes, threshold 0.05, nms_max_output_size = 1000 - the
NMS cap binds, the radix-select path runs, the mask head sees ~1000 RoIs per image over P3-P5 and
1000 masks per image are pasted to 1024x512 - against the C restatement of the reference path
(oracle/c) on two complete frames."""
import numpy as np
import pytest
import torch

import synth
from oracle import c_oracle as co

pytestmark = pytest.mark.gpu

KW = dict(min_confidence=0.05, nms_iou_threshold=0.4, post_iou_threshold=0.65, nms_max_output_size=1000,
          max_k=2, base_size=64)


@pytest.mark.parametrize("ratios,Cf", [(synth.DEFAULT_RATIOS, 128), (synth.A9_RATIOS, 256)])
def test_stress_frames_equal_c_oracle(ratios, Cf):
    import masklab_b200 as ml
    B, H, W, C = 2, 512, 1024, 6
    cfgp = synth.prior_config(ratios=ratios)
    N = synth.num_anchors(cfgp, H, W)
    assert N in (163680, 98208)
    loc, cls = synth.head_tensors(B, N, C, mu=-5.0, seed=4321)
    fmaps = synth.fpn_maps(B, H, W, Cf, seed=4322)
    pipe = ml.PostProcessPipeline(cfgp, (H, W), (H, W), C, Cf, B, ml.DetectionConfig(**KW))
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    rois = pipe.detect_and_align(d(loc), d(cls), [d(f) for f in fmaps])
    crops, roi_boxes = pipe.roi_views(rois)
    R = roi_boxes.shape[1]
    probs = synth.mask_probs(B, R, C, seed=4323)
    pipe.trim_and_paste(rois, d(probs))
    det_i, pasted = pipe.result_views()
    M = int(rois.m_dev.item())
    want = co.full_path(loc, cls, fmaps, lambda f, b: probs, cfgp, (H, W), (H, W), binary=True, **KW)
    counts = rois.counts.cpu().numpy()
    assert counts.max() >= 900, counts                    # the stress regime: close to / at the cap
    assert np.array_equal(rois.det[:, :M].cpu().numpy(), want["proposed"])
    assert np.array_equal(roi_boxes.cpu().numpy(), want["roi_boxes"])
    for a, b in zip(crops, want["roi_fmaps"]):
        assert np.array_equal(a.cpu().numpy(), b)
    assert np.array_equal(det_i.cpu().numpy(), want["det_i"])
    assert np.array_equal(pasted.cpu().numpy(), want["binary"])


def test_streaming_shape_frames_equal_c_oracle():
    """BASELINE.json configs[4] shapes: model at 540x960 (N = 163,275), masks pasted at 1080x1920."""
    import masklab_b200 as ml
    B, H, W, PH, PW, C, Cf = 2, 540, 960, 1080, 1920, 6, 128
    kw = dict(min_confidence=0.05, nms_iou_threshold=0.4, post_iou_threshold=0.65, nms_max_output_size=100,
              max_k=2, base_size=36)
    cfgp = synth.prior_config()
    N = synth.num_anchors(cfgp, H, W)
    assert N == 163275
    loc, cls = synth.head_tensors(B, N, C, mu=-5.8, seed=555)
    fmaps = synth.fpn_maps(B, H, W, Cf, seed=556)
    pipe = ml.PostProcessPipeline(cfgp, (H, W), (PH, PW), C, Cf, B, ml.DetectionConfig(**kw))
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    rois = pipe.detect_and_align(d(loc), d(cls), [d(f) for f in fmaps])
    crops, roi_boxes = pipe.roi_views(rois)
    probs = synth.mask_probs(B, roi_boxes.shape[1], C, seed=557)
    pipe.trim_and_paste(rois, d(probs))
    det_i, pasted = pipe.result_views()
    M = int(rois.m_dev.item())
    want = co.full_path(loc, cls, fmaps, lambda f, b: probs, cfgp, (H, W), (PH, PW), binary=True, **kw)
    assert np.array_equal(rois.det[:, :M].cpu().numpy(), want["proposed"])
    for a, b in zip(crops, want["roi_fmaps"]):
        assert np.array_equal(a.cpu().numpy(), b)
    assert np.array_equal(det_i.cpu().numpy(), want["det_i"])
    assert np.array_equal(pasted.cpu().numpy(), want["binary"])
