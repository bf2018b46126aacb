"""Randomised parity sweep on the GPU: odd shapes and unusual hyper-parameters through the fused
and the staged pipeline, checked bit for bit against the C restatement of the reference path
(oracle/c, itself proven identical to the NumPy oracle in tests/test_c_oracle.py)."""
import numpy as np
import pytest
import torch

import synth
from oracle import c_oracle as co
from oracle import masklab_oracle as mo

pytestmark = pytest.mark.gpu
F32 = np.float32


def _case(seed):
    rng = np.random.default_rng(seed)
    strides = [(8, 16, 32, 64, 128), (8, 16, 32, 64), (4, 8, 16), (16, 32, 64)][rng.integers(4)]
    ratios = [synth.DEFAULT_RATIOS, synth.A9_RATIOS, (1 / 2, 1, 2, 5, 8), (1,)][rng.integers(4)]
    scales = [synth.DEFAULT_SCALES, (1,), (1, 1.5)][rng.integers(3)]
    H = int(rng.integers(40, 200))
    W = int(rng.integers(40, 260))
    B = int(rng.integers(1, 6))
    C = int(rng.integers(1, 8))
    Cf = int(rng.choice([3, 4, 8, 20, 32, 130]))
    max_k = int(rng.integers(0, min(3, len(strides))))
    crop = [(14, 14), (7, 7), (1, 1), (5, 9), (2, 3)][rng.integers(5)]
    mask = [(28, 28), (14, 14), (7, 9), (33, 35), (32, 32), (1, 1)][rng.integers(6)]
    ratio = [1.0, 2.0, 1.5][rng.integers(3)]
    PH, PW = int(H * ratio), int(W * ratio)
    if rng.random() < 0.5:
        PW = (PW // 16) * 16 or 16                                  # vector paste path
    kw = dict(min_confidence=float(rng.choice([0.05, 0.3, 0.0, 0.5])),
              nms_iou_threshold=float(rng.choice([0.4, 0.1, 0.9, 0.0])),
              post_iou_threshold=float(rng.choice([0.65, 0.2, 1.0])),
              nms_max_output_size=int(rng.choice([1, 7, 50, 300])),
              max_k=max_k, base_size=float(rng.choice([36, 64, 10])))
    mu = float(rng.choice([-5.0, -3.0, -1.5]))
    padding = ["same", "valid"][rng.integers(2)]
    return dict(strides=strides, ratios=ratios, scales=scales, H=H, W=W, B=B, C=C, Cf=Cf, crop=crop,
                mask=mask, PH=PH, PW=PW, kw=kw, mu=mu, padding=padding)


@pytest.mark.parametrize("seed", list(range(24)))
def test_random_configuration_matches_c_oracle(seed):
    import masklab_b200 as ml
    c = _case(1000 + seed)
    cfgp = synth.prior_config(strides=c["strides"], scales=c["scales"], ratios=c["ratios"])
    B, H, W, C, Cf = c["B"], c["H"], c["W"], c["C"], c["Cf"]
    N = synth.num_anchors(cfgp, H, W, c["padding"])
    if N == 0:
        pytest.skip("no anchors")
    if c["kw"]["min_confidence"] == 0.0 and B * N * C > 60000:
        c["kw"]["min_confidence"] = 0.05                            # keep the oracle's NMS affordable
    loc, cls = synth.head_tensors(B, N, C, mu=c["mu"], seed=seed)
    if seed % 3 == 0 and B > 1:
        cls[B - 1] = 0
    fmaps = synth.fpn_maps(B, H, W, Cf, strides=c["strides"][:c["kw"]["max_k"] + 1], seed=seed + 1,
                           padding=c["padding"])
    probs = {}

    def head(roi_fmaps, roi_boxes):
        probs["m"] = synth.mask_probs(B, roi_boxes.shape[1], C, mask_hw=c["mask"], seed=seed + 2)
        return probs["m"]

    want = co.full_path(loc, cls, fmaps, head, cfgp, (H, W), (c["PH"], c["PW"]), crop_size=c["crop"],
                        padding=c["padding"], binary=True, **c["kw"])
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    # fused / staged chain, then the round-2 variants of the fused tail: background fill on a second stream
    # (prefill; silently off for widths the vector paste cannot take) and the planar mask-head layout
    for fused, prefill, planar in ((True, False, False), (False, False, False), (True, True, False),
                                   (True, False, True), (True, True, True)):
        cfg = ml.DetectionConfig(crop_size=c["crop"], mask_size=c["mask"], padding=c["padding"],
                                 fused=fused, prefill=prefill, mask_layout="planar" if planar else "interleaved",
                                 **c["kw"])
        pipe = ml.PostProcessPipeline(cfgp, (H, W), (c["PH"], c["PW"]), C, Cf, B, cfg)
        pipe.pasted.fill_(7)                                           # stale bytes the paste (or the fill) has to clear
        rois = pipe.detect_and_align(d(loc), d(cls), [d(f) for f in fmaps])
        crops, roi_boxes = pipe.roi_views(rois)
        M = int(rois.m_dev.item())
        assert np.array_equal(rois.det[:, :M].cpu().numpy(), want["proposed"]), (seed, fused, c)
        assert np.array_equal(roi_boxes.cpu().numpy(), want["roi_boxes"]), (seed, fused, c)
        for g, w in zip(crops, want["roi_fmaps"]):
            assert np.array_equal(g.cpu().numpy(), w), (seed, fused, c)
        masks = probs["m"].transpose(0, 1, 4, 2, 3) if planar else probs["m"]
        pipe.trim_and_paste(rois, d(masks))
        det_i, pasted = pipe.result_views()
        assert np.array_equal(det_i.cpu().numpy(), want["det_i"]), (seed, fused, prefill, planar, c)
        assert np.array_equal(pasted.cpu().numpy(), want["binary"]), (seed, fused, prefill, planar, c)


def test_negative_and_nonfinite_scores_are_handled():
    """Scores outside (0,1): negative thresholds keep negative scores (ordered-key transform),
    NaN never passes `>=`, +inf sorts first."""
    import masklab_b200 as ml
    rng = np.random.default_rng(5)
    B, N, C = 2, 400, 3
    cls = rng.standard_normal((B, N, C)).astype(F32)
    cls[0, 3, 1] = np.inf
    cls[0, 5, :] = np.nan
    cls[1, 7, 2] = -0.0
    boxes = np.concatenate([rng.uniform(0, 200, (B, N, 2)), rng.uniform(5, 60, (B, N, 2))], -1).astype(F32)
    kw = dict(min_confidence=-0.5, nms_iou_threshold=0.3, post_iou_threshold=0.5, nms_max_output_size=40)
    want = mo.detection_proposal(cls, boxes, **kw)
    got = ml.DetectionProposal(**kw)([torch.from_numpy(cls).cuda(), torch.from_numpy(boxes).cuda(), None])
    assert np.array_equal(got.cpu().numpy(), want)
    assert want[0, 0, 5] == np.inf
