"""Integration of the drop-in layers in the reference's serving order (retinamasklab.py:613-636 and
road_project/setup/serving.py:29-48) against the same composition of the oracles."""
import numpy as np
import pytest
import torch

import synth
from oracle import draw_oracle as do
from oracle import jpeg_oracle as jo
from oracle import masklab_oracle as mo
from oracle import semantic_oracle as seo
from oracle import summary_oracle as so

pytestmark = pytest.mark.gpu

F32 = np.float32


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("kernels,mask_output", [((0, 0, 0), "float32"), ((4, 6, 0), "float32"), ((0, 0, 0), "uint8")])
def test_serving_outputs_match_the_oracle_composition(kernels, mask_output):
    import masklab_b200 as ml
    B, PH, PW, h, w, C = 2, 216, 384, 108, 192, 5
    frames = np.random.default_rng(1).integers(0, 256, (B, PH, PW, 3)).astype(np.uint8)
    # what the model hands over: RoI boxes / mask-head output (level-major, -1 padded) and semantic probabilities
    det = synth.detections(B, 9, C, h, w, seed=2, lo=8.0, hi=90.0, pad_tail=2)
    dist = mo.mask_distribute(det, 2, 36)
    fmaps = synth.fpn_maps(B, h, w, 4, seed=3)
    _, roi_boxes = mo.pyramid_roi_align(fmaps, dist, (h, w))
    roi_masks = synth.mask_probs(B, roi_boxes.shape[1], C, seed=4)
    hs, ws = h // 4, w // 4
    rng = np.random.default_rng(5)
    seg_pred = np.clip(seo.resize_bilinear_nhwc(rng.random((B, 7, 12, 3)).astype(F32) ** 2, hs, ws) * 1.4, 0, 1).astype(F32)
    cfg = ml.PostProcessConfig(resolution=(h, w), smoothing_kernel_sizes=kernels, smoothing_weights=(1.0, 1.1, 0.9))

    # --- oracle composition
    d_o, i_o = mo.trim_instances(roi_boxes, roi_masks)
    posts = [seo.semantic_smoothing(seg_pred[..., i:i + 1], k, wt) for i, (k, wt) in
             enumerate(zip(cfg.smoothing_kernel_sizes, cfg.smoothing_weights))]
    sem = seo.resize_bilinear_nhwc(np.concatenate(posts, axis=-1), h, w)
    det_i, ins_i = mo.upsample_output(d_o, i_o, (h, w), (PH, PW))
    seg_i = seo.upsample_semantic(sem, (PH, PW))
    pasted = mo.crop_and_pad_mask((PH, PW), det_i, ins_i)
    masks_o = pasted if mask_output == "float32" else (pasted > 0.5).astype(F32)
    vis_o = do.draw_segmentation(do.draw_instance(do.draw_boxes(frames, det_i), det_i, masks_o, cfg.instance_colors,
                                                  cfg.instance_alpha), seg_i, cfg.semantic_colors, cfg.semantic_alpha)
    sum_o = so.summary_output(det_i, seg_i, masks_o, cfg.default_road_size)

    # --- the drop-in layers
    vis, summary, det_outs, ins_outs, seg_outs, contents = ml.serving_outputs(
        dev(frames), (h, w), dev(roi_boxes), dev(roi_masks), dev(seg_pred), cfg, mask_output=mask_output, encode=True)
    assert contents == jo.encode_image_content(vis_o)                  # serving.py:41: the JPEG of frame 0
    assert np.array_equal(det_outs.cpu().numpy(), det_i) and np.array_equal(ins_outs.cpu().numpy(), ins_i)
    assert np.array_equal(seg_outs.cpu().numpy(), seg_i)
    assert np.array_equal(vis.cpu().numpy(), vis_o)
    got = summary.cpu().numpy()
    assert got.shape == sum_o.shape
    assert np.array_equal(got[..., :6], sum_o[..., :6]) and np.array_equal(got[..., 10], sum_o[..., 10])
    np.testing.assert_allclose(got[..., 6:10], sum_o[..., 6:10], rtol=1e-6)


def test_resize_like_both_modes():
    import masklab_b200 as ml
    x = np.random.default_rng(7).random((2, 9, 13, 3)).astype(F32)
    for align in (True, False):
        want = seo.resize_bilinear_nhwc(x, 20, 31, align_corners=align)
        got = ml.ResizeLike(align_corners=align)(dev(x), target=(20, 31))
        assert np.array_equal(got.cpu().numpy(), want)
    assert ml.ResizeLike(False).get_config()["align_corners"] is False
