"""GPU parity of the SummaryOutput family against oracle/summary_oracle.py,
through the C ABI (mlp_road_scan, mlp_summary_output) and the drop-in layers.

Integer-valued columns, the metres-per-pixel table and include_my_road are bit-exact; the four
float reductions are float64 sums rounded to float32 on both sides - equal up to the summation
order of the float64 accumulation, compared with rtol 1e-6."""
import numpy as np
import pytest
import torch

import synth
from oracle import masklab_oracle as mo
from oracle import summary_oracle as so

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def pasted(B, M, C, PH, PW, seed, pad_tail=1):
    det = synth.int_detections(B, M, C, PH, PW, seed=seed, pad_tail=pad_tail)
    rng = np.random.default_rng(seed + 1)
    ins = (rng.random((B, M, 28, 28)) > 0.45).astype(np.int32)
    return det, mo.crop_and_pad_mask((PH, PW), det, ins)


def check_summary(got, want):
    assert got.shape == want.shape
    assert np.array_equal(got[..., :6], want[..., :6])
    assert np.array_equal(got[..., 10], want[..., 10])
    np.testing.assert_allclose(got[..., 6:10], want[..., 6:10], rtol=1e-6, atol=0)


@pytest.mark.parametrize("PH,PW", [(48, 80), (64, 1100), (37, 53)])
def test_unit_lengths_bitexact(PH, PW):
    import masklab_b200 as ml
    from masklab_b200.layers import summary as ls
    B = 3
    seg = synth.semantic_map(B, PH, PW, seed=PH)
    seg[2, :, :, 1] = 0                                         # an image without road
    ctx = ml.Context.get()
    unit, bits, cbits, box = ls.road_scan(ctx, dev(seg), 3.25, True)
    want = np.stack([so.road_unit_lengths(seg[b, :, :, 1]) for b in range(B)])
    assert np.array_equal(unit.cpu().numpy(), want)
    packed = np.packbits(seg[..., 1].astype(np.uint8), axis=-1, bitorder="little")
    gotbits = bits.cpu().numpy().view(np.uint8)[..., :packed.shape[-1]]
    assert np.array_equal(gotbits, packed)
    cpacked = np.packbits((seg[..., 2] != 0).astype(np.uint8), axis=-1, bitorder="little")
    assert np.array_equal(cbits.cpu().numpy().view(np.uint8)[..., :cpacked.shape[-1]], cpacked)
    idx = np.argwhere(seg[..., 2] != 0)
    if idx.size:
        assert box[:4].tolist() == [idx[:, 1].min(), idx[:, 2].min(), idx[:, 1].max(), idx[:, 2].max()]


@pytest.mark.parametrize("PH,PW,M,dtype", [(48, 80, 5, "f32"), (64, 1100, 3, "f32"), (37, 53, 4, "f32"),
                                           (48, 80, 5, "u8"), (40, 1030, 2, "u8")])
def test_summary_output_matches_oracle(PH, PW, M, dtype):
    import masklab_b200 as ml
    B, C = 2, 4
    seg = synth.semantic_map(B, PH, PW, seed=7 + PW)
    seg[0, PH // 2:PH // 2 + 4, 3:PW // 2, 2] = 1                # make sure there is a crack region
    det, masks = pasted(B, M, C, PH, PW, seed=PW)
    if dtype == "u8":
        masks = (masks > 0.5).astype(np.uint8)
    want = so.summary_output(det, seg, masks.astype(np.float32))
    got = ml.SummaryOutput(default_road_size=3.25)([dev(det), dev(seg), dev(masks)])
    assert want.shape[1] == M + 1
    check_summary(got.cpu().numpy(), want)


def test_summary_without_crack_and_without_road():
    import masklab_b200 as ml
    B, M, C, PH, PW = 2, 3, 4, 40, 64
    seg = synth.semantic_map(B, PH, PW, seed=3, crack=False, road=False)
    det, masks = pasted(B, M, C, PH, PW, seed=9)
    want = so.summary_output(det, seg, masks)
    got = ml.SummaryOutput()([dev(det), dev(seg), dev(masks)]).cpu().numpy()
    assert want.shape == (B, M, 11)
    check_summary(got, want)
    # zero-area crack region (one row): conf = 0 -> not appended
    seg[1, 5, 3:20, 2] = 1
    want = so.summary_output(det, seg, masks)
    got = ml.SummaryOutput()([dev(det), dev(seg), dev(masks)]).cpu().numpy()
    assert want.shape == (B, M, 11)
    check_summary(got, want)


def test_member_layers():
    import masklab_b200 as ml
    B, M, C, PH, PW = 2, 4, 4, 56, 96
    seg = synth.semantic_map(B, PH, PW, seed=21)
    det, masks = pasted(B, M, C, PH, PW, seed=22)
    inc = ml.IncludeMyRoad(threshold=0.1)([dev(seg), dev(masks)]).cpu().numpy()
    assert np.array_equal(inc, so.include_my_road(seg, masks, 0.1))
    inc5 = ml.IncludeMyRoad(threshold=0.5)([dev(seg), dev(masks)]).cpu().numpy()
    assert np.array_equal(inc5, so.include_my_road(seg, masks, 0.5))
    size = ml.CalculateInstanceSize(default_road_size=2.5)([dev(seg), dev(masks)]).cpu().numpy()
    np.testing.assert_allclose(size, so.calculate_instance_size(seg, masks, 2.5), rtol=1e-6)
    cdet, cseg = ml.CrackToInstance()(dev(seg[..., 2]))
    wdet, wseg = so.crack_to_instance(seg[..., 2])
    assert np.array_equal(cdet.cpu().numpy(), wdet) and np.array_equal(cseg.cpu().numpy(), wseg)
    e_det, _ = ml.CrackToInstance()(dev(np.zeros((2, 8, 8), dtype=np.int32)))
    assert np.array_equal(e_det.cpu().numpy(), so.crack_to_instance(np.zeros((2, 8, 8), np.int32))[0])
    cfg = ml.SummaryOutput(default_road_size=3.0).get_config()
    assert cfg["default_road_size"] == 3.0 and "SummaryOutput" in ml.get_custom_objects()


def test_summary_rejects_cpu_tensors():
    import masklab_b200 as ml
    det = torch.zeros((1, 1, 6), dtype=torch.int32)
    with pytest.raises(ml.InvalidArgumentError):
        ml.SummaryOutput()([det, torch.zeros((1, 8, 8, 3), dtype=torch.int32), torch.zeros((1, 1, 8, 8))])


@pytest.mark.parametrize("PH,PW,M", [(48, 80, 5), (64, 1100, 3), (37, 53, 4), (300, 420, 6)])
def test_summary_from_tiles_equals_summary_of_pasted_masks(PH, PW, M):
    """The tile path never materialises [B,M,PH,PW]; it must agree with the oracle's summary of the
    pasted float32 masks (and hence with the drop-in layer fed CropAndPadMask's output)."""
    import masklab_b200 as ml
    B, C = 2, 4
    seg = synth.semantic_map(B, PH, PW, seed=31 + PW)
    seg[1, PH // 2:PH // 2 + 3, 2:PW // 3, 2] = 1
    det = synth.int_detections(B, M, C, PH, PW, seed=PW + 1, pad_tail=1)
    det[0, 0] = [PW // 2, PH // 2, 2 * PW, 2 * PH, 1, 99]        # a box clipped by the whole frame
    det[0, 1, 5] = 40                                             # below the batch confidence threshold
    rng = np.random.default_rng(PW)
    ins = (rng.random((B, M, 28, 28)) > 0.45).astype(np.int32)
    masks = mo.crop_and_pad_mask((PH, PW), det, ins)
    want = so.summary_output(det, seg, masks)
    layer = ml.SummaryOutput()
    got = layer.from_tiles([dev(det), dev(seg), dev(ins)]).cpu().numpy()
    check_summary(got, want)
    got2 = layer([dev(det), dev(seg), ml.CropAndPadMask()([(PH, PW), dev(det), dev(ins)])]).cpu().numpy()
    check_summary(got2, want)


def test_summary_from_tiles_narrow_and_wide_boxes():
    """Every lane layout of the box reduction: boxes 1..40 columns wide (8 lanes per row), 33..64 (16), wider (32),
    more than one 128-column chunk, single-row and single-column boxes, boxes hanging over the frame edge."""
    import masklab_b200 as ml
    B, C, PH, PW, M = 2, 4, 150, 400, 24
    rng = np.random.default_rng(77)
    seg = synth.semantic_map(B, PH, PW, seed=78)
    det = np.zeros((B, M, 6), np.int32)
    widths = [1, 2, 7, 8, 9, 16, 17, 31, 32, 33, 47, 63, 64, 65, 100, 127, 128, 129, 200, 300, 5, 12, 40, 70]
    for b in range(B):
        det[b, :, 2] = widths if b == 0 else widths[::-1]
        det[b, :, 3] = rng.integers(1, 140, M)
        det[b, :, 0] = rng.integers(0, PW, M)
        det[b, :, 1] = rng.integers(0, PH, M)
        det[b, :, 4] = rng.integers(0, C, M)
        det[b, :, 5] = rng.integers(55, 100, M)
    det[0, 3, 3] = 1                                                  # one row
    det[1, 5, :2] = [PW - 2, PH - 1]                                  # over the corner
    ins = (rng.random((B, M, 28, 28)) > 0.4).astype(np.int32)
    masks = mo.crop_and_pad_mask((PH, PW), det, ins)
    want = so.summary_output(det, seg, masks)
    got = ml.SummaryOutput().from_tiles([dev(det), dev(seg), dev(ins)]).cpu().numpy()
    check_summary(got, want)


def test_pipeline_trim_and_summarize():
    import masklab_b200 as ml
    B, H, W, C, Cf = 2, 128, 256, 3, 16
    PH, PW = 256, 512
    cfgp = synth.prior_config()
    N = synth.num_anchors(cfgp, H, W)
    loc, cls = synth.head_tensors(B, N, C, mu=-5.0, seed=41)
    fmaps = synth.fpn_maps(B, H, W, Cf, seed=42)
    kw = dict(min_confidence=0.05, nms_iou_threshold=0.4, post_iou_threshold=0.65,
              nms_max_output_size=30, max_k=2, base_size=36)
    probs = {}

    def mask_head(roi_fmaps, roi_boxes):
        probs["m"] = synth.mask_probs(B, roi_boxes.shape[1], C, seed=43)
        return probs["m"]

    want = mo.full_path(loc, cls, fmaps, mask_head, cfgp, (H, W), (PH, PW), **kw)
    seg = synth.semantic_map(B, PH, PW, seed=44)
    seg[0, 100:104, 20:200, 2] = 1
    want_sum = so.summary_output(want["det_i"], seg, want["pasted"])
    for paste in (False, True):
        pipe = ml.PostProcessPipeline(cfgp, (H, W), (PH, PW), C, Cf, B, ml.DetectionConfig(**kw))
        rois = pipe.detect_and_align(dev(loc), dev(cls), [dev(f) for f in fmaps])
        pipe.trim_and_summarize(rois, dev(probs["m"]), dev(seg), paste=paste)
        got = pipe.summary_view().cpu().numpy()
        check_summary(got, want_sum)
        if paste:
            det_i, pasted = pipe.result_views()
            assert np.array_equal(pasted.cpu().numpy(), want["binary"])
            assert np.array_equal(det_i.cpu().numpy(), want["det_i"])


def test_serving_consumers_at_1080p():
    """Frame-resolution check of both consumers (1080 x 1920, the streaming configuration): summary from
    tiles and from materialised masks, overlays from tiles, against the oracles."""
    import masklab_b200 as ml
    from oracle import draw_oracle as do
    B, M, C, PH, PW = 1, 7, 5, 1080, 1920
    seg = synth.semantic_map(B, PH, PW, seed=77)
    seg[0, 600:640, 100:1500, 2] = 1
    det = synth.int_detections(B, M, C, PH, PW, seed=78, pad_tail=1)
    det[0, 0] = [PW // 2, PH // 2, PW + 100, PH + 100, 1, 99]     # frame-sized box: 15 column chunks
    det[0, 1] = [1500, 900, 700, 300, 2, 88]
    rng = np.random.default_rng(79)
    ins = (rng.random((B, M, 28, 28)) > 0.4).astype(np.int32)
    masks = mo.crop_and_pad_mask((PH, PW), det, ins)
    want = so.summary_output(det, seg, masks)
    layer = ml.SummaryOutput()
    check_summary(layer.from_tiles([dev(det), dev(seg), dev(ins)]).cpu().numpy(), want)
    check_summary(layer([dev(det), dev(seg), dev(masks)]).cpu().numpy(), want)
    check_summary(layer([dev(det), dev(seg), dev((masks > 0.5).astype(np.uint8))]).cpu().numpy(),
                  so.summary_output(det, seg, (masks > 0.5).astype(np.float32)))
    img = np.random.default_rng(80).integers(0, 256, (B, PH, PW, 3)).astype(np.uint8)
    colors = [[192, 32, 128], [160, 96, 0], [96, 0, 128], [32, 96, 192], [96, 32, 128]]
    sem_colors = [[64, 0, 128], [128, 96, 0], [128, 192, 0]]
    want_vis = do.draw_segmentation(do.draw_instance(do.draw_boxes(img, det), det, masks, colors, 0.3), seg,
                                    sem_colors, 0.3)
    boxed = ml.DrawBoxes()([dev(img), dev(det)])
    got_vis = ml.DrawInstance(colors, 0.3).from_tiles([boxed, dev(det), dev(ins)], seg_outs=dev(seg),
                                                      semantic_colors=sem_colors, semantic_alpha=0.3)
    assert np.array_equal(got_vis.cpu().numpy(), want_vis)


def test_fractional_semantic_maps_are_rejected_integral_floats_accepted():
    """the kernels read the semantic map as int32; a fractional float map (0.7 -> 0) would silently
    differ from the reference's `> 0.5` / `> 0` / `!= 0` tests, so it is rejected; an integral float map is fine."""
    import masklab_b200 as ml
    from masklab_b200 import runtime as rt
    B, PH, PW, M = 1, 64, 96, 3
    seg = synth.semantic_map(B, PH, PW, seed=3)
    det = synth.int_detections(B, M, 5, PH, PW, seed=4)
    masks = (np.random.default_rng(5).random((B, M, PH, PW)) > 0.7).astype(np.float32)
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    want = ml.SummaryOutput()([d(det), d(seg), d(masks)])
    same = ml.SummaryOutput()([d(det), d(seg.astype(np.float32)), d(masks)])
    assert torch.equal(want, same)
    soft = seg.astype(np.float32) * 0.7
    with pytest.raises(rt.InvalidArgumentError):
        ml.SummaryOutput()([d(det), d(soft), d(masks)])
    with pytest.raises(rt.InvalidArgumentError):
        ml.CrackToInstance()(d(soft[..., 2]))
