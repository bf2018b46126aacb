"""GPU parity tests: every drop-in layer (CUDA through the C ABI) against the CPU oracle on
the same seeded inputs.  Bars (BASELINE.json north_star): kept indices, class ids and
binary masks bit-exact; decoded boxes <= 1e-5 relative; RoIAlign features <= 1e-4 absolute.
"""
import numpy as np
import pytest
import torch

import synth
from oracle import masklab_oracle as mo

pytestmark = pytest.mark.gpu

F32 = np.float32


@pytest.fixture(scope="module")
def ml():
    import masklab_b200
    masklab_b200.Context.get(0)
    return masklab_b200


def dev(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def host(t):
    return t.detach().cpu().numpy()


# ------------------------------------------------------------------ a1/a2 -----
@pytest.mark.parametrize("hw,padding", [((64, 128), "same"), ((540, 960), "same"), ((100, 75), "same"),
                                        ((100, 75), "valid"), ((512, 1024), "same")])
def test_prior_layer_exact(ml, hw, padding):
    cfg = synth.prior_config()
    table = mo.prior_table(**cfg)
    want = mo.prior_layer(table, hw[0], hw[1], padding)
    layer = ml.PriorLayer(cfg, padding=padding)
    images = torch.zeros((2, hw[0], hw[1], 3), dtype=torch.uint8, device="cuda")
    got = host(layer(images))
    assert got.dtype == np.int32 and got.shape == (2,) + want.shape
    assert np.array_equal(got[0], want) and np.array_equal(got[1], want)
    assert layer.get_config()["prior"] == cfg and layer.get_config()["padding"] == padding
    assert layer.trainable is False


def test_prior_layer_road_project_config(ml):
    cfg = synth.prior_config(strides=(8, 16, 32, 64), ratios=(1 / 2, 1, 2, 5, 8))
    want = mo.prior_layer(mo.prior_table(**cfg), 540, 960)
    got = host(ml.PriorLayer(ml.PriorBoxes(**cfg))(torch.zeros((1, 540, 960, 3), device="cuda")))
    assert np.array_equal(got[0], want)


# --------------------------------------------------------------------- a3 -----
def test_restore_boxes(ml):
    cfg = synth.prior_config()
    H, W, B = 128, 256, 3
    pr = mo.prior_layer(mo.prior_table(**cfg), H, W)
    N = pr.shape[0]
    loc, _ = synth.head_tensors(B, N, 1, seed=5)
    loc[0, :8] = [[0, 0, 0, 0], [2, -2, 2, -2], [-2, 2, -2, 2], [1e-8, -1e-8, 1e-8, -1e-8],
                  [0.5, 0.25, 0.125, 1.0], [-0.0, 0.0, -0.0, 0.0], [1.5, 1.5, 1.5, 1.5],
                  [-1.25, 0.75, -0.3, 0.3]]
    prb = np.broadcast_to(pr[None], (B,) + pr.shape)
    want = mo.restore_boxes(loc, prb)
    got_i = host(ml.RestoreBoxes()([dev(loc), dev(prb)]))
    got_f = host(ml.RestoreBoxes()([dev(loc), dev(prb.astype(F32))]))
    for got in (got_i, got_f):
        rel = np.abs(got - want) / np.maximum(np.abs(want), 1e-12)
        assert rel.max() <= 1e-5                                    # north_star tolerance
        assert np.mean(got != want) <= 1e-6                         # and in fact (almost surely) exact


def test_restore_boxes_from_prior_matches_layer(ml):
    import ctypes
    from masklab_b200 import runtime as rt
    cfg = synth.prior_config()
    H, W, B = 96, 160, 2
    prior = ml.PriorBoxes(**cfg)
    pr = mo.prior_layer(mo.prior_table(**cfg), H, W)
    loc, _ = synth.head_tensors(B, pr.shape[0], 1, seed=6)
    want = mo.restore_boxes(loc, np.broadcast_to(pr[None], (B,) + pr.shape))
    ctx = ml.Context.get(0)
    loc_d = dev(loc)
    out = torch.empty_like(loc_d)
    pc = prior.to_c("same")
    rt.check(ctx.lib.mlp_restore_boxes_from_prior(ctx.handle, ctypes.byref(pc), ctx.view(loc_d), B, H, W,
                                                  ctx.view(out), ctx.stream()))
    assert np.array_equal(host(out), want)


# --------------------------------------------------------------------- a4 -----
def test_normalize_boxes_exact(ml):
    rng = np.random.default_rng(0)
    boxes = (rng.random((4, 37, 4), dtype=F32) * 500).astype(F32)
    for shape in (None, (540.0, 960.0), (512, 1024)):
        want = mo.normalize_boxes(boxes) if shape is None else mo.normalize_boxes(boxes, shape)
        kw = {} if shape is None else {"shape": torch.tensor(shape)}
        got = host(ml.NormalizeBoxes()(dev(boxes), **kw))
        assert np.array_equal(got, want)
    rows7 = (rng.random((5, 7), dtype=F32) * 100).astype(F32)
    assert np.array_equal(host(ml.NormalizeBoxes()(dev(rows7), shape=(64, 64))),
                          mo.normalize_boxes(rows7, (64, 64)))


# ------------------------------------------------------------------ a5-a8 -----
def _proposal_case(B, H, W, C, mu, seed, ratios=synth.DEFAULT_RATIOS):
    cfg = synth.prior_config(ratios=ratios)
    pr = mo.prior_layer(mo.prior_table(**cfg), H, W)
    N = pr.shape[0]
    loc, cls = synth.head_tensors(B, N, C, mu=mu, seed=seed)
    boxes = mo.restore_boxes(loc, np.broadcast_to(pr[None], (B,) + pr.shape))
    return cfg, loc, cls, boxes


def _check_proposal(ml, cls, boxes, **kw):
    want, dbg = mo.detection_proposal(cls, boxes, return_debug=True, **kw)
    layer = ml.DetectionProposal(**kw)
    got = host(layer([dev(cls), dev(boxes), None]))
    assert got.shape == want.shape, (got.shape, want.shape)
    assert np.array_equal(got, want)
    # kept (anchor, class) indices, bit exact, in output order
    keep = host(layer.last_keep)
    counts = host(layer.last_counts)
    final = dbg["keep"]
    for b in range(cls.shape[0]):
        rows = final[final[:, 0] == b]
        assert counts[b] == rows.shape[0]
        assert np.array_equal(keep[b, :counts[b]], rows[:, 1:].astype(np.int32))
        assert np.all(keep[b, counts[b]:] == -1)
    return want, dbg


@pytest.mark.parametrize("mu,max_out", [(-5.8, 100), (-5.0, 100), (-4.0, 50), (-5.0, 1000)])
def test_detection_proposal_matches_oracle(ml, mu, max_out):
    _, _, cls, boxes = _proposal_case(3, 128, 256, 5, mu, seed=11)
    _check_proposal(ml, cls, boxes, min_confidence=0.05, nms_iou_threshold=0.4,
                    post_iou_threshold=0.65, nms_max_output_size=max_out)


def test_detection_proposal_dense_overlaps(ml):
    """Low thresholds -> heavy suppression chains, several NMS chunks, cap not binding."""
    _, _, cls, boxes = _proposal_case(2, 96, 96, 3, -2.0, seed=12)
    want, dbg = _check_proposal(ml, cls, boxes, min_confidence=0.05, nms_iou_threshold=0.1,
                                post_iou_threshold=0.2, nms_max_output_size=1000)
    assert dbg["candidates"].shape[0] > 3000


def test_detection_proposal_many_candidates_radix_path(ml):
    """More candidates per (image,class) than the shared-memory sort buffer holds."""
    _, _, cls, boxes = _proposal_case(1, 160, 160, 2, -1.0, seed=13)
    want, dbg = _check_proposal(ml, cls, boxes, min_confidence=0.05, nms_iou_threshold=0.3,
                                post_iou_threshold=0.5, nms_max_output_size=600)
    per_group = np.bincount(dbg["candidates"][:, 2])
    assert per_group.max() > 4096


def test_detection_proposal_score_ties_break_on_anchor_index(ml):
    _, _, cls, boxes = _proposal_case(2, 64, 64, 3, -5.0, seed=14)
    cls = np.round(cls * 20).astype(F32) / 20                      # heavy score quantisation
    cls[:, ::7, :] = F32(0.25)
    want, dbg = _check_proposal(ml, cls, boxes, min_confidence=0.05, nms_iou_threshold=0.4,
                                post_iou_threshold=0.65, nms_max_output_size=200)
    assert dbg["candidates"].shape[0] > 100


def test_detection_proposal_empty_and_ragged(ml):
    _, _, cls, boxes = _proposal_case(4, 64, 128, 4, -5.0, seed=15)
    cls[1] = 0.0                                                   # image with no candidate
    cls[3, :, 1:] = 0.0                                            # image with a single class
    _check_proposal(ml, cls, boxes, nms_max_output_size=30)
    none = np.zeros_like(cls)                                      # nothing anywhere -> [B,1,6] of -1
    want = mo.detection_proposal(none, boxes)
    got = host(ml.DetectionProposal()([dev(none), dev(boxes), None]))
    assert want.shape == (4, 1, 6) and np.array_equal(got, want)


def test_detection_proposal_batch_limit(ml):
    cls = torch.zeros((33, 16, 2), device="cuda")
    boxes = torch.ones((33, 16, 4), device="cuda")
    with pytest.raises(ValueError):                                # misc.py:275, 32 partitions
        ml.DetectionProposal()([cls, boxes, None])
    out = ml.DetectionProposal(max_batch_size=None)([cls, boxes, None])
    assert tuple(out.shape) == (33, 1, 6)


def test_detection_proposal_config_roundtrip(ml):
    layer = ml.DetectionProposal(min_confidence=0.3, nms_iou_threshold=0.5, post_iou_threshold=0.7,
                                 nms_max_output_size=77, max_batch_size=8, name="dp")
    clone = ml.DetectionProposal.from_config(layer.get_config())
    assert clone.get_config() == layer.get_config()
    assert ml.get_custom_objects()["DetectionProposal"] is ml.DetectionProposal


def test_cpu_tensor_is_rejected(ml):
    with pytest.raises(ValueError):
        ml.DetectionProposal()([torch.zeros((1, 4, 2)), torch.zeros((1, 4, 4)), None])


# --------------------------------------------------------------------- a9 -----
def test_mask_distribute_exact(ml):
    det = synth.detections(3, 200, 5, 512, 1024, seed=21, lo=4.0, hi=900.0, pad_tail=17)
    det[0, 0, 2:4] = [64, 64]            # exactly on a level boundary
    det[0, 1, 2:4] = [128, 128]
    det[0, 2, 2:4] = [256, 256]
    det[0, 3, 2:4] = [36, 36]
    for max_k, base in ((2, 64), (2, 36), (3, 32)):
        want, margin = mo.mask_distribute(det, max_k, base, return_margin=True)
        got = host(ml.MaskDistribute(max_k=max_k, base_size=base)(dev(det)))
        assert np.array_equal(got, want)
    assert ml.MaskDistribute(max_k=1, base_size=40).get_config()["base_size"] == 40


# -------------------------------------------------------------------- a10 -----
@pytest.mark.parametrize("Cf,crop", [(128, (14, 14)), (32, (7, 7)), (20, (14, 14)), (6, (3, 5))])
def test_pyramid_roi_align(ml, Cf, crop):
    B, H, W, M = 3, 128, 256, 40
    det = synth.detections(B, M, 4, H, W, seed=31, lo=8.0, hi=300.0, pad_tail=5)
    det[1, 7:] = -1                                                # ragged counts
    det[2, 0, :4] = [10.0, 10.0, 80.0, 80.0]                       # sticks out of the frame
    det[2, 1, :4] = [300.0, 200.0, 20.0, 20.0]                     # fully outside
    dist = mo.mask_distribute(det, 2, 36)
    fmaps = synth.fpn_maps(B, H, W, Cf, seed=32)
    want_f, want_b = mo.pyramid_roi_align(fmaps, dist, (H, W), crop)
    images = torch.zeros((B, H, W, 3), device="cuda")
    got_f, got_b = ml.PyramidRoiAlign(crop_size=crop)([[dev(f) for f in fmaps], dev(dist), images])
    assert np.array_equal(host(got_b), want_b)
    assert len(got_f) == 3
    for g, w in zip(got_f, want_f):
        g = host(g)
        assert g.shape == w.shape
        assert np.abs(g - w).max() <= 1e-4                         # north_star tolerance
        assert np.array_equal(g, w)                                # and exact (no FMA, same order)


def test_pyramid_roi_align_level_without_boxes(ml):
    B, H, W = 2, 64, 64
    det = synth.detections(B, 6, 2, H, W, seed=33, lo=8.0, hi=20.0)   # everything lands on level 0
    dist = mo.mask_distribute(det, 2, 64)
    fmaps = synth.fpn_maps(B, H, W, 8, seed=34)
    want_f, want_b = mo.pyramid_roi_align(fmaps, dist, (H, W))
    got_f, got_b = ml.PyramidRoiAlign()([[dev(f) for f in fmaps], dev(dist),
                                         torch.zeros((B, H, W, 3), device="cuda")])
    assert [tuple(t.shape) for t in got_f] == [w.shape for w in want_f]
    assert want_f[1].shape[1] == 1 and np.all(want_f[1] == -1)     # one all -1 slot (MoldBatch)
    for g, w in zip(got_f, want_f):
        assert np.array_equal(host(g), w)
    assert np.array_equal(host(got_b), want_b)


# -------------------------------------------------------------------- a11 -----
def test_trim_instances_exact(ml):
    B, H, W, C = 3, 128, 256, 5
    det = synth.detections(B, 30, C, H, W, seed=41, lo=8.0, hi=300.0, pad_tail=3)
    det[2, 11:] = -1
    dist = mo.mask_distribute(det, 2, 36)
    fmaps = synth.fpn_maps(B, H, W, 4, seed=42)
    _, roi_boxes = mo.pyramid_roi_align(fmaps, dist, (H, W))
    roi_masks = synth.mask_probs(B, roi_boxes.shape[1], C, seed=43)
    want_b, want_m = mo.trim_instances(roi_boxes, roi_masks)
    got_b, got_m = ml.TrimInstances()([dev(roi_boxes), dev(roi_masks)])
    assert np.array_equal(host(got_b), want_b) and np.array_equal(host(got_m), want_m)
    flat_b, flat_m = ml.TrimInstances(mold=False)([dev(roi_boxes), dev(roi_masks)])
    wb, wm = mo.trim_instances(roi_boxes, roi_masks, mold=False)
    assert np.array_equal(host(flat_b), wb) and np.array_equal(host(flat_m), wm)


def test_trim_instances_all_padding(ml):
    rb = -np.ones((2, 3, 6), F32)
    rm = synth.mask_probs(2, 3, 2, seed=44)
    wb, wm = mo.trim_instances(rb, rm)
    gb, gm = ml.TrimInstances()([dev(rb), dev(rm)])
    assert wb.shape == (2, 1, 6) and np.array_equal(host(gb), wb) and np.array_equal(host(gm), wm)


# -------------------------------------------------------------------- a12 -----
def test_upsample_output_exact(ml):
    det = synth.detections(2, 50, 5, 540, 960, seed=51, pad_tail=4)
    masks = synth.mask_probs(2, 50, 1, seed=52)[..., 0]
    masks[0, 0, 0, :4] = [0.5, np.nextafter(F32(0.5), F32(1)), 0.49999997, 1.0]
    sem = torch.zeros((2, 540, 960, 3), device="cuda")
    for dst in ((1080, 1920), (540, 960), (720, 1000)):
        target = torch.zeros((2, dst[0], dst[1], 3), device="cuda")
        want_b, want_m = mo.upsample_output(det, masks, (540, 960), dst)
        got_b, got_m, sem_out = ml.UpSampleOutput(semantic=False)([dev(det), dev(masks), sem], target=target)
        assert got_b.dtype == torch.int32 and got_m.dtype == torch.int32
        assert np.array_equal(host(got_b), want_b) and np.array_equal(host(got_m), want_m)
        assert sem_out is sem


@pytest.mark.parametrize("src,dst", [((27, 48), (54, 96)), ((30, 41), (77, 53)), ((20, 32), (20, 32)), ((9, 7), (1, 1))])
def test_upsample_output_semantic_half(ml, src, dst):
    """misc.py:190-195: resize_bilinear(align_corners=True) to the frame, > 0.5, int32."""
    from oracle import semantic_oracle as so
    rng = np.random.default_rng(src[0])
    sem = rng.random((2, src[0], src[1], 3)).astype(F32)
    det = synth.detections(2, 4, 5, src[0], src[1], seed=1)
    masks = synth.mask_probs(2, 4, 1, seed=2)[..., 0]
    target = torch.zeros((2, dst[0], dst[1], 3), device="cuda")
    _, _, got = ml.UpSampleOutput()([dev(det), dev(masks), dev(sem)], target=target)
    assert got.dtype == torch.int32 and np.array_equal(host(got), so.upsample_semantic(sem, dst))


@pytest.mark.parametrize("hw,target", [((108, 192), (54, 96)), ((100, 100), (54, 96)), ((50, 333), (54, 96))])
def test_downsample_input(ml, hw, target):
    """misc.py:143-154: min ratio, truncated size, bilinear align_corners=True, float32 out."""
    from oracle import semantic_oracle as so
    frames = np.random.default_rng(hw[1]).integers(0, 256, (2, hw[0], hw[1], 3)).astype(np.uint8)
    want = so.downsample_input(frames, target)
    got = ml.DownSampleInput(target)(dev(frames))
    assert got.dtype == torch.float32 and tuple(got.shape) == want.shape
    assert np.array_equal(host(got), want)
    got_f = ml.DownSampleInput(target)(dev(frames.astype(F32)))
    assert np.array_equal(host(got_f), want)
    assert ml.DownSampleInput(target).get_config()["target_size"] == target


# ---------------------------------------------------------------- a13/a14 -----
def _paste_case(B, M, PH, PW, seed, pad_tail=2):
    det = synth.detections(B, M, 5, PH, PW, seed=seed, lo=2.0, hi=float(min(PH, PW)), pad_tail=pad_tail)
    masks = synth.mask_probs(B, M, 1, seed=seed + 1)[..., 0]
    det_i, mask_i = mo.upsample_output(det, masks, (PH, PW), (PH, PW))
    return det_i, mask_i


@pytest.mark.parametrize("PH,PW", [(128, 256), (135, 240), (64, 100), (50, 37), (96, 1024), (40, 1920)])
def test_crop_and_pad_mask_exact(ml, PH, PW):
    B, M = 2, 24
    det_i, mask_i = _paste_case(B, M, PH, PW, seed=61)
    det_i[0, 0, :4] = [PW + 50, 10, 20, 20]          # box fully right of the frame -> zero size
    det_i[0, 1, :4] = [5, 5, 40, 40]                 # clipped at the top-left corner
    det_i[0, 2, :4] = [PW // 2, PH // 2, 1, 1]       # 1x1 box
    det_i[0, 3, :4] = [PW // 2, PH // 2, 2 * PW, 2 * PH]   # covers the whole frame
    det_i[1, 0, 5] = 50                              # conf == threshold passes (>=)
    det_i[1, 1, 5] = 49                              # below threshold -> zeros
    want = mo.crop_and_pad_mask((PH, PW), det_i, mask_i)
    images = torch.zeros((B, PH, PW, 3), dtype=torch.uint8, device="cuda")
    got = host(ml.CropAndPadMask()([images, dev(det_i), dev(mask_i)]))
    assert got.dtype == np.float32 and np.array_equal(got, want)
    got8 = host(ml.CropAndPadMask(output="uint8")([images, dev(det_i), dev(mask_i)]))
    assert got8.dtype == np.uint8 and np.array_equal(got8, mo.binary_masks(want))
    if PW % 8 == 0:
        bits = host(ml.CropAndPadMask(output="bits")([images, dev(det_i), dev(mask_i)]))
        assert bits.shape == (B, M, PH, PW // 8) and np.array_equal(bits, mo.packed_masks(want))
    else:
        with pytest.raises(ValueError):
            ml.CropAndPadMask(output="bits")([images, dev(det_i), dev(mask_i)])
    assert want[1, 1].sum() == 0 and want[0, 0].sum() == 0


def test_crop_and_pad_mask_low_confidence_branch(ml):
    """max(conf) <= 50 -> threshold -100: every row, padding included, is pasted (misc.py:366-370)."""
    PH, PW = 64, 96
    det_i, mask_i = _paste_case(2, 10, PH, PW, seed=63)
    det_i[..., 5] = np.minimum(det_i[..., 5], 40)
    want = mo.crop_and_pad_mask((PH, PW), det_i, mask_i)
    got = host(ml.CropAndPadMask()([(PH, PW), dev(det_i), dev(mask_i)]))
    assert np.array_equal(got, want)


# ------------------------------------------------------------------- mold -----
def test_mold_batch(ml):
    rng = np.random.default_rng(71)
    x = rng.random((57, 3, 5), dtype=F32)
    bi = rng.integers(0, 6, 57)
    bi[bi == 4] = 2                                               # image 4 gets nothing
    want = mo.mold_batch(x, bi, 6)
    got = host(ml.MoldBatch(max_batch_size=64)(dev(x), batch_indices=dev(bi), batch_size=6))
    assert np.array_equal(got, want)
    empty = host(ml.MoldBatch()(dev(np.zeros((0, 6), F32)), batch_indices=dev(np.zeros((0,), np.int64)),
                                batch_size=3))
    assert empty.shape == (3, 1, 6) and np.all(empty == -1)
    with pytest.raises(ValueError):
        ml.MoldBatch(max_batch_size=64)(dev(x), batch_indices=dev(bi), batch_size=33)


# --------------------------------------------------------------- whole path ---
@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("paste", ["uint8", "float32", "bits"])
def test_pipeline_matches_oracle(ml, paste, fused):
    B, H, W, C, Cf = 3, 128, 256, 4, 16
    PH, PW = 256, 512
    cfgp = synth.prior_config()
    N = synth.num_anchors(cfgp, H, W)
    loc, cls = synth.head_tensors(B, N, C, mu=-5.0, seed=81)
    cls[1] = 0                                                    # an image without detections
    fmaps = synth.fpn_maps(B, H, W, Cf, seed=82)
    kw = dict(min_confidence=0.05, nms_iou_threshold=0.4, post_iou_threshold=0.65,
              nms_max_output_size=60, max_k=2, base_size=36)
    probs = {}

    def mask_head(roi_fmaps, roi_boxes):
        probs["m"] = synth.mask_probs(B, roi_boxes.shape[1], C, seed=83)
        return probs["m"]

    want = mo.full_path(loc, cls, fmaps, mask_head, cfgp, (H, W), (PH, PW), **kw)
    cfg = ml.DetectionConfig(paste_output=paste, fused=fused, **kw)
    pipe = ml.PostProcessPipeline(cfgp, (H, W), (PH, PW), C, Cf, B, cfg)
    rois = pipe.detect_and_align(dev(loc), dev(cls), [dev(f) for f in fmaps])
    crops, roi_boxes = pipe.roi_views(rois)
    M = int(rois.m_dev.item())
    assert np.array_equal(host(rois.det[:, :M]), want["proposed"])
    assert np.array_equal(host(roi_boxes), want["roi_boxes"])
    for g, w in zip(crops, want["roi_fmaps"]):
        assert np.array_equal(host(g), w)
    pipe.trim_and_paste(rois, dev(probs["m"]))
    det_i, pasted = pipe.result_views()
    assert np.array_equal(host(det_i), want["det_i"])
    if paste == "uint8":
        assert np.array_equal(host(pasted), want["binary"])
    elif paste == "bits":
        assert np.array_equal(host(pasted), mo.packed_masks(want["pasted"]))
    else:
        assert np.array_equal(host(pasted), want["pasted"])


@pytest.mark.parametrize("fused", [True, False])
def test_pipeline_no_detections_and_high_confidence(ml, fused):
    """Both confidence-threshold branches of the paste through the pipeline: (a) nothing detected
    anywhere (one all -1 slot per image), (b) confident detections (max conf > 50)."""
    B, H, W, C, Cf = 2, 96, 160, 3, 8
    cfgp = synth.prior_config()
    N = synth.num_anchors(cfgp, H, W)
    fmaps = synth.fpn_maps(B, H, W, Cf, seed=92)
    kw = dict(min_confidence=0.05, nms_iou_threshold=0.4, post_iou_threshold=0.65,
              nms_max_output_size=20, max_k=2, base_size=36)
    for case in ("none", "confident"):
        loc, cls = synth.head_tensors(B, N, C, mu=-5.0, seed=91)
        if case == "none":
            cls[:] = 0
        else:
            cls[:, ::97, :] = np.float32(0.93)
        probs = {}

        def mask_head(roi_fmaps, roi_boxes):
            probs["m"] = synth.mask_probs(B, roi_boxes.shape[1], C, seed=93)
            return probs["m"]

        want = mo.full_path(loc, cls, fmaps, mask_head, cfgp, (H, W), (H, W), **kw)
        pipe = ml.PostProcessPipeline(cfgp, (H, W), (H, W), C, Cf, B, ml.DetectionConfig(fused=fused, **kw))
        rois = pipe.detect_and_align(dev(loc), dev(cls), [dev(f) for f in fmaps])
        crops, roi_boxes = pipe.roi_views(rois)
        assert np.array_equal(host(roi_boxes), want["roi_boxes"])
        pipe.trim_and_paste(rois, dev(probs["m"]))
        det_i, pasted = pipe.result_views()
        assert np.array_equal(host(det_i), want["det_i"])
        assert np.array_equal(host(pasted), want["binary"])
        if case == "confident":
            assert want["det_i"][..., 5].max() > 50 and want["binary"].sum() > 0


@pytest.mark.parametrize("shape,k,weight", [((2, 40, 56, 3), 10, 1.0), ((1, 33, 29, 2), 4, 2.0), ((2, 9, 11, 3), 1, 0.5),
                                            ((1, 6, 7, 1), 10, 1.0), ((2, 12, 12, 3), 0, 0.25), ((1, 50, 70, 4), 7, 1.5)])
def test_semantic_smoothing(ml, shape, k, weight):
    """semantic.py:270-284: erosion2d then dilation2d with a flat k x k kernel, SAME, times weight."""
    from oracle import semantic_oracle as so
    rng = np.random.default_rng(shape[1] + k)
    x = rng.random(shape).astype(F32)
    x[x < 0.3] = 0.0                                                 # plateaus and specks
    want = so.semantic_smoothing(x, k, weight)
    got = ml.SemanticSmoothing(kernel_size=k, weight=weight)(dev(x))
    assert got.dtype == torch.float32 and np.array_equal(host(got), want)


@pytest.mark.parametrize("k", [2, 3, 9, 16, 17])
def test_semantic_smoothing_tiles_and_window_sizes(ml, k):
    """Several 32 x 64 tiles with partial ones at the right / bottom edge, window sizes either side of the fused
    kernel's limit (16); the tile-fused opening, the four streaming passes and the oracle agree bit for bit."""
    import os
    from oracle import semantic_oracle as so
    rng = np.random.default_rng(100 + k)
    x = rng.random((2, 100, 203, 2)).astype(F32)
    x[x < 0.35] = 0.0
    want = so.semantic_smoothing(x, k, 1.25)
    layer = ml.SemanticSmoothing(kernel_size=k, weight=1.25)
    got = host(layer(dev(x)))
    assert np.array_equal(got, want)
    os.environ["MLP_SMOOTH_PASSES"] = "1"
    try:
        passes = host(layer(dev(x)))
    finally:
        del os.environ["MLP_SMOOTH_PASSES"]
    assert np.array_equal(passes, want)
