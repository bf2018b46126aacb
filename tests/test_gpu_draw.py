"""GPU parity of the overlay layers against oracle/draw_oracle.py through the
C ABI (mlp_draw_segmentation, mlp_draw_instance, mlp_draw_tiles): uint8 images, bit-exact."""
import numpy as np
import pytest
import torch

import synth
from oracle import draw_oracle as do
from oracle import masklab_oracle as mo

pytestmark = pytest.mark.gpu

INST_COLORS = [[192, 32, 128], [160, 96, 0], [96, 0, 128], [32, 96, 192], [96, 32, 128]]   # config.py:32-36
SEM_COLORS = [[64, 0, 128], [128, 96, 0], [128, 192, 0]]                                    # config.py:39-41


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def frames(B, PH, PW, seed):
    return np.random.default_rng(seed).integers(0, 256, (B, PH, PW, 3)).astype(np.uint8)


def scene(B, M, PH, PW, seed, C=5):
    det = synth.int_detections(B, M, C, PH, PW, seed=seed, pad_tail=1)
    det[0, 0] = [PW // 2, PH // 2, 2 * PW, 2 * PH, 1, 99]        # frame-sized box
    det[0, 1] = det[0, 2]                                         # two instances of one class on one box
    det[0, 1, 5] = 95
    if M > 4:
        det[1, 3, 4] = 7                                          # class outside the colour table
    rng = np.random.default_rng(seed + 1)
    ins = (rng.random((B, M, 28, 28)) > 0.45).astype(np.int32)
    return det, ins, mo.crop_and_pad_mask((PH, PW), det, ins)


@pytest.mark.parametrize("PH,PW,img_dtype,seg_dtype", [(40, 64, "u8", "i32"), (33, 50, "f32", "i32"),
                                                       (24, 130, "u8", "f32")])
def test_draw_segmentation(PH, PW, img_dtype, seg_dtype):
    import masklab_b200 as ml
    B = 2
    img = frames(B, PH, PW, 1)
    seg = synth.semantic_map(B, PH, PW, seed=2)
    if img_dtype == "f32":
        img = img.astype(np.float32) + np.float32(0.25)
    if seg_dtype == "f32":
        seg = (seg * np.random.default_rng(3).random(seg.shape)).astype(np.float32)   # soft maps
    want = do.draw_segmentation(img, seg, SEM_COLORS, 0.3)
    got = ml.DrawSegmentation(SEM_COLORS, 0.3)([dev(img), dev(seg)])
    assert got.dtype == torch.uint8 and np.array_equal(got.cpu().numpy(), want)


@pytest.mark.parametrize("PH,PW,M,mask_dtype", [(48, 80, 5, "f32"), (37, 53, 4, "f32"), (48, 80, 5, "u8"),
                                                (64, 300, 7, "f32")])
def test_draw_instance_matches_oracle(PH, PW, M, mask_dtype):
    import masklab_b200 as ml
    B = 2
    img = frames(B, PH, PW, 5)
    det, ins, masks = scene(B, M, PH, PW, seed=PW)
    if mask_dtype == "u8":
        masks = (masks > 0.5).astype(np.uint8)
    want = do.draw_instance(img, det, masks, INST_COLORS, 0.3)
    got = ml.DrawInstance(INST_COLORS, 0.3)([dev(img), dev(det), dev(masks)]).cpu().numpy()
    assert np.array_equal(got, want)
    assert (want != img).any()


@pytest.mark.parametrize("PH,PW,M", [(48, 80, 5), (37, 53, 4), (64, 300, 7), (200, 333, 12)])
def test_draw_from_tiles_equals_draw_of_pasted_masks(PH, PW, M):
    import masklab_b200 as ml
    B = 2
    img = frames(B, PH, PW, 7)
    det, ins, masks = scene(B, M, PH, PW, seed=PW + 3)
    layer = ml.DrawInstance(INST_COLORS, 0.3)
    want = do.draw_instance(img, det, masks, INST_COLORS, 0.3)
    got = layer.from_tiles([dev(img), dev(det), dev(ins)]).cpu().numpy()
    assert np.array_equal(got, want)
    # with the semantic overlay of serving.py:38-40 in the same pass
    seg = synth.semantic_map(B, PH, PW, seed=9)
    want2 = do.draw_segmentation(want, seg, SEM_COLORS, 0.3)
    got2 = layer.from_tiles([dev(img), dev(det), dev(ins)], seg_outs=dev(seg), semantic_colors=SEM_COLORS,
                            semantic_alpha=0.3).cpu().numpy()
    assert np.array_equal(got2, want2)
    # ... and with DrawBoxes in front (serving.py:34), uint8 and float32 frames
    want3 = do.draw_segmentation(do.draw_instance(do.draw_boxes(img, det), det, masks, INST_COLORS, 0.3), seg, SEM_COLORS, 0.3)
    got3 = layer.from_tiles([dev(img), dev(det), dev(ins)], seg_outs=dev(seg), semantic_colors=SEM_COLORS,
                            semantic_alpha=0.3, boxes=True).cpu().numpy()
    assert np.array_equal(got3, want3)
    imgf = (img.astype(np.float32) * 1.1 - 9.5)                        # outside [0, 255] in places, fractional
    want4 = do.draw_instance(do.draw_boxes(imgf, det), det, masks, INST_COLORS, 0.3)
    got4 = layer.from_tiles([dev(imgf), dev(det), dev(ins)], boxes=True).cpu().numpy()
    assert np.array_equal(got4, want4)


@pytest.mark.parametrize("M,PH,PW", [(90, 70, 131), (40, 130, 200)])
def test_draw_from_tiles_crowded_blocks(M, PH, PW):
    """Many boxes over the same pixels: with 90 instances more than 64 boxes touch one 64x64 block (the per-block
    geometry cache overflows and the kernel walks the whole candidate list), with 40 the grouped fast path runs with a
    full list; overlapping instances of one class add up before the > 0.5 test.  Frame sizes that are no multiple of
    the block or of 4 pixels."""
    import masklab_b200 as ml
    B, C = 2, 5
    rng = np.random.default_rng(M + PW)
    det = np.zeros((B, M, 6), np.int32)
    det[..., 0] = rng.integers(0, PW, (B, M))
    det[..., 1] = rng.integers(0, PH, (B, M))
    det[..., 2] = rng.integers(PW // 3, 2 * PW, (B, M))             # wide boxes: every block sees most of them
    det[..., 3] = rng.integers(PH // 3, 2 * PH, (B, M))
    det[..., 4] = rng.integers(0, C, (B, M))
    det[..., 5] = rng.integers(60, 100, (B, M))
    det[1, M - 3:] = -1                                               # padding rows
    ins = (rng.random((B, M, 28, 28)) > 0.8).astype(np.int32)         # sparse tiles: sums of fractions matter
    masks = mo.crop_and_pad_mask((PH, PW), det, ins)
    img = frames(B, PH, PW, 11)
    seg = synth.semantic_map(B, PH, PW, seed=12)
    layer = ml.DrawInstance(INST_COLORS, 0.3)
    want = do.draw_instance(img, det, masks, INST_COLORS, 0.3)
    got = layer.from_tiles([dev(img), dev(det), dev(ins)]).cpu().numpy()
    assert np.array_equal(got, want)
    want2 = do.draw_segmentation(do.draw_instance(do.draw_boxes(img, det), det, masks, INST_COLORS, 0.3), seg, SEM_COLORS, 0.3)
    got2 = layer.from_tiles([dev(img), dev(det), dev(ins)], seg_outs=dev(seg), semantic_colors=SEM_COLORS,
                            semantic_alpha=0.3, boxes=True).cpu().numpy()
    assert np.array_equal(got2, want2)


@pytest.mark.parametrize("kind", ["int_nonbinary", "float_fraction", "float_onehot"])
def test_draw_from_tiles_semantic_map_values(kind):
    """The overlay kernel looks DrawSegmentation's colour term up in a table when an int32 map holds only 0 / 1; every
    other map - int32 values beyond 1 (also negative), float32 maps, fractional or not - takes the arithmetic path.
    One frame mixes both kinds of warps (binary rows above, other values below)."""
    import masklab_b200 as ml
    B, M, PH, PW = 2, 6, 96, 128
    img = frames(B, PH, PW, 21)
    det, ins, masks = scene(B, M, PH, PW, seed=22)
    seg = synth.semantic_map(B, PH, PW, seed=23)
    rng = np.random.default_rng(24)
    if kind == "int_nonbinary":
        seg = seg.astype(np.int32)
        seg[:, PH // 2:] = rng.integers(-2, 4, seg[:, PH // 2:].shape)
        seg[0, 3, 5, 1] = 7
    elif kind == "float_fraction":
        seg = seg.astype(np.float32)
        seg[:, PH // 2:] = rng.random(seg[:, PH // 2:].shape, dtype=np.float32)
    else:
        seg = seg.astype(np.float32)
    layer = ml.DrawInstance(INST_COLORS, 0.3)
    want = do.draw_segmentation(do.draw_instance(do.draw_boxes(img, det), det, masks, INST_COLORS, 0.3), seg, SEM_COLORS, 0.4)
    got = layer.from_tiles([dev(img), dev(det), dev(ins)], seg_outs=dev(seg), semantic_colors=SEM_COLORS,
                           semantic_alpha=0.4, boxes=True).cpu().numpy()
    assert np.array_equal(got, want)


def test_pipeline_draw():
    import masklab_b200 as ml
    B, H, W, C, Cf = 2, 128, 256, 3, 16
    PH, PW = 256, 512
    cfgp = synth.prior_config()
    N = synth.num_anchors(cfgp, H, W)
    loc, cls = synth.head_tensors(B, N, C, mu=-5.0, seed=51)
    fmaps = synth.fpn_maps(B, H, W, Cf, seed=52)
    kw = dict(min_confidence=0.05, nms_iou_threshold=0.4, post_iou_threshold=0.65,
              nms_max_output_size=30, max_k=2, base_size=36)
    probs = {}

    def mask_head(roi_fmaps, roi_boxes):
        probs["m"] = synth.mask_probs(B, roi_boxes.shape[1], C, seed=53)
        return probs["m"]

    want = mo.full_path(loc, cls, fmaps, mask_head, cfgp, (H, W), (PH, PW), **kw)
    img = frames(B, PH, PW, 54)
    seg = synth.semantic_map(B, PH, PW, seed=55)
    vis_i = do.draw_instance(img, want["det_i"], want["pasted"], INST_COLORS[:C], 0.3)
    vis = do.draw_segmentation(vis_i, seg, SEM_COLORS, 0.3)
    pipe = ml.PostProcessPipeline(cfgp, (H, W), (PH, PW), C, Cf, B, ml.DetectionConfig(**kw))
    rois = pipe.detect_and_align(dev(loc), dev(cls), [dev(f) for f in fmaps])
    pipe.trim_and_summarize(rois, dev(probs["m"]), dev(seg))
    got_i = pipe.draw(rois, dev(probs["m"]), dev(img), INST_COLORS[:C], 0.3).cpu().numpy().copy()
    assert np.array_equal(got_i, vis_i)
    got = pipe.draw(rois, dev(probs["m"]), dev(img), INST_COLORS[:C], 0.3, seg_outs=dev(seg),
                    semantic_colors=SEM_COLORS, semantic_alpha=0.3).cpu().numpy()
    assert np.array_equal(got, vis)
    # the whole visualisation branch of serving.py:34-40: boxes, instances, semantic
    vis_b = do.draw_boxes(img, want["det_i"])
    vis_all = do.draw_segmentation(do.draw_instance(vis_b, want["det_i"], want["pasted"], INST_COLORS[:C], 0.3),
                                   seg, SEM_COLORS, 0.3)
    got_all = pipe.draw(rois, dev(probs["m"]), dev(img), INST_COLORS[:C], 0.3, seg_outs=dev(seg),
                        semantic_colors=SEM_COLORS, semantic_alpha=0.3, boxes=True).cpu().numpy()
    assert np.array_equal(got_all, vis_all)
    # ... and its JPEG encode (serving.py:41), every frame of the batch
    from oracle import jpeg_oracle as jo
    files, lengths = pipe.encode()
    files, lengths = files.cpu().numpy(), lengths.cpu().numpy()
    for b in range(B):
        assert files[b, :lengths[b]].tobytes() == jo.encode_jpeg(vis_all[b])
    # the whole serving tail as one CUDA graph: same summary, overlay and files, also after other inputs passed through
    d_loc, d_cls, d_fm, d_m, d_seg, d_img = dev(loc), dev(cls), [dev(f) for f in fmaps], dev(probs["m"]), dev(seg), dev(img)
    ref_summary = pipe.summary.clone()
    graph, _ = pipe.capture_serving(d_loc, d_cls, d_fm, d_m, d_seg, d_img, INST_COLORS[:C], 0.3,
                                    semantic_colors=SEM_COLORS, semantic_alpha=0.3, boxes=True)
    keep = d_img.clone()
    d_img.copy_(255 - keep)
    graph.replay()
    torch.cuda.synchronize()
    assert not np.array_equal(pipe.vis.cpu().numpy(), vis_all)
    d_img.copy_(keep)
    graph.replay()
    torch.cuda.synchronize()
    assert np.array_equal(pipe.vis.cpu().numpy(), vis_all)
    f2, l2 = pipe.jpeg_files.cpu().numpy(), pipe.jpeg_len.cpu().numpy()
    for b in range(B):
        assert f2[b, :l2[b]].tobytes() == files[b, :lengths[b]].tobytes()
    Mo = int(pipe.summary_m.item())
    assert torch.equal(pipe.summary[:B * Mo * 11], ref_summary[:B * Mo * 11])
    assert ml.DrawInstance(INST_COLORS).get_config()["alpha"] == 0.3


def test_draw_rejects_cpu_tensors():
    import masklab_b200 as ml
    with pytest.raises(ml.InvalidArgumentError):
        ml.DrawSegmentation(SEM_COLORS)([torch.zeros((1, 4, 4, 3), dtype=torch.uint8),
                                         torch.zeros((1, 4, 4, 3), dtype=torch.int32)])


@pytest.mark.parametrize("PH,PW,img_dtype", [(48, 80, "u8"), (37, 53, "f32"), (200, 333, "u8")])
def test_draw_boxes(PH, PW, img_dtype):
    import masklab_b200 as ml
    B, M = 2, 9
    img = frames(B, PH, PW, 11)
    if img_dtype == "f32":
        img = img.astype(np.float32) * np.float32(1.1) - np.float32(10.0)       # needs the clip
    det = synth.int_detections(B, M, 5, PH, PW, seed=PH, pad_tail=2)
    det[0, 0] = [1, 1, 3, 3, 0, 90]                     # corner in (-1, 0) px: truncation draws the line
    det[0, 1] = [1, 1, 6, 6, 0, 90]                     # corner beyond -1 px: line outside
    det[0, 2] = [PW, PH, 40, 40, 0, 90]                 # sticks out bottom/right
    det[1, 0] = [PW // 2, PH // 2, 0, 0, 0, 90]         # zero-size box: a single pixel
    det[1, 1] = [5 * PW, 5 * PH, 4, 4, 0, 90]           # entirely outside
    want = do.draw_boxes(img, det)
    got = ml.DrawBoxes()([dev(img), dev(det)]).cpu().numpy()
    assert np.array_equal(got, want)
    assert (want == 255).sum() > (np.clip(img, 0, 255).astype(np.uint8) == 255).sum()
