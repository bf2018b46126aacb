"""Round-2 features of the fused pipeline on the GPU, each against the C restatement of the reference
path (oracle/c) or against the unchanged pipeline on the same inputs:

* the background fill of CropAndPadMask on a second stream (mlp_paste_prefill + MLP_PASTE_PREFILLED) at
  the full cfg-2 size, eager and as one captured CUDA graph with a fork / join inside;
* the planar mask-head layout ([B,R,C,mh,mw]) and the bulk-copy tail preparation against the register
  gather it replaced (MLP_TAIL_TMA=0);
* scratch safety: a captured pipeline owns a private, frozen ctx and a larger shape raises MLP_EFROZEN
  instead of freeing memory a graph references;
* determinism of the atomics-based finalisers (box summary, JPEG placement): 100 replays, equal bytes.
"""
import os

import numpy as np
import pytest
import torch

import synth
from oracle import c_oracle as co

pytestmark = pytest.mark.gpu

WL = dict(B=32, H=512, W=1024, C=5, Cf=128, mu=-5.8)
KW = dict(min_confidence=0.05, nms_iou_threshold=0.4, post_iou_threshold=0.6, nms_max_output_size=100,
          max_k=2, base_size=36)


def _d(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.fixture(scope="module")
def cfg2():
    import masklab_b200 as ml
    B, H, W, C, Cf = WL["B"], WL["H"], WL["W"], WL["C"], WL["Cf"]
    cfgp = synth.prior_config()
    N = synth.num_anchors(cfgp, H, W)
    loc, cls = synth.head_tensors(B, N, C, mu=WL["mu"], seed=4321)
    cls[3] = 0                                                  # a frame without detections
    fmaps = synth.fpn_maps(B, H, W, Cf, seed=4322)
    base = ml.PostProcessPipeline(cfgp, (H, W), (H, W), C, Cf, B, ml.DetectionConfig(**KW))
    ins = (_d(loc), _d(cls), [_d(f) for f in fmaps])
    rois = base.detect_and_align(*ins)
    _, R = rois.shapes()
    probs = synth.mask_probs(B, R, C, seed=4323)
    d_probs = _d(probs)
    base.trim_and_paste(rois, d_probs)
    det_i, pasted = base.result_views()
    want = co.full_path(loc, cls, fmaps, lambda f, b: probs, cfgp, (H, W), (H, W), binary=True, **KW)
    assert np.array_equal(det_i.cpu().numpy(), want["det_i"])
    assert np.array_equal(pasted.cpu().numpy(), want["binary"])
    return dict(ml=ml, cfgp=cfgp, ins=ins, probs=probs, d_probs=d_probs, want=want, M=int(base.trim_m.item()),
                det_i=det_i.clone(), pasted=pasted.clone())


@pytest.mark.parametrize("mode", ["uint8", "bits", "float32"])
def test_prefilled_paste_equals_plain_paste_full_size(cfg2, mode):
    ml = cfg2["ml"]
    B, H, W, C, Cf = WL["B"], WL["H"], WL["W"], WL["C"], WL["Cf"]
    outs = []
    for prefill in (False, True):
        cfg = ml.DetectionConfig(paste_output=mode, prefill=prefill, **KW)
        pipe = ml.PostProcessPipeline(cfg2["cfgp"], (H, W), (H, W), C, Cf, B, cfg)
        pipe.pasted.fill_(3)                                    # stale bytes: the fill has to clear them
        for _ in range(2):                                      # second pass: fill after a previous batch's boxes
            rois = pipe.detect_and_align(*cfg2["ins"])
            pipe.trim_and_paste(rois, cfg2["d_probs"])
        det_i, pasted = pipe.result_views()
        outs.append((det_i.clone(), pasted.clone()))
        del pipe
    assert torch.equal(outs[0][0], outs[1][0])
    assert torch.equal(outs[0][1], outs[1][1])
    if mode == "uint8":
        assert torch.equal(outs[1][1], cfg2["pasted"])


def test_prefilled_graph_replay_full_size(cfg2):
    """One CUDA graph with the fork (fill beside RoIAlign + tail prep) and the join inside."""
    ml = cfg2["ml"]
    B, H, W, C, Cf = WL["B"], WL["H"], WL["W"], WL["C"], WL["Cf"]
    pipe = ml.PostProcessPipeline(cfg2["cfgp"], (H, W), (H, W), C, Cf, B, ml.DetectionConfig(prefill=True, **KW))
    graph, _ = pipe.capture(*cfg2["ins"], cfg2["d_probs"])
    assert not pipe.ctx.shared and pipe.ctx.frozen
    for _ in range(3):
        pipe.pasted.fill_(9)
        graph.replay()
    torch.cuda.synchronize()
    det_i, pasted = pipe.result_views()
    assert torch.equal(det_i, cfg2["det_i"]) and torch.equal(pasted, cfg2["pasted"])


def test_planar_mask_layout_full_size(cfg2):
    ml = cfg2["ml"]
    B, H, W, C, Cf = WL["B"], WL["H"], WL["W"], WL["C"], WL["Cf"]
    planar = _d(cfg2["probs"].transpose(0, 1, 4, 2, 3))
    pipe = ml.PostProcessPipeline(cfg2["cfgp"], (H, W), (H, W), C, Cf, B,
                                  ml.DetectionConfig(mask_layout="planar", **KW))
    rois = pipe.detect_and_align(*cfg2["ins"])
    pipe.trim_and_paste(rois, planar)
    det_i, pasted = pipe.result_views()
    assert torch.equal(det_i, cfg2["det_i"]) and torch.equal(pasted, cfg2["pasted"])


@pytest.mark.parametrize("planar", [False, True])
@pytest.mark.parametrize("shape", [(28, 28, 5), (28, 28, 6), (14, 14, 3), (32, 32, 8), (7, 9, 4), (33, 20, 2), (28, 28, 1)])
def test_tail_bulk_copy_equals_register_gather(shape, planar):
    """tail_prep_tma_kernel (one cp.async.bulk per RoI block) against tail_prep_kernel, every tile shape the bit
    path takes (mask width <= 32), both layouts; shapes whose block is not a multiple of 16 bytes fall back."""
    import masklab_b200 as ml
    mh, mw, C = shape
    B, H, W, Cf = 3, 96, 160, 8
    cfgp = synth.prior_config(strides=(8, 16, 32))
    N = synth.num_anchors(cfgp, H, W)
    loc, cls = synth.head_tensors(B, N, C, mu=-3.0, seed=77)
    cls[1] = 0
    fmaps = synth.fpn_maps(B, H, W, Cf, seed=78)
    kw = dict(min_confidence=0.05, nms_iou_threshold=0.4, post_iou_threshold=0.65, nms_max_output_size=40,
              max_k=2, base_size=36)
    outs = []
    for tma in ("1", "0"):
        os.environ["MLP_TAIL_TMA"] = tma
        try:
            cfg = ml.DetectionConfig(mask_size=(mh, mw), mask_layout="planar" if planar else "interleaved", **kw)
            pipe = ml.PostProcessPipeline(cfgp, (H, W), (2 * H, 2 * W), C, Cf, B, cfg)
            rois = pipe.detect_and_align(_d(loc), _d(cls), [_d(f) for f in fmaps])
            _, R = rois.shapes()
            probs = synth.mask_probs(B, R, C, mask_hw=(mh, mw), seed=79)
            pipe.trim_and_paste(rois, _d(probs.transpose(0, 1, 4, 2, 3) if planar else probs))
            det_i, pasted = pipe.result_views()
            outs.append((det_i.cpu().numpy(), pasted.cpu().numpy()))
        finally:
            os.environ.pop("MLP_TAIL_TMA", None)
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
    want = co.full_path(loc, cls, fmaps, lambda f, b: probs, cfgp, (H, W), (2 * H, 2 * W), binary=True, **kw)
    assert np.array_equal(outs[0][0], want["det_i"]) and np.array_equal(outs[0][1], want["binary"])


def test_frozen_ctx_refuses_to_regrow_scratch():
    """a captured graph bakes scratch pointers in.  capture() moves the pipeline to a private
    ctx and freezes it; a later call on that ctx with a larger shape must fail loudly, not free the arena."""
    import masklab_b200 as ml
    from masklab_b200 import runtime as rt
    B, H, W, C, Cf = 2, 64, 96, 3, 8
    cfgp = synth.prior_config(strides=(8, 16, 32))
    N = synth.num_anchors(cfgp, H, W)
    loc, cls = synth.head_tensors(B, N, C, mu=-3.0, seed=5)
    fmaps = synth.fpn_maps(B, H, W, Cf, seed=6)
    kw = dict(min_confidence=0.05, nms_iou_threshold=0.4, post_iou_threshold=0.65, nms_max_output_size=20,
              max_k=2, base_size=36)
    pipe = ml.PostProcessPipeline(cfgp, (H, W), (H, W), C, Cf, B, ml.DetectionConfig(**kw))
    assert pipe.ctx.shared
    ins = (_d(loc), _d(cls), [_d(f) for f in fmaps])
    rois = pipe.detect_and_align(*ins)
    _, R = rois.shapes()
    probs = _d(synth.mask_probs(B, R, C, seed=7))
    graph, _ = pipe.capture(*ins, probs)
    assert not pipe.ctx.shared and pipe.ctx.frozen
    before = pipe.ctx.scratch_bytes()
    # same ctx, much larger shape: the DETECT arena would have to grow
    big = ml.PostProcessPipeline(cfgp, (8 * H, 8 * W), (H, W), C, Cf, B, ml.DetectionConfig(**kw))
    big.ctx = pipe.ctx
    N2 = synth.num_anchors(cfgp, 8 * H, 8 * W)
    loc2, cls2 = synth.head_tensors(B, N2, C, mu=-5.0, seed=8)
    fm2 = synth.fpn_maps(B, 8 * H, 8 * W, Cf, seed=9)
    with pytest.raises(rt.MaskLabError) as e:
        big.detect_and_align(_d(loc2), _d(cls2), [_d(f) for f in fm2])
    assert e.value.code == rt.MLP_EFROZEN and "frozen" in str(e.value)
    assert pipe.ctx.scratch_bytes() == before
    graph.replay()                                              # the graph still runs on intact scratch
    torch.cuda.synchronize()
    det_i, pasted = pipe.result_views()
    want = co.full_path(loc, cls, fmaps, lambda f, b: probs.cpu().numpy(), cfgp, (H, W), (H, W), binary=True, **kw)
    assert np.array_equal(det_i.cpu().numpy(), want["det_i"]) and np.array_equal(pasted.cpu().numpy(), want["binary"])
    pipe.ctx.freeze(False)                                      # thawed: growth is allowed again
    big.detect_and_align(_d(loc2), _d(cls2), [_d(f) for f in fm2])
    torch.cuda.synchronize()
    # the shared ctx was never frozen: stand-alone layers keep working whatever they need
    assert not rt.Context.get().frozen


def test_profiling_survives_a_capture():
    """event brackets are skipped while the stream is capturing, so profile_read works afterwards."""
    import masklab_b200 as ml
    B, H, W, C, Cf = 2, 64, 96, 3, 8
    cfgp = synth.prior_config(strides=(8, 16, 32))
    N = synth.num_anchors(cfgp, H, W)
    loc, cls = synth.head_tensors(B, N, C, mu=-3.0, seed=15)
    fmaps = synth.fpn_maps(B, H, W, Cf, seed=16)
    kw = dict(min_confidence=0.05, nms_iou_threshold=0.4, post_iou_threshold=0.65, nms_max_output_size=20,
              max_k=2, base_size=36)
    pipe = ml.PostProcessPipeline(cfgp, (H, W), (H, W), C, Cf, B, ml.DetectionConfig(**kw), private_context=True)
    ins = (_d(loc), _d(cls), [_d(f) for f in fmaps])
    rois = pipe.detect_and_align(*ins)
    _, R = rois.shapes()
    probs = _d(synth.mask_probs(B, R, C, seed=17))
    pipe.ctx.profile(True)
    graph, _ = pipe.capture(*ins, probs)
    graph.replay()
    rois = pipe.detect_and_align(*ins)
    pipe.trim_and_paste(rois, probs)
    stages = pipe.ctx.profile_read()
    assert stages["paste"][1] >= 1 and stages["paste"][0] > 0.0
    pipe.ctx.profile(False)


def test_atomic_finalisers_are_deterministic(cfg2):
    """box_summary_kernel ("last CTA to arrive writes the row") and jpeg_place_kernel (atomic first/last words)
    100 times over the same batch: every replay must give the same bytes."""
    ml = cfg2["ml"]
    B, H, W, C, Cf = WL["B"], WL["H"], WL["W"], WL["C"], WL["Cf"]
    pipe = ml.PostProcessPipeline(cfg2["cfgp"], (H, W), (H, W), C, Cf, B, ml.DetectionConfig(**KW), private_context=True)
    seg = _d(synth.semantic_map(B, H, W, seed=31))
    img = _d(synth.road_frames(B, H, W, seed=32))
    colors = [[192, 32, 128], [160, 96, 0], [96, 0, 128], [32, 96, 192], [96, 32, 128]]
    sem = [[64, 0, 128], [128, 96, 0], [128, 192, 0]]
    graph, _ = pipe.capture_serving(*cfg2["ins"], cfg2["d_probs"], seg, img, colors, 0.3, semantic_colors=sem,
                                    semantic_alpha=0.3, boxes=True)
    ref = None
    for i in range(100):
        graph.replay()
        torch.cuda.synchronize()
        cur = (pipe.summary.clone(), pipe.jpeg_files.clone(), pipe.jpeg_len.clone(), pipe.vis.clone())
        if ref is None:
            ref = cur
            continue
        Mo = int(pipe.summary_m.item())
        n = B * Mo * 11
        assert torch.equal(cur[0][:n].view(torch.int32), ref[0][:n].view(torch.int32)), f"summary differs at replay {i}"
        assert torch.equal(cur[2], ref[2]), f"JPEG lengths differ at replay {i}"
        for b in range(B):
            L = int(cur[2][b])
            assert torch.equal(cur[1][b, :L], ref[1][b, :L]), f"JPEG bytes of frame {b} differ at replay {i}"
        assert torch.equal(cur[3], ref[3])


def test_speculative_fill_follows_a_changing_m():
    """The background fill is sized by the PREVIOUS batch's M before this batch's counts exist, and completed after
    the NMS kernels.  Batches whose M grows, shrinks, drops to 1 (no detections at all) and grows again must all
    equal the plain pipeline bit for bit - stale boxes of earlier batches included (same output buffer)."""
    import masklab_b200 as ml
    B, H, W, C, Cf = 3, 96, 160, 4, 8
    cfgp = synth.prior_config(strides=(8, 16, 32))
    N = synth.num_anchors(cfgp, H, W)
    kw = dict(min_confidence=0.05, nms_iou_threshold=0.4, post_iou_threshold=0.65, nms_max_output_size=400,
              max_k=2, base_size=36)
    fmaps = [_d(f) for f in synth.fpn_maps(B, H, W, Cf, seed=51)]
    plain = ml.PostProcessPipeline(cfgp, (H, W), (2 * H, 2 * W), C, Cf, B, ml.DetectionConfig(**kw))
    fast = ml.PostProcessPipeline(cfgp, (H, W), (2 * H, 2 * W), C, Cf, B, ml.DetectionConfig(prefill=True, **kw))
    fast.pasted.fill_(5)
    seen = []
    for step, mu in enumerate((-6.0, -3.0, -5.0, None, -4.0, -6.5)):
        loc, cls = synth.head_tensors(B, N, C, mu=mu if mu is not None else -9.0, seed=60 + step)
        if mu is None:
            cls[:] = 0                                           # nothing passes the threshold: M = 1, all padding
        outs = []
        for pipe in (plain, fast):
            rois = pipe.detect_and_align(_d(loc), _d(cls), fmaps)
            _, R = rois.shapes()
            pipe.trim_and_paste(rois, _d(synth.mask_probs(B, R, C, seed=70 + step)))
            det_i, pasted = pipe.result_views()
            outs.append((det_i.clone(), pasted.clone()))
        assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1]), (step, mu)
        seen.append(int(fast.trim_m.item()))
    assert len(set(seen)) >= 3 and min(seen) == 1, seen            # M really moved both ways


@pytest.mark.parametrize("seed", [0, 3, 7, 11])
def test_fused_cross_class_nms_equals_oracle(seed):
    """MLP_NMS_FUSE=1: the last class CTA of an image to finish runs the image's cross-class NMS (no second kernel).
    Same detections, RoIs and masks as the C oracle, incl. images without candidates and single-class images."""
    import masklab_b200 as ml
    rng = np.random.default_rng(seed)
    B, H, W, C, Cf = int(rng.integers(1, 6)), 96, 160, int(rng.integers(1, 7)), 8
    cfgp = synth.prior_config(strides=(8, 16, 32))
    N = synth.num_anchors(cfgp, H, W)
    loc, cls = synth.head_tensors(B, N, C, mu=float(rng.choice([-4.0, -3.0, -2.0])), seed=100 + seed)
    if B > 1:
        cls[0] = 0
    if B > 2 and C > 1:
        cls[2, :, 1:] = 0
    fmaps = synth.fpn_maps(B, H, W, Cf, seed=200 + seed)
    kw = dict(min_confidence=0.05, nms_iou_threshold=0.4, post_iou_threshold=0.65,
              nms_max_output_size=int(rng.choice([5, 40, 300])), max_k=2, base_size=36)
    os.environ["MLP_NMS_FUSE"] = "1"
    try:
        pipe = ml.PostProcessPipeline(cfgp, (H, W), (H, W), C, Cf, B, ml.DetectionConfig(**kw))
        rois = pipe.detect_and_align(_d(loc), _d(cls), [_d(f) for f in fmaps])
        _, R = rois.shapes()
        probs = synth.mask_probs(B, R, C, seed=300 + seed)
        pipe.trim_and_paste(rois, _d(probs))
        det_i, pasted = pipe.result_views()
        M = int(rois.m_dev.item())
        got_det = rois.det[:, :M].cpu().numpy()
        # stand-alone layer (boxes given, no prior decode) through the same kernels
        boxes = co.restore_boxes(loc, np.broadcast_to(co.prior_layer(cfgp, H, W)[None], (B, N, 4)))
        layer = ml.DetectionProposal(**{k: kw[k] for k in ("min_confidence", "nms_iou_threshold", "post_iou_threshold",
                                                           "nms_max_output_size")})([_d(cls), _d(boxes), None])
    finally:
        os.environ.pop("MLP_NMS_FUSE", None)
    want = co.full_path(loc, cls, fmaps, lambda f, b: probs, cfgp, (H, W), (H, W), binary=True, **kw)
    assert np.array_equal(got_det, want["proposed"]) and np.array_equal(layer.cpu().numpy(), want["proposed"])
    assert np.array_equal(det_i.cpu().numpy(), want["det_i"]) and np.array_equal(pasted.cpu().numpy(), want["binary"])
