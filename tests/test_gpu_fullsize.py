"""Full-size parity on the GPU (BASELINE.json configs[1] shapes: B=32, 1024x512, N=163,680, C=5,
Cf=128): the whole fused pipeline against the C restatement of the reference path (oracle/c, itself
bit-identical to the NumPy oracle) on the complete batch, plus size-independent properties of
the outputs (sortedness, NMS invariants, masks confined to their boxes, idempotence)."""
import numpy as np
import pytest
import torch

import synth
from oracle import c_oracle as co

pytestmark = pytest.mark.gpu

WL = dict(B=32, H=512, W=1024, C=5, Cf=128, mu=-5.8, min_confidence=0.05, nms_iou_threshold=0.4,
          post_iou_threshold=0.6, nms_max_output_size=100, max_k=2, base_size=36)
KW = {k: WL[k] for k in ("min_confidence", "nms_iou_threshold", "post_iou_threshold",
                         "nms_max_output_size", "max_k", "base_size")}


@pytest.fixture(scope="module")
def run():
    import masklab_b200 as ml
    B, H, W, C, Cf = WL["B"], WL["H"], WL["W"], WL["C"], WL["Cf"]
    cfgp = synth.prior_config()
    N = synth.num_anchors(cfgp, H, W)
    loc, cls = synth.head_tensors(B, N, C, mu=WL["mu"], seed=1234)
    cls[5] = 0                                   # one frame without any detection
    cls[9, :, 1:] = 0                            # one frame with a single class
    fmaps = synth.fpn_maps(B, H, W, Cf, seed=1235)
    pipe = ml.PostProcessPipeline(cfgp, (H, W), (H, W), C, Cf, B, ml.DetectionConfig(**KW))
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    rois = pipe.detect_and_align(d(loc), d(cls), [d(f) for f in fmaps])
    crops, roi_boxes = pipe.roi_views(rois)
    R = roi_boxes.shape[1]
    probs = synth.mask_probs(B, R, C, seed=1236)
    pipe.trim_and_paste(rois, d(probs))
    det_i, pasted = pipe.result_views()
    M = int(rois.m_dev.item())
    got = dict(det=rois.det[:, :M].cpu().numpy(), keep=rois.keep[:, :M].cpu().numpy(),
               counts=rois.counts.cpu().numpy(), roi_boxes=roi_boxes.cpu().numpy(),
               crops=[c.cpu().numpy() for c in crops], det_i=det_i.cpu().numpy(),
               pasted=pasted.cpu().numpy())
    want = co.full_path(loc, cls, fmaps, lambda f, b: probs, cfgp, (H, W), (H, W), binary=True, **KW)
    return dict(got=got, want=want, loc=loc, cls=cls, pipe=pipe, ml=ml, inputs=(d(loc), d(cls), [d(f) for f in fmaps]),
                probs=d(probs))


def test_full_batch_equals_c_oracle(run):
    g, w = run["got"], run["want"]
    assert g["det"].shape == w["proposed"].shape == (32, 100, 6)
    assert np.array_equal(g["det"], w["proposed"])                 # boxes, class ids, scores
    assert np.array_equal(g["roi_boxes"], w["roi_boxes"])
    for a, b in zip(g["crops"], w["roi_fmaps"]):
        assert a.shape == b.shape
        assert np.abs(a - b).max() <= 1e-4                         # north-star tolerance
        assert np.array_equal(a, b)                                # and exact
    assert np.array_equal(g["det_i"], w["det_i"])
    assert g["pasted"].dtype == np.uint8 and np.array_equal(g["pasted"], w["binary"])   # 1.68 GB of masks


def test_detection_invariants_at_full_size(run):
    det, keep, counts, cls = run["got"]["det"], run["got"]["keep"], run["got"]["counts"], run["cls"]
    assert counts[5] == 0 and np.all(det[5] == -1)
    assert set(det[9, :counts[9], 4].tolist()) <= {0.0}
    for b in range(det.shape[0]):
        n = counts[b]
        assert 0 <= n <= 100 and np.all(det[b, n:] == -1) and np.all(keep[b, n:] == -1)
        s = det[b, :n, 5]
        assert np.all(s[:-1] >= s[1:])                             # selection order = descending score
        assert np.all(s >= np.float32(0.05))
        # kept (anchor, class) really carries that score, and no (anchor,class) is kept twice
        assert np.array_equal(cls[b, keep[b, :n, 0], keep[b, :n, 1]], s)
        assert len({(int(a), int(c)) for a, c in keep[b, :n]}) == n
        # no surviving pair exceeds the cross-class IoU threshold (NormalizeBoxes + TF IoU, f32)
        cx, cy, w, h = (det[b, :n, i] for i in range(4))
        y1, x1, y2, x2 = cy - h / 2, cx - w / 2, cy + h / 2, cx + w / 2
        area = (y2 - y1) * (x2 - x1)
        ih = np.maximum(np.minimum(y2[:, None], y2[None]) - np.maximum(y1[:, None], y1[None]), 0)
        iw = np.maximum(np.minimum(x2[:, None], x2[None]) - np.maximum(x1[:, None], x1[None]), 0)
        inter = (ih * iw).astype(np.float32)
        with np.errstate(divide="ignore", invalid="ignore"):
            iou = inter / (area[:, None] + area[None] - inter)
        np.fill_diagonal(iou, 0)
        assert not np.any(iou > np.float32(0.6))


def test_masks_confined_to_boxes_at_full_size(run):
    det_i, pasted = run["got"]["det_i"], run["got"]["pasted"]
    B, M, PH, PW = pasted.shape
    thr = 50 if det_i[..., 5].max() > 50 else -100
    rows_any = pasted.any(axis=3)
    cols_any = pasted.any(axis=2)
    for b in range(B):
        for j in range(M):
            if det_i[b, j, 5] < thr:                               # filtered row -> all zeros
                assert not rows_any[b, j].any()
                continue
            cx, cy, w, h = np.maximum(det_i[b, j, :4], 1)
            ys = np.nonzero(rows_any[b, j])[0]
            xs = np.nonzero(cols_any[b, j])[0]
            if ys.size:
                assert ys[0] >= np.ceil(cy - h / 2) and ys[-1] < np.ceil(cy + h / 2)
                assert xs[0] >= np.ceil(cx - w / 2) and xs[-1] < np.ceil(cx + w / 2)


def test_idempotent_and_deterministic_at_full_size(run):
    pipe = run["pipe"]
    loc, cls, fmaps = run["inputs"]
    r = pipe.detect_and_align(loc, cls, fmaps)
    pipe.trim_and_paste(r, run["probs"])
    det_i, pasted = pipe.result_views()
    assert np.array_equal(det_i.cpu().numpy(), run["got"]["det_i"])
    assert torch.equal(pasted.cpu(), torch.from_numpy(run["got"]["pasted"]))
    # NMS is idempotent: feeding the kept boxes back keeps every one of them, in the same order
    ml = run["ml"]
    det, counts = run["got"]["det"], run["got"]["counts"]
    b = int(np.argmax(counts))
    n = counts[b]
    boxes = torch.from_numpy(det[b:b + 1, :n, :4].copy()).cuda()
    C = 5
    cls2 = np.zeros((1, n, C), np.float32)
    cls2[0, np.arange(n), det[b, :n, 4].astype(int)] = det[b, :n, 5]
    again = ml.DetectionProposal(**{k: KW[k] for k in ("min_confidence", "nms_iou_threshold",
                                                       "post_iou_threshold", "nms_max_output_size")})(
        [torch.from_numpy(cls2).cuda(), boxes, None]).cpu().numpy()
    assert again.shape[1] == n and np.array_equal(again[0], det[b, :n])


def test_cuda_graph_replay_reproduces_the_batch(run):
    """The whole batch captured as one CUDA graph (no host round trip inside the path) gives the same
    bytes, also after the inputs are refilled with another batch and back."""
    pipe = run["pipe"]
    d_loc, d_cls, d_fm = run["inputs"]
    ref_det = pipe.det.clone()
    ref_pasted = pipe.pasted.clone()
    graph, rois = pipe.capture(d_loc, d_cls, d_fm, run["probs"])
    keep_cls = d_cls.clone()
    d_cls.copy_(torch.roll(keep_cls, 1, dims=0))             # another batch through the same buffers
    graph.replay()
    torch.cuda.synchronize()
    assert not torch.equal(pipe.det, ref_det)
    d_cls.copy_(keep_cls)
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(pipe.det, ref_det) and torch.equal(pipe.pasted, ref_pasted)
