"""Times the two consumers of the pasted masks on cfg-2 both ways: drop-in layers reading the
materialised [B,M,PH,PW] masks (uint8 and float32) vs the tile path that never writes them."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import synth
import bench
import masklab_b200 as ml
from masklab_b200.layers import summary as ls


def timeit(fn, n=20):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


wl = bench.WORKLOADS["cfg2"]
B, C, PH, PW = wl["B"], wl["C"], wl["PH"], wl["PW"]
cfgp, N, loc, cls, fmaps = bench.make_inputs(wl, B, seed=100)
d = lambda a: torch.from_numpy(a).cuda()
d_loc, d_cls, d_fm = d(loc), d(cls), [d(f) for f in fmaps]
seg = d(synth.semantic_map(B, PH, PW, seed=500))
img = d(np.random.default_rng(600).integers(0, 256, (B, PH, PW, 3)).astype(np.uint8))
out = {}
for mode in ("uint8", "float32"):
    cfg = ml.DetectionConfig(paste_output=mode, **bench.kwargs_of(wl))
    pipe = ml.PostProcessPipeline(cfgp, (wl["H"], wl["W"]), (PH, PW), C, wl["Cf"], B, cfg)
    r = pipe.detect_and_align(d_loc, d_cls, d_fm)
    mf, R = r.shapes()
    masks = d(synth.mask_probs(B, R, C, seed=300))
    pipe.trim_and_paste(r, masks)
    det_i, pasted = pipe.result_views()
    det_i = det_i.contiguous()
    out[f"paste_{mode}_ms"] = timeit(lambda: pipe.trim_and_paste(r, masks))
    out[f"SummaryOutput_on_{mode}_masks_ms"] = timeit(lambda: ml.SummaryOutput()([det_i, seg, pasted]))
    out[f"DrawInstance_on_{mode}_masks_ms"] = timeit(lambda: ml.DrawInstance(bench.INST_COLORS[:C])([img, det_i, pasted]))
    if mode == "uint8":
        out["tiles_trim_and_summarize_ms"] = timeit(lambda: pipe.trim_and_summarize(r, masks, seg))
        out["tiles_draw_instance_and_semantic_ms"] = timeit(
            lambda: pipe.draw(r, masks, img, bench.INST_COLORS[:C], 0.3, seg_outs=seg, semantic_colors=bench.SEM_COLORS))
    del pipe, pasted
    torch.cuda.empty_cache()
sem_lo = torch.rand((B, 540, 960, 3), device="cuda")
frames_hi = torch.randint(0, 256, (B, 1080, 1920, 3), dtype=torch.uint8, device="cuda")
out["UpSampleOutput_semantic_540x960_to_1080x1920_ms"] = timeit(
    lambda: ml.layers.misc.resize_bilinear(ml.Context.get(), sem_lo, 1080, 1920, threshold=True))
out["DownSampleInput_1080x1920_to_540x960_ms"] = timeit(lambda: ml.DownSampleInput((540, 960))(frames_hi))
# training-side target assignment at cfg-2 scale: 32 images, 163,680 priors, 50 ground truths each
prb = ml.PriorLayer(cfgp)(torch.zeros((B, wl["H"], wl["W"], 3), dtype=torch.uint8, device="cuda"))   # [B,N,4] int32
gt = torch.from_numpy(synth.detections(B, 50, C, wl["H"], wl["W"], seed=9, lo=12.0, hi=300.0, pad_tail=5)).cuda()
gt[..., 5] = torch.where(gt[..., 0] == -1, -1.0, 1.0)
out["AssignBoxes_B32_N163680_G50_ms"] = timeit(lambda: ml.AssignBoxes(C)([gt, prb]))
gm = torch.zeros((B, 50, wl["H"], wl["W"]), device="cuda")
out["AssignMasks_B32_R118_G50_ms"] = timeit(lambda: ml.AssignMasks()([gt[:, :118 % 50 + 50].repeat(1, 3, 1)[:, :118], (28, 28, C), gt, gm]))
out["DetectionIOUMetric_B32_P100_G50_ms"] = timeit(lambda: ml.DetectionIOUMetric()([gt.repeat(1, 2, 1), gt]))
out["DrawSegmentation_ms"] = timeit(lambda: ml.DrawSegmentation(bench.SEM_COLORS)([img, seg]))
for k, v in out.items():
    print(f"{k:48s} {v:8.3f}")
