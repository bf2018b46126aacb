"""Randomised soak of the whole pipeline against the C restatement of the reference path: the configuration generator of
tests/test_gpu_random.py over many more seeds, with exact score ties injected (quantised scores: the anchor-index tie
rule decides), every tail variant (fused / staged / background fill / planar masks / box-clipped) and both forms of
the cross-class NMS (own kernel, fused behind the per-class kernel).

    python tools/fuzz_gpu.py [first_seed] [count]          # prints one line per failure and a summary; exit 1 on failure
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import synth  # noqa: E402
import test_gpu_random as tr  # noqa: E402
from oracle import c_oracle as co  # noqa: E402
import masklab_b200 as ml  # noqa: E402


def run(seed):
    c = tr._case(5000 + seed)
    cfgp = synth.prior_config(strides=c["strides"], scales=c["scales"], ratios=c["ratios"])
    B, H, W, C, Cf = c["B"], c["H"], c["W"], c["C"], c["Cf"]
    N = synth.num_anchors(cfgp, H, W, c["padding"])
    if N == 0:
        return "skip"
    if c["kw"]["min_confidence"] == 0.0 and B * N * C > 60000:
        c["kw"]["min_confidence"] = 0.05
    loc, cls = synth.head_tensors(B, N, C, mu=c["mu"], seed=seed)
    if seed % 2 == 0:                                        # exact ties: scores on a grid of 1/64
        cls = (np.round(cls * 64) / 64).astype(np.float32)
    if seed % 5 == 0 and B > 1:
        cls[B - 1] = 0
    fmaps = synth.fpn_maps(B, H, W, Cf, strides=c["strides"][:c["kw"]["max_k"] + 1], seed=seed + 1, padding=c["padding"])
    if any(f.size == 0 for f in fmaps):
        return "skip"                                        # a pyramid level without a single cell ('valid' padding, tiny frame)
    probs = {}

    def head(roi_fmaps, roi_boxes):
        probs["m"] = synth.mask_probs(B, roi_boxes.shape[1], C, mask_hw=c["mask"], seed=seed + 2)
        return probs["m"]

    want = co.full_path(loc, cls, fmaps, head, cfgp, (H, W), (c["PH"], c["PW"]), crop_size=c["crop"],
                        padding=c["padding"], binary=True, **c["kw"])
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    bad = []
    for fuse in ("0", "1"):
        os.environ["MLP_NMS_FUSE"] = fuse
        for fused, prefill, planar in ((True, False, False), (False, False, False), (True, True, False), (True, False, True)):
            if fuse == "1" and not fused:
                continue
            cfg = ml.DetectionConfig(crop_size=c["crop"], mask_size=c["mask"], padding=c["padding"], fused=fused,
                                     prefill=prefill, mask_layout="planar" if planar else "interleaved", **c["kw"])
            pipe = ml.PostProcessPipeline(cfgp, (H, W), (c["PH"], c["PW"]), C, Cf, B, cfg)
            pipe.pasted.fill_(7)
            rois = pipe.detect_and_align(d(loc), d(cls), [d(f) for f in fmaps])
            crops, roi_boxes = pipe.roi_views(rois)
            M = int(rois.m_dev.item())
            tag = f"seed {seed} fuse {fuse} fused {fused} prefill {prefill} planar {planar}"
            if not np.array_equal(rois.det[:, :M].cpu().numpy(), want["proposed"]):
                bad.append(tag + ": detections")
                continue
            if not np.array_equal(roi_boxes.cpu().numpy(), want["roi_boxes"]):
                bad.append(tag + ": roi boxes")
            if not all(np.array_equal(g.cpu().numpy(), w) for g, w in zip(crops, want["roi_fmaps"])):
                bad.append(tag + ": roi features")
            masks = probs["m"].transpose(0, 1, 4, 2, 3) if planar else probs["m"]
            pipe.trim_and_paste(rois, d(masks))
            det_i, pasted = pipe.result_views()
            if not (np.array_equal(det_i.cpu().numpy(), want["det_i"]) and np.array_equal(pasted.cpu().numpy(), want["binary"])):
                bad.append(tag + ": masks")
            if fused and not prefill:
                _, geom, pool, used = pipe.trim_and_clip(rois, d(masks))
                torch.cuda.synchronize()
                Mt = want["binary"].shape[1]
                dense = ml.expand_clipped(geom.cpu().numpy(), pool.cpu().numpy(), (c["PH"], c["PW"]), m_rows=Mt)
                if not np.array_equal(dense, want["binary"]):
                    bad.append(tag + ": clipped masks")
    os.environ.pop("MLP_NMS_FUSE", None)
    return bad


def main():
    first = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    count = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    failures, ran = [], 0
    for s in range(first, first + count):
        r = run(s)
        if r == "skip":
            continue
        ran += 1
        for line in r:
            print("FAIL", line, flush=True)
        failures += r
    print(f"fuzz: {ran} configurations x up to 7 pipeline variants, {len(failures)} failures")
    return 1 if failures else 0


if __name__ == "__main__":
    sys.exit(main())
