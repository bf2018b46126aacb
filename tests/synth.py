"""Seeded synthetic inputs of SURVEY.md §8(d), shared by the tests and bench.py (NumPy, CPU).

Everything is float32 NHWC.  `head_tensors` follows the classification-head statistics of
the reference (bias init -log(99), /root/reference/engine/layers/detection.py:200): scores
are sigmoid(N(mu,1)); mu=-5.8 is the "realistic" setting (~0.21% of scores pass 0.05),
mu=-5.0 the "stress" one (~2%).
"""
import numpy as np

F32 = np.float32

DEFAULT_STRIDES = (8, 16, 32, 64, 128)
DEFAULT_SCALES = (2 ** 0, 2 ** (1 / 3), 2 ** (2 / 3))
DEFAULT_RATIOS = (1 / 3, 1 / 2, 1, 2, 3)
A9_RATIOS = (1 / 2, 1, 2)


def prior_config(strides=DEFAULT_STRIDES, scales=DEFAULT_SCALES, ratios=DEFAULT_RATIOS):
    strides = [int(s) for s in strides]
    return dict(strides=strides, sizes=[4 * s for s in strides], pr_scales=list(scales),
                pr_ratios=list(ratios))


def num_anchors(prior_cfg, H, W, padding="same"):
    A = len(prior_cfg["pr_scales"]) * len(prior_cfg["pr_ratios"])
    n = 0
    for s in prior_cfg["strides"]:
        hf = -(-H // s) if padding == "same" else H // s
        wf = -(-W // s) if padding == "same" else W // s
        n += hf * wf * A
    return n


def head_tensors(B, N, C, mu=-5.8, seed=0):
    """loc_pred [B,N,4] ~ N(0,0.5^2) clipped to +-2; cls_pred [B,N,C] = sigmoid(N(mu,1))."""
    rng = np.random.default_rng(seed)
    loc = np.clip(rng.standard_normal((B, N, 4), dtype=F32) * F32(0.5), -2, 2).astype(F32)
    z = rng.standard_normal((B, N, C), dtype=F32) + F32(mu)
    cls = (1.0 / (1.0 + np.exp(-z.astype(np.float64)))).astype(F32)
    return loc, cls


def fpn_maps(B, H, W, Cf, strides=(8, 16, 32), seed=1, padding="same"):
    rng = np.random.default_rng(seed)
    out = []
    for s in strides:
        hf = -(-H // s) if padding == "same" else H // s
        wf = -(-W // s) if padding == "same" else W // s
        out.append(rng.standard_normal((B, hf, wf, Cf), dtype=F32))
    return out


def mask_probs(B, R, C, mask_hw=(28, 28), seed=2):
    """Stand-in for the MaskSubNet output: U(0,1) [B,R,mh,mw,C]."""
    rng = np.random.default_rng(seed)
    return rng.random((B, R, mask_hw[0], mask_hw[1], C), dtype=F32)


def detections(B, M, C, H, W, seed=3, lo=16.0, hi=400.0, pad_tail=0):
    """Exactly M synthetic detections per image (cfg-4): w,h ~ logU(lo,hi), centres uniform
    in-frame, classes uniform, scores U(0.5,1) sorted descending like NMS output; the last
    `pad_tail` rows of every image are -1 padding."""
    rng = np.random.default_rng(seed)
    w = np.exp(rng.uniform(np.log(lo), np.log(hi), (B, M))).astype(F32)
    h = np.exp(rng.uniform(np.log(lo), np.log(hi), (B, M))).astype(F32)
    cx = rng.uniform(0, W, (B, M)).astype(F32)
    cy = rng.uniform(0, H, (B, M)).astype(F32)
    cls = rng.integers(0, C, (B, M)).astype(F32)
    score = -np.sort(-rng.uniform(0.5, 1.0, (B, M)), axis=1)
    det = np.stack([cx, cy, w, h, cls, score.astype(F32)], axis=-1).astype(F32)
    if pad_tail:
        det[:, M - pad_tail:] = -1
    return det


def semantic_map(B, PH, PW, S=3, seed=4, crack=True, road=True):
    """Stand-in for UpSampleOutput's thresholded semantic map, int32 {0,1} [B,PH,PW,S]:
    channel 1 (my_road) is a noisy trapezoid widening towards the bottom of the frame with a few
    gaps (rows without road, single-pixel rows), channel 2 (crack) a few thin random strokes,
    channel 0 background."""
    rng = np.random.default_rng(seed)
    seg = np.zeros((B, PH, PW, S), dtype=np.int32)
    ys = np.arange(PH)
    for b in range(B):
        if road:
            top = int(PH * rng.uniform(0.25, 0.45))
            cx = PW * rng.uniform(0.4, 0.6)
            half_top, half_bot = PW * rng.uniform(0.02, 0.06), PW * rng.uniform(0.25, 0.45)
            t = np.clip((ys - top) / max(PH - 1 - top, 1), 0, 1)
            left = np.round(cx - (half_top + (half_bot - half_top) * t) + rng.normal(0, 1.5, PH)).astype(int)
            right = np.round(cx + (half_top + (half_bot - half_top) * t) + rng.normal(0, 1.5, PH)).astype(int)
            for y in range(top, PH):
                if rng.random() < 0.03:
                    continue                                   # a row without road pixels
                lo, hi = max(left[y], 0), min(right[y], PW - 1)
                if rng.random() < 0.02:
                    hi = lo                                    # single-pixel row (x_min == x_max)
                if hi >= lo:
                    seg[b, y, lo:hi + 1, 1] = 1
                    holes = rng.random(hi + 1 - lo) < 0.01     # interior holes do not move the borders
                    holes[[0, -1]] = False
                    seg[b, y, lo:hi + 1, 1][holes] = 0
        if crack and S > 2:
            for _ in range(int(rng.integers(0, 3))):
                y0, x0 = int(rng.integers(PH // 2, PH)), int(rng.integers(0, PW))
                for k in range(int(rng.integers(5, 40))):
                    y, x = y0 + k // 2, x0 + k + int(rng.integers(-1, 2))
                    if 0 <= y < PH and 0 <= x < PW:
                        seg[b, y, x, 2] = 1
        seg[b, :, :, 0] = 1 - np.clip(seg[b, :, :, 1:].sum(-1), 0, 1)
    return seg


def int_detections(B, M, C, PH, PW, seed=5, pad_tail=0):
    """det_outs as UpSampleOutput returns them: int32 [B,M,6] (cx,cy,w,h,class,conf*100)."""
    rng = np.random.default_rng(seed)
    det = np.stack([rng.integers(0, PW, (B, M)), rng.integers(0, PH, (B, M)),
                    rng.integers(2, max(3, PW // 3), (B, M)), rng.integers(2, max(3, PH // 3), (B, M)),
                    rng.integers(0, C, (B, M)), rng.integers(51, 100, (B, M))], axis=-1).astype(np.int32)
    if pad_tail:
        det[:, M - pad_tail:] = np.array([-1, -1, -1, -1, -1, -100], dtype=np.int32)
    return det


def road_frames(B, PH, PW, seed=0):
    """uint8 [B,PH,PW,3] camera-like frames: smooth shading, a textured band and mild sensor noise (uniform noise
    would be meaningless input for the JPEG encode at the end of the serving graph)."""
    rng = np.random.default_rng(seed)
    yy = np.arange(PH, dtype=np.float32)[:, None]
    xx = np.arange(PW, dtype=np.float32)[None, :]
    out = np.empty((B, PH, PW, 3), dtype=np.uint8)
    for b in range(B):
        base = np.stack([120 + 70 * np.sin(xx / (41.0 + b) + yy / 67.0), 115 + 60 * np.cos(xx / 93.0 - yy / (29.0 + b)),
                         95 + yy * (110.0 / PH) + 20 * np.sin(xx / 11.0)], -1)
        base += rng.normal(0, 3.0, base.shape).astype(np.float32)
        out[b] = np.clip(base, 0, 255).astype(np.uint8)
    return out
