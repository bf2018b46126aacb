"""Randomised soak of the two consumers of the mask tiles in the serving graph - SummaryOutput and the
DrawBoxes + DrawInstance + DrawSegmentation overlay - evaluated straight from the tiles (mlp_tile_summary,
mlp_draw_tiles_boxes), against the NumPy oracles fed with the oracle's pasted masks.  Random frame sizes (also no
multiple of 4 / 32 / 64), instance counts from one to crowded, boxes from one pixel to larger than the frame and partly
or wholly off-frame, widths on every lane-layout boundary of the box reduction (8 / 16 / 32 lanes, several 128-column
chunks), tile densities from empty to full, classes inside and outside the colour table, padding rows, confidences on
both sides of the batch threshold, binary / non-binary int32 and float32 semantic maps, uint8 and float32 frames.

    python tools/fuzz_consumers.py [first_seed] [count]      # one line per failure and a summary; exit 1 on failure
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import synth  # noqa: E402
from oracle import draw_oracle as do  # noqa: E402
from oracle import masklab_oracle as mo  # noqa: E402
from oracle import summary_oracle as so  # noqa: E402
import masklab_b200 as ml  # noqa: E402

INST_COLORS = [[192, 32, 128], [160, 96, 0], [96, 0, 128], [32, 96, 192], [96, 32, 128], [10, 200, 30], [250, 250, 5]]
SEM_COLORS = [[64, 0, 128], [128, 96, 0], [128, 192, 0]]
WIDTHS = [1, 2, 7, 8, 9, 15, 16, 17, 31, 32, 33, 47, 63, 64, 65, 96, 127, 128, 129, 191, 255, 256, 257, 300]


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def case(seed):
    rng = np.random.default_rng(seed)
    B = int(rng.integers(1, 4))
    PH = int(rng.choice([17, 33, 48, 64, 70, 96, 130, 200]))
    PW = int(rng.choice([20, 50, 64, 100, 128, 131, 200, 333, 512]))
    M = int(rng.choice([1, 2, 5, 12, 40, 90, 150, 300], p=[.15, .15, .2, .2, .12, .1, .05, .03]))   # > 256: two search rounds
    C = int(rng.integers(1, 8))
    det = np.zeros((B, M, 6), np.int32)
    det[..., 0] = rng.integers(-PW // 4, PW + PW // 4, (B, M))
    det[..., 1] = rng.integers(-PH // 4, PH + PH // 4, (B, M))
    mode = int(rng.integers(0, 3))
    if mode == 0:                                            # widths on the lane-layout boundaries
        det[..., 2] = rng.choice(WIDTHS, (B, M))
        det[..., 3] = rng.integers(1, PH + 10, (B, M))
    elif mode == 1:                                          # small boxes
        det[..., 2] = rng.integers(1, 40, (B, M))
        det[..., 3] = rng.integers(1, 40, (B, M))
    else:                                                    # anything up to twice the frame
        det[..., 2] = rng.integers(1, 2 * PW, (B, M))
        det[..., 3] = rng.integers(1, 2 * PH, (B, M))
    det[..., 4] = rng.integers(0, C + (1 if seed % 7 == 0 else 0), (B, M))      # sometimes a class beyond the table
    det[..., 5] = rng.integers(55, 100, (B, M)) if seed % 3 else rng.integers(20, 100, (B, M))
    if seed % 11 == 0:
        det[..., 5] = rng.integers(0, 50, (B, M))            # nothing above 50: the batch threshold drops to -100
    npad = int(rng.integers(0, M))
    if npad and seed % 2:
        det[B - 1, M - npad:] = np.array([-1, -1, -1, -1, -1, -100], np.int32)
    dens = float(rng.choice([0.0, 0.1, 0.5, 0.9, 1.0]))
    ins = (rng.random((B, M, 28, 28)) < dens).astype(np.int32)
    seg = synth.semantic_map(B, PH, PW, seed=seed + 1)
    kind = seed % 4
    if kind == 1:
        seg = seg.astype(np.int32)
        seg[:, PH // 2:] = rng.integers(-1, 4, seg[:, PH // 2:].shape)
    elif kind == 2:
        seg = (seg * rng.random(seg.shape)).astype(np.float32)
    elif kind == 3:
        seg = seg.astype(np.float32)
    img = rng.integers(0, 256, (B, PH, PW, 3)).astype(np.uint8)
    if seed % 5 == 0:
        img = (img.astype(np.float32) * 1.1 - 9.5)
    return dict(B=B, PH=PH, PW=PW, M=M, C=C, det=det, ins=ins, seg=seg, img=img)


def run(seed):
    c = case(seed)
    det, ins, seg, img = c["det"], c["ins"], c["seg"], c["img"]
    PH, PW, C = c["PH"], c["PW"], c["C"]
    masks = mo.crop_and_pad_mask((PH, PW), det, ins)
    bad = []
    # overlay
    colors = INST_COLORS[:C]
    alpha_s = 0.3 + 0.1 * (seed % 3)
    want = do.draw_segmentation(do.draw_instance(do.draw_boxes(img, det), det, masks, colors, 0.3), seg, SEM_COLORS, alpha_s)
    got = ml.DrawInstance(colors, 0.3).from_tiles([dev(img), dev(det), dev(ins)], seg_outs=dev(seg),
                                                  semantic_colors=SEM_COLORS, semantic_alpha=alpha_s,
                                                  boxes=True).cpu().numpy()
    if not np.array_equal(got, want):
        bad.append("overlay (%d bytes differ)" % int((got != want).sum()))
    want_i = do.draw_instance(img, det, masks, colors, 0.3)
    got_i = ml.DrawInstance(colors, 0.3).from_tiles([dev(img), dev(det), dev(ins)]).cpu().numpy()
    if not np.array_equal(got_i, want_i):
        bad.append("instance overlay (%d bytes differ)" % int((got_i != want_i).sum()))
    # summary (integral semantic maps only: SummaryOutput casts the map to int)
    if seg.dtype == np.int32 and seg.min() >= 0 and seg.max() <= 1:
        want_s = so.summary_output(det, seg, masks)
        got_s = ml.SummaryOutput().from_tiles([dev(det), dev(seg), dev(ins)]).cpu().numpy()
        if got_s.shape != want_s.shape:
            bad.append("summary shape %s vs %s" % (got_s.shape, want_s.shape))
        else:
            if not (np.array_equal(got_s[..., :6], want_s[..., :6]) and np.array_equal(got_s[..., 10], want_s[..., 10])):
                bad.append("summary integer columns")
            if not np.allclose(got_s[..., 6:10], want_s[..., 6:10], rtol=1e-6, atol=0):
                bad.append("summary size columns")
    return bad


def main():
    first = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    count = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    fails = 0
    for seed in range(first, first + count):
        try:
            bad = run(seed)
        except Exception as e:                               # noqa: BLE001
            bad = ["exception %r" % (e,)]
        if bad:
            fails += 1
            c = case(seed)
            print("seed", seed, {k: c[k] for k in ("B", "PH", "PW", "M", "C")}, bad, flush=True)
    print("fuzz_consumers: %d configurations, %d failures" % (count, fails))
    sys.exit(1 if fails else 0)


if __name__ == "__main__":
    main()
