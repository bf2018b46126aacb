"""How far does the float64 contract of SummaryOutput's float columns (oracle/summary_oracle.py) sit
from what TensorFlow's float32 arithmetic would produce?  ("next round" item 1b.)

Two things in the reference are float32 where the contract is float64:
  * _calculate_theta (misc.py:706-718): float32 X^T X, float32 2x2 inverse, float32 products -> `unit` [PH];
  * tf.reduce_sum of float32 products (misc.py:633-658) -> instance / horizontal / vertical size.
This script emulates both in float32 (two accumulation orders each) on the synthetic road maps the tests and
bench.py use and prints the maximum relative deviations.  CPU only; run from the repo root:

    python tools/summary_f32_deviation.py > profiles/summary_f32_deviation_r02.txt
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth  # noqa: E402
from oracle import summary_oracle as so  # noqa: E402

F32, F64 = np.float32, np.float64


def rel(a, b):
    a, b = np.asarray(a, F64), np.asarray(b, F64)
    d = np.abs(a - b) / np.maximum(np.abs(b), 1e-30)
    return float(d.max()) if d.size else 0.0


def reductions_f32(unit, mask, order):
    """instance / horizontal / vertical size of ONE instance with float32 accumulation."""
    unit2 = (unit * unit).astype(F32)
    prod2 = (unit2[:, None] * mask).astype(F32)
    prod1 = (unit[:, None] * mask).astype(F32)
    vert_terms = (unit * (mask > 0.5).any(axis=1).astype(F32)).astype(F32)
    if order == "pairwise":                          # numpy's float32 sum is pairwise: a tree, like Eigen's
        inst = F32(prod2.sum(dtype=F32))
        horiz = F32(prod1.sum(axis=0, dtype=F32).max())
        vert = F32(vert_terms.sum(dtype=F32))
    else:                                            # one float32 add per term, row-major
        inst = F32(np.cumsum(prod2.reshape(-1), dtype=F32)[-1]) if prod2.size else F32(0)
        horiz = F32(np.cumsum(prod1, axis=0, dtype=F32)[-1].max())
        vert = F32(np.cumsum(vert_terms, dtype=F32)[-1])
    return inst, horiz, vert


def main():
    rng = np.random.default_rng(0)
    print("# float64 contract vs float32 emulation of TensorFlow's arithmetic (max relative deviation)")
    print("# frames: tests/synth.semantic_map (noisy trapezoid road, gaps, single-pixel rows); 16 frames per size")
    print("size        fit order    theta0      theta1      unit[PH]    | reductions order  instance    horizontal  vertical")
    for PH, PW in ((512, 1024), (1080, 1920), (540, 960)):
        seg = synth.semantic_map(16, PH, PW, seed=1000 + PH)
        worst = {}
        for order in ("blas", "sequential"):
            dt0 = dt1 = du = 0.0
            for b in range(seg.shape[0]):
                road = seg[b, :, :, 1]
                left, right = so.road_marginals(road)
                for pts in (left, right):
                    a = so.fit_line(pts)
                    e = so.fit_line_f32(pts, order)
                    dt0, dt1 = max(dt0, rel(e[0], a[0])), max(dt1, rel(e[1], a[1]))
                u64 = so.road_unit_lengths(road)
                u32 = so.road_unit_lengths(road, fit=lambda p: so.fit_line_f32(p, order))
                du = max(du, rel(u32, u64))
            worst[order] = (dt0, dt1, du)
        # reductions: contract (f64 accumulation of the f32 terms) vs f32 accumulation, same unit
        red = {}
        for order in ("pairwise", "sequential"):
            di = dh = dv = 0.0
            for b in range(4):
                unit = so.road_unit_lengths(seg[b, :, :, 1])
                for _ in range(6):
                    w, h = int(rng.integers(8, PW // 2)), int(rng.integers(8, PH // 2))
                    x0, y0 = int(rng.integers(0, PW - w)), int(rng.integers(0, PH - h))
                    m = np.zeros((PH, PW), F32)
                    m[y0:y0 + h, x0:x0 + w] = rng.random((h, w), dtype=F32)          # soft paste values
                    want = so.instance_reductions(seg[b:b + 1], m[None, None])[0, 0]
                    got = reductions_f32(unit, m, order)
                    di, dh, dv = max(di, rel(got[0], want[1])), max(dh, rel(got[1], want[2])), max(dv, rel(got[2], want[3]))
            red[order] = (di, dh, dv)
        for (fo, fv), (ro, rv) in zip(worst.items(), red.items()):
            print(f"{PW}x{PH:<5} {fo:<11} {fv[0]:.3e}   {fv[1]:.3e}   {fv[2]:.3e}   | {ro:<16}  {rv[0]:.3e}   {rv[1]:.3e}   {rv[2]:.3e}")


if __name__ == "__main__":
    main()
