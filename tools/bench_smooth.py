"""Times SemanticSmoothing (mlp_semantic_smoothing) on the serving shapes: python tools/bench_smooth.py [k]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import masklab_b200 as ml  # noqa: E402


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    k = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    g = torch.Generator(device="cuda").manual_seed(0)
    for S in (1, 3):
        x = torch.rand((32, 540, 960, S), device="cuda", generator=g)
        layer = ml.SemanticSmoothing(kernel_size=k, weight=1.1)
        ms = timed(lambda: layer(x))
        gb = x.numel() * 4 * 2 / 1e9
        print(f"SemanticSmoothing k={k} [32,540,960,{S}]: {ms:.3f} ms, {gb / ms * 1e3:.0f} GB/s of map in + out "
              f"({'passes' if os.environ.get('MLP_SMOOTH_PASSES') else 'fused'})")


if __name__ == "__main__":
    main()
