"""Times EncodeImageContent.encode_batch (mlp_jpeg_encode) on overlay-like frames: CUDA events around `iters` calls.
usage: python tools/bench_jpeg.py [B H W iters]"""
import sys

import numpy as np
import torch

import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import masklab_b200 as ml  # noqa: E402


def frames(B, H, W, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    yy = torch.arange(H, device="cuda", dtype=torch.float32)[:, None]
    xx = torch.arange(W, device="cuda", dtype=torch.float32)[None, :]
    out = []
    for b in range(B):
        img = torch.stack([120 + 90 * torch.sin(xx / (23.0 + b) + yy / 31.0), 110 + 80 * torch.cos(xx / 57.0 - yy / (19.0 + b)),
                           90 + yy * (120.0 / H) + 30 * torch.sin(xx / 7.0)], -1)
        img = img + torch.randn(img.shape, device="cuda", generator=g) * 6
        out.append(img.clamp(0, 255).to(torch.uint8))
    return torch.stack(out)


def main():
    B, H, W, iters = (int(a) for a in sys.argv[1:5]) if len(sys.argv) >= 5 else (32, 512, 1024, 20)
    fr = frames(B, H, W)
    enc = ml.EncodeImageContent()
    out, lengths = enc.encode_batch(fr)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        enc.encode_batch(fr, out=out, lengths=lengths)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        enc.encode_batch(fr, out=out, lengths=lengths)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    n = lengths.cpu().numpy()
    print(f"jpeg B={B} {H}x{W}: {ms:.3f} ms per batch = {B / ms * 1e3:.0f} frames/s, "
          f"{B * H * W * 3 / ms / 1e6:.1f} GB/s of pixels, mean file {n.mean() / 1024:.1f} KiB ({n.mean() / (H * W):.3f} B/px)")


if __name__ == "__main__":
    main()
