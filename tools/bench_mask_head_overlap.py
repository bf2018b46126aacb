"""What the background fill (DetectionConfig(prefill=True)) is for: a pipeline WITH a mask head between the two halves.

The mask head (MaskSubNet, engine/layers/instance.py:158-240: conv stack on the 14x14 RoI features, 2x deconvolution,
1x1 sigmoid) is out of this repository's scope - it is dense tensor-core work that stays in cuDNN.  bench.py therefore
feeds synthetic mask probabilities and the two halves run back to back, where the fill has nothing to hide behind
(DESIGN.md 6g).  Here a STAND-IN head (torch convolutions, same shapes as the reference's default MaskSubNet: four 3x3
convs 128->128, a 2x2 stride-2 transposed conv, a 1x1 conv to C classes, channels-last, TF32) runs between
detect_and_align and trim_and_paste, as in a deployment, and the whole batch is timed with the fill off and on.
The head is library code and not part of any reported metric; only the DIFFERENCE between the two modes matters.

    python tools/bench_mask_head_overlap.py > profiles/mask_head_overlap_r02.txt
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import torch.nn as nn  # noqa: E402

import bench  # noqa: E402
import masklab_b200 as ml  # noqa: E402

torch.backends.cudnn.allow_tf32 = True
torch.backends.cuda.matmul.allow_tf32 = True
torch.backends.cudnn.benchmark = True


class StandInMaskHead(nn.Module):
    def __init__(self, cf, classes, depth=4):
        super().__init__()
        layers = []
        for _ in range(depth):
            layers += [nn.Conv2d(cf, cf, 3, padding=1), nn.ReLU(inplace=True)]
        layers += [nn.ConvTranspose2d(cf, cf, 2, stride=2), nn.ReLU(inplace=True), nn.Conv2d(cf, classes, 1), nn.Sigmoid()]
        self.net = nn.Sequential(*layers)

    def forward(self, x):                       # x [n,14,14,Cf] NHWC -> [n,28,28,C] NHWC
        y = self.net(x.permute(0, 3, 1, 2))     # channels-last memory, NCHW view
        return y.permute(0, 2, 3, 1).contiguous()


def main():
    wl = bench.WORKLOADS["cfg2"]
    B, C, Cf = wl["B"], wl["C"], wl["Cf"]
    cfgp, N, loc, cls, fmaps = bench.make_inputs(wl, B, seed=100)
    d = lambda a: torch.from_numpy(a).cuda()
    d_loc, d_cls, d_fmaps = d(loc), d(cls), [d(f) for f in fmaps]
    head = StandInMaskHead(Cf, C).cuda().eval().to(memory_format=torch.channels_last)
    print(f"# {torch.cuda.get_device_name(0)}, cfg-2 (B={B}), stand-in mask head between the halves; ms per batch, CUDA events")
    print("# mode                         eager (one stream)   one CUDA graph")
    for prefill in (False, True):
        cfg = ml.DetectionConfig(paste_output="uint8", prefill=prefill, **bench.kwargs_of(wl))
        pipe = ml.PostProcessPipeline(cfgp, (wl["H"], wl["W"]), (wl["PH"], wl["PW"]), C, Cf, B, cfg, private_context=True)
        rois = pipe.detect_and_align(d_loc, d_cls, d_fmaps)
        mf, R = rois.shapes()
        masks = torch.empty((B, R, 28, 28, C), device="cuda")

        def batch():
            r = pipe.detect_and_align(d_loc, d_cls, d_fmaps)
            with torch.no_grad():
                off = 0
                for f, m in enumerate(mf):          # one call per pyramid level, as PyramidRoiAlign hands them over
                    x = r.crops[f][:B * m * 14 * 14 * Cf].view(B * m, 14, 14, Cf)
                    masks[:, off:off + m] = head(x).view(B, m, 28, 28, C)
                    off += m
            pipe.trim_and_paste(r, masks)

        def timed(fn, n=50):
            for _ in range(5):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / n

        eager = timed(batch)
        pipe._own_context()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=torch.cuda.Stream(priority=-1)):
            batch()
        graph = timed(g.replay)
        print(f"background fill {'ON ' if prefill else 'OFF'}            {eager:8.3f}            {graph:8.3f}")
        del g, pipe
    # the head alone, for scale
    x = torch.randn(B * 118, 14, 14, Cf, device="cuda")
    with torch.no_grad():
        for _ in range(5):
            head(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            head(x)
        e1.record()
        torch.cuda.synchronize()
    print(f"# stand-in head alone on {B * 118} RoIs: {e0.elapsed_time(e1) / 50:.3f} ms")


if __name__ == "__main__":
    main()
