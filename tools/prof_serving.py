"""A short single-stream run of the SERVING TAIL for ncu: detect_and_align -> trim_and_summarize -> draw -> encode.

    python tools/prof_serving.py [workload] [steps]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import bench  # noqa: E402
import synth  # noqa: E402
import masklab_b200 as ml  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
wl = bench.WORKLOADS[name]
B, C, PH, PW = wl["B"], wl["C"], wl["PH"], wl["PW"]
cfgp, N, loc, cls, fmaps = bench.make_inputs(wl, B, seed=100)
cfg = ml.DetectionConfig(paste_output="uint8", **bench.kwargs_of(wl))
pipe = ml.PostProcessPipeline(cfgp, (wl["H"], wl["W"]), (PH, PW), C, wl["Cf"], B, cfg, private_context=True)
d = lambda a: torch.from_numpy(a).cuda()
d_loc, d_cls, d_fmaps = d(loc), d(cls), [d(f) for f in fmaps]
d_seg = d(synth.semantic_map(B, PH, PW, seed=500))
d_img = d(synth.road_frames(B, PH, PW, seed=600))
rois = pipe.detect_and_align(d_loc, d_cls, d_fmaps)
_, R = rois.shapes()
d_masks = d(synth.mask_probs(B, R, C, seed=300))


def serve():
    r = pipe.detect_and_align(d_loc, d_cls, d_fmaps, prefill=False)
    pipe.trim_and_summarize(r, d_masks, d_seg)
    pipe.draw(r, d_masks, d_img, bench.INST_COLORS[:C], 0.3, seg_outs=d_seg, semantic_colors=bench.SEM_COLORS,
              semantic_alpha=0.3, boxes=True)
    pipe.encode()


serve()
torch.cuda.synchronize()
for _ in range(steps):
    serve()
torch.cuda.synchronize()
print("ok", name, "steps", steps, "launches", pipe.ctx.launch_count())
