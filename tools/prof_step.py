"""A short single-stream run of the hot path for ncu (launch lists and --set full captures):

    python tools/prof_step.py [workload] [steps] [prefill 0|1] [planar 0|1]

One warm-up pass, then `steps` eager passes of detect_and_align + trim_and_paste on one stream."""
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
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
prefill = bool(int(sys.argv[3])) if len(sys.argv) > 3 else False
planar = bool(int(sys.argv[4])) if len(sys.argv) > 4 else False
wl = bench.WORKLOADS[name]
B, C = wl["B"], wl["C"]
cfgp, N, loc, cls, fmaps = bench.make_inputs(wl, B, seed=100)
cfg = ml.DetectionConfig(paste_output="uint8", prefill=prefill, mask_layout="planar" if planar else "interleaved",
                         **bench.kwargs_of(wl))
pipe = ml.PostProcessPipeline(cfgp, (wl["H"], wl["W"]), (wl["PH"], wl["PW"]), C, wl["Cf"], B, cfg, private_context=True)
d = lambda a: torch.from_numpy(a).cuda()
d_loc, d_cls, d_fmaps = d(loc), d(cls), [d(f) for f in fmaps]
rois = pipe.detect_and_align(d_loc, d_cls, d_fmaps)
_, R = rois.shapes()
probs = synth.mask_probs(B, R, C, seed=300)
d_masks = d(probs.transpose(0, 1, 4, 2, 3).copy() if planar else probs)
pipe.trim_and_paste(rois, d_masks)
torch.cuda.synchronize()
for _ in range(steps):
    r = pipe.detect_and_align(d_loc, d_cls, d_fmaps)
    pipe.trim_and_paste(r, d_masks)
torch.cuda.synchronize()
print("ok", name, "steps", steps, "prefill", prefill, "planar", planar, "M", int(pipe.trim_m.item()), "R", R)
