"""Paste-kernel-only timing (cfg-2 shape by default) for tuning; prints GB/s of the uint8 output."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import synth
import masklab_b200 as ml

B, M, PH, PW = int(os.environ.get("B", 32)), int(os.environ.get("M", 100)), 512, 1024
det = synth.detections(B, M, 5, PH, PW, seed=1)
masks = synth.mask_probs(B, M, 1, seed=2)[..., 0]
det_i, mask_i, _ = ml.UpSampleOutput(semantic=False)([torch.from_numpy(det).cuda(), torch.from_numpy(masks).cuda(),
                                                      (PH, PW)], target=(PH, PW))
det_i, mask_i = det_i.cpu().numpy(), mask_i.cpu().numpy()
if os.environ.get("ZERO"):
    det_i[..., 0] = PW + 1000      # every box off-frame: the kernel degenerates to a pure zero fill
if os.environ.get("SMALL"):
    det_i[..., 2:4] = np.minimum(det_i[..., 2:4], int(os.environ["SMALL"]))
d, m = torch.from_numpy(det_i).cuda(), torch.from_numpy(mask_i).cuda()
layer = ml.CropAndPadMask(output=os.environ.get("OUT", "uint8"))
ctx = ml.Context.get(0)
for _ in range(3):
    out = layer([(PH, PW), d, m])
torch.cuda.synchronize()
ctx.profile(True)
for _ in range(20):
    out = layer([(PH, PW), d, m])
st = ctx.profile_read()
ms = st["paste"][0] / st["paste"][1]
nbytes = out.numel() * out.element_size()
print(f"ctas/sm={os.environ.get('MLP_PASTE_CTAS_PER_SM','occ')} band_kb={os.environ.get('MLP_PASTE_BAND_KB','64')} "
      f"paste {ms*1e3:.1f} us  {nbytes/ms/1e6:.0f} GB/s  (thr kernel {st['paste_threshold'][0]/st['paste_threshold'][1]*1e3:.1f} us)")
