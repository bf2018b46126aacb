"""Phase timeline of the two NMS kernels (CTA 0) from a -DMLP_NMS_TIMING build (tuning aid):

    python tools/nms_timing.py [workload]
"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import masklab_b200.build as b  # noqa: E402

lib_path = os.path.join(b.HERE, "libmasklab_b200_nmst.so")
b.build_library(force=True, out=lib_path, extra=["-DMLP_NMS_TIMING"])
os.environ["MASKLAB_B200_LIB"] = lib_path
import torch  # noqa: E402

import bench  # noqa: E402
import masklab_b200 as ml  # noqa: E402
from masklab_b200 import runtime as rt  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
wl = bench.WORKLOADS[name]
B, C = wl["B"], wl["C"]
cfgp, N, loc, cls, fmaps = bench.make_inputs(wl, B, seed=100)
pipe = ml.PostProcessPipeline(cfgp, (wl["H"], wl["W"]), (wl["PH"], wl["PW"]), C, wl["Cf"], B,
                              ml.DetectionConfig(**bench.kwargs_of(wl)), private_context=True)
d = lambda a: torch.from_numpy(a).cuda()
ins = (d(loc), d(cls), [d(f) for f in fmaps])
for _ in range(3):
    pipe.detect_and_align(*ins)
torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * 64)()
lib = rt.load_library()
lib.mlp_debug_nms_timing.argtypes = [ctypes.c_void_p]
assert lib.mlp_debug_nms_timing(buf) == 0
for k, kname in enumerate(("nms_per_class (CTA 0)", "nms_cross_class (CTA 0)")):
    t = list(buf[k * 32:(k + 1) * 32])
    t0 = t[0]
    marks = [(i, v) for i, v in enumerate(t) if v]
    print(kname)
    prev = t0
    for i, v in sorted(marks, key=lambda m: m[1]):
        print(f"   mark {i:2d}  +{(v - prev) / 1e3:7.2f} us   (at {(v - t0) / 1e3:7.2f} us)")
        prev = v
