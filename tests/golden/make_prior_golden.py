"""Generate tests/golden/prior_tables.json by importing the reference's own
engine/prior.py (the only hot-path file of the reference that runs without
TensorFlow).  Run in the authoring container only:

    python tests/golden/make_prior_golden.py

/root/reference is read-only and absent on the GPU box; the JSON this script
writes is what travels.  NumPy >= 1.24 removed `np.int`, which
engine/prior.py:60-66 uses, so the harness (not the reference) aliases it.
"""
import importlib.util
import json
import os
import sys

import numpy as np

REF = "/root/reference/engine/prior.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "prior_tables.json")


def load_reference_prior():
    if not hasattr(np, "int"):
        np.int = int  # harness-side shim, see module docstring
    spec = importlib.util.spec_from_file_location("_ref_prior", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.PriorBoxes


def table(PriorBoxes, strides, scales, ratios):
    sizes = [4 * s for s in strides]          # engine/retinamasklab.py:46-49
    pb = PriorBoxes(strides=strides, sizes=sizes, pr_scales=scales, pr_ratios=ratios)
    rows = [[int(r.stride), int(r.w), int(r.h)] for r in pb.boxes.itertuples()]
    return {"strides": strides, "sizes": sizes, "pr_scales": scales, "pr_ratios": ratios,
            "num_anchors": len(pb), "rows": rows}


def main():
    PriorBoxes = load_reference_prior()
    cases = {
        # engine/config.py:60-61 + backbone_outputs C3..P7 (config.py:53)
        "default": table(PriorBoxes, [8, 16, 32, 64, 128],
                         [2 ** 0, 2 ** (1 / 3), 2 ** (2 / 3)], [1 / 3, 1 / 2, 1, 2, 3]),
        # road_project/train.py:37,44-45
        "road_project": table(PriorBoxes, [8, 16, 32, 64],
                              [2 ** 0, 2 ** (1 / 3), 2 ** (2 / 3)], [1 / 2, 1, 2, 5, 8]),
        # BASELINE.json configs[2]: ~100k anchors (A=9)
        "stress_a9": table(PriorBoxes, [8, 16, 32, 64, 128],
                           [2 ** 0, 2 ** (1 / 3), 2 ** (2 / 3)], [1 / 2, 1, 2]),
    }
    with open(OUT, "w") as f:
        json.dump(cases, f, indent=1)
    print("wrote", OUT, {k: len(v["rows"]) for k, v in cases.items()})


if __name__ == "__main__":
    sys.exit(main())
