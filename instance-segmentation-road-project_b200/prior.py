"""PriorBoxes — mirror of /root/reference/engine/prior.py:9-71 (same constructor, `config`,
`get_config()`, `len()` and `.boxes` table), plus the packing of that table into the C
struct the kernels take.  Host-side, runs once at graph-construction time.
"""
import numpy as np

from . import runtime as rt


class PriorBoxes:
    """Default-box (anchor) configuration.  `boxes` rows are (stride, w, h) in the
    reference's order: for (size, stride) / for scale / for ratio (prior.py:55-67)."""

    def __init__(self, strides, sizes, pr_scales, pr_ratios):
        def tolist(v):
            return v.tolist() if isinstance(v, np.ndarray) else list(v)
        self.strides = tolist(strides)
        self.sizes = tolist(sizes)
        self.pr_scales = tolist(pr_scales)
        self.pr_ratios = tolist(pr_ratios)
        self.setup()
        assert len(self.strides) == len(self.sizes), \
            "the number of strides and of sizes must be equal"          # prior.py:40
        self.config = {"strides": self.strides, "sizes": self.sizes,
                       "pr_scales": self.pr_scales, "pr_ratios": self.pr_ratios}

    def __len__(self):
        return len(self.pr_scales) * len(self.pr_ratios)                 # prior.py:48-53

    def setup(self):
        rows = []
        for size, stride in zip(self.sizes, self.strides):
            for wh_size in self.pr_scales:
                for wh_ratio in self.pr_ratios:
                    # float64, round-half-even, then int — exactly prior.py:60-61
                    w = int(np.round(size * wh_size * np.sqrt(wh_ratio)))
                    h = int(np.round(size * wh_size / np.sqrt(wh_ratio)))
                    rows.append((int(stride), w, h))
        self.table = np.asarray(rows, dtype=np.int64).reshape(-1, 3)

    @property
    def boxes(self):
        """pandas DataFrame[stride,w,h] indexed from 1, like the reference's attribute."""
        import pandas as pd
        df = pd.DataFrame(self.table, columns=["stride", "w", "h"])
        df.index = df.index + 1
        return df

    def get_config(self):
        return self.config

    # ---- packing for the C ABI (grouped by stride ascending like
    #      PriorLayer.__init__, engine/layers/detection.py:260-262) ----
    def grouped(self):
        out = []
        for stride in sorted(set(self.table[:, 0].tolist())):
            rows = self.table[self.table[:, 0] == stride]
            out.append((int(stride), [(int(w), int(h)) for _, w, h in rows]))
        return out

    def to_c(self, padding="same"):
        groups = self.grouped()
        if len(groups) > rt.MLP_MAX_LEVELS:
            raise rt.InvalidArgumentError(rt.MLP_EINVAL, f"more than {rt.MLP_MAX_LEVELS} strides")
        c = rt.PriorConfigC()
        c.num_levels = len(groups)
        c.padding_same = 1 if padding == "same" else 0
        for l, (stride, whs) in enumerate(groups):
            if len(whs) > rt.MLP_MAX_ANCHORS:
                raise rt.InvalidArgumentError(
                    rt.MLP_EINVAL, f"more than {rt.MLP_MAX_ANCHORS} anchors at stride {stride}")
            c.stride[l] = stride
            c.num_anchors[l] = len(whs)
            for a, (w, h) in enumerate(whs):
                c.anchor_w[l][a] = w
                c.anchor_h[l][a] = h
        return c
