"""Reads one bench.py JSON line from stdin and prints a one-line stage summary (tuning helper)."""
import json
import sys

d = json.loads(sys.stdin.read())
s = d["roofline"]["timed_region"]["stage_bracket_ms_per_step"]
tag = sys.argv[1] if len(sys.argv) > 1 else ""
print(tag, "step_us", round(d["ms_per_step"] * 1e3, 1), "fps", round(d["value"]),
      "frac", round(d["roofline"]["whole_step"]["frac"], 3),
      " ".join(f"{k}={v * 1e3:.1f}" for k, v in s.items()))
