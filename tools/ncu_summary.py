"""Summarise an .ncu-rep (raw page) and a launch-list CSV into the markdown kept under profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep gpurun_out/launches.csv > profiles/ncu_summary_rNN.md
"""
import csv
import io
import json
import subprocess
import sys
from collections import defaultdict

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__grid_size", "grid"),
    ("smsp__inst_executed.sum", "warp instr"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio_throttle"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
]


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main():
    """argv: report.ncu-rep launches.csv [traffic.json workload]; with the last two the DRAM bytes per launch are also
    written as JSON together with the hash of the CUDA sources they were captured from (bench.py uses the file only
    when that hash matches the sources it runs)."""
    rep, launches = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    print(f"# ncu summary of `{rep}` (ncu --set full --clock-control none, one launch per kernel)\n")
    traffic = {}
    for r in rows[2:]:
        name = r[ki].split("(")[0].replace("<unnamed>::", "").replace("void ", "")
        print(f"## {name}\n")
        print("| metric | value |\n|---|---|")
        rd = wr = None
        for m, label in METRICS:
            if m in hdr:
                i = hdr.index(m)
                print(f"| {label} | {r[i]} {units[i]} |")
                if m == "dram__bytes_read.sum":
                    rd = to_bytes(r[i], units[i])
                if m == "dram__bytes_write.sum":
                    wr = to_bytes(r[i], units[i])
        if rd is not None and wr is not None:
            traffic[name] = rd + wr
            print(f"| dram traffic (read+write) | {(rd + wr) / 1e6:.1f} MB |")
        print()
    rows = [r for r in csv.reader(open(launches)) if len(r) > 5]
    h = rows[0]
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    d = defaultdict(list)
    for r in rows[1:]:
        d[r[ki].split("(")[0].replace("<unnamed>::", "").replace("void ", "")].append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in d.values())
    print(f"# launch list `{launches}` (gpu__time_duration.sum; cold-cache, serialised: compare shares)\n")
    print("| kernel | launches | avg us | share |\n|---|---|---|---|")
    for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
        print(f"| {k} | {len(v)} | {sum(v) / len(v) / 1000:.2f} | {sum(v) / tot * 100:.1f}% |")
    print("\n<!-- traffic-json: " + json.dumps(traffic) + " -->")
    if len(sys.argv) > 4:
        import os
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        import bench
        with open(sys.argv[3], "w") as f:
            json.dump({"source": f"{os.path.basename(rep)} (ncu --set full --clock-control none, single stream)",
                       "workload": sys.argv[4], "source_hash": bench.source_hash(),
                       "dram_bytes_per_launch": traffic}, f, indent=1)


if __name__ == "__main__":
    main()
