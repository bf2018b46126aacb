"""bench.py's contract, the parts that can be checked without a GPU: the reference arm (`--impl reference`) runs the C
restatement on the host cores, honours --steps / --warmup, prints exactly one JSON line with the keys the driver reads,
and describes the same `config` the GPU arm would; the GPU arm's helpers (algorithmic bytes, source hash) are pure."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                       timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines                                  # ONE JSON line on stdout
    return json.loads(lines[0])


def test_reference_arm_prints_the_contract_line():
    line = _run("--impl", "reference", "--workload", "tiny", "--gpus", "1", "--steps", "3", "--warmup", "1")
    assert line["impl"] == "reference" and line["steps"] == 3 and line["warmup"] == 1
    assert line["metric"].startswith("frames/sec") and line["unit"] == "frames/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["ms_per_step"] > 0 and line["vs_baseline"] is None and line["data"] == "synthetic"
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "oracle/c" in cb["sample"]
    e2e = line["e2e"]
    assert e2e["value"] == line["value"] and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"] and line["config"]["workload"].startswith("tiny")


def test_non_zero_ranks_of_the_reference_arm_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "tiny",
                        "--gpus", "2", "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=300,
                       cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_algorithmic_bytes_and_source_hash():
    sys.path.insert(0, ROOT)
    import bench
    wl = bench.WORKLOADS["cfg2"]
    n = 163680
    got = bench.algorithmic_bytes(wl, n, 100)
    # 2.38 GB per cfg-2 batch (decode + det + FPN maps + crops + trim + uint8 paste)
    assert abs(got / 1e9 - 2.374) < 0.01
    assert set(bench.WORKLOADS) >= {"cfg1", "cfg2", "stress", "cfg5"}
    h = bench.source_hash()
    assert len(h) == 16 and h == bench.source_hash()
