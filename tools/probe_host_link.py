"""Host-link ceiling of the box: concurrent pinned host<->device copies on 1, 2, 4, ... GPUs at once.

    python tools/probe_host_link.py [MiB per copy, default 1024] > profiles/host_link_r02.txt

For every GPU count n (powers of two up to the visible GPUs) and every direction mix (D2H only, H2D only, both
at once) one large `cudaMemcpyAsync` per direction per GPU is enqueued on its own stream from/to its own pinned
host buffer, all GPUs at the same time, five times over; the table reports per-GPU and aggregate GB/s of the
best round (wall clock around a device-wide synchronise; plain copies, no kernels).  VERDICT r1 item 7: bench.py's
end-to-end leg moves 1.68 GB device->host and 0.42 GB host->device per step and rank - if the aggregate stops
growing with n, the e2e curve is the host's, not the path's.
"""
import sys
import time

import torch


def run(n, mib, d2h, h2d, rounds=5):
    nbytes = mib << 20
    devs = list(range(n))
    bufs = []
    for d in devs:
        torch.cuda.set_device(d)
        item = dict(dev=d)
        if d2h:
            item["d_out"] = torch.empty(nbytes, dtype=torch.uint8, device=f"cuda:{d}")
            item["h_out"] = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
            item["s_out"] = torch.cuda.Stream(device=d)
        if h2d:
            item["d_in"] = torch.empty(nbytes, dtype=torch.uint8, device=f"cuda:{d}")
            item["h_in"] = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
            item["s_in"] = torch.cuda.Stream(device=d)
        bufs.append(item)

    def sync():
        for d in devs:
            torch.cuda.synchronize(d)

    best = None
    for it in range(rounds + 1):
        sync()
        t0 = time.perf_counter()
        for b in bufs:
            if d2h:
                with torch.cuda.stream(b["s_out"]):
                    b["h_out"].copy_(b["d_out"], non_blocking=True)
            if h2d:
                with torch.cuda.stream(b["s_in"]):
                    b["d_in"].copy_(b["h_in"], non_blocking=True)
        sync()
        dt = time.perf_counter() - t0
        if it > 0:                                  # round 0 = warm-up
            best = dt if best is None else min(best, dt)
    per_dir = nbytes / best / 1e9
    return per_dir, per_dir * n * (int(d2h) + int(h2d))


def main():
    mib = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    total = torch.cuda.device_count()
    print(f"# {total} x {torch.cuda.get_device_name(0)}, {mib} MiB per copy, pinned host memory, best of 5 rounds")
    print("gpus  mix        GB/s per GPU and direction   aggregate GB/s")
    n = 1
    while n <= total:
        for name, d2h, h2d in (("D2H", True, False), ("H2D", False, True), ("D2H+H2D", True, True)):
            per, agg = run(n, mib, d2h, h2d)
            print(f"{n:<5} {name:<10} {per:10.1f}                    {agg:10.1f}", flush=True)
        n *= 2


if __name__ == "__main__":
    main()
