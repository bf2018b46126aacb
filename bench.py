#!/usr/bin/env python
"""bench.py — frames/s of the MaskLab post-backbone path (decode -> NMS -> RoIAlign -> trim ->
mask paste) on 1..8 B200, with the roofline of the dominant kernel and a CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|stress|cfg5]

One "step" = one pass of the whole hot path over one synthetic batch of 32 frames per GPU
(BASELINE.json configs[1]: 1024x512 frames, N=163,680 anchors, C=5, FPN Cf=128).  Frames are
sharded by image: every rank runs its own batch on its own stream (weak scaling), the only
collective is one NCCL all_gather of the fixed-capacity detection records at the end of the
timed region.  Rank 0 prints ONE JSON line.
"""
import argparse
import contextlib
import io
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

WORKLOADS = {
    # BASELINE.json configs[0] / SURVEY §8 cfg-1: the serving graph's own shape - ONE 1920x1080 frame, the model at
    # DownSampleInput resolution 540x960 (N=163,275), masks pasted at 1080x1920 (retinamasklab.py:601,
    # road_project/setup/serving.py, engine/config.py:19); 5 instance classes (README: car/bump/manhole/steel/pothole)
    "cfg1": dict(B=1, H=540, W=960, PH=1080, PW=1920, C=5, Cf=128, ratios="default", mu=-5.8,
                 min_confidence=0.05, nms_iou_threshold=0.4, post_iou_threshold=0.6,
                 nms_max_output_size=100, max_k=2, base_size=36),
    # BASELINE.json configs[1] / SURVEY §8 cfg-2: ResNeXt default ModelConfiguration shapes
    # (engine/config.py:53,60-64,83-86,150) with the north-star score threshold 0.05.
    "cfg2": dict(B=32, H=512, W=1024, PH=512, PW=1024, C=5, Cf=128, ratios="default", mu=-5.8,
                 min_confidence=0.05, nms_iou_threshold=0.4, post_iou_threshold=0.6,
                 nms_max_output_size=100, max_k=2, base_size=36),
    # configs[2]+[3]: NMS/top-k stress + 1000 RoIs/image mask paste
    "stress": dict(B=32, H=512, W=1024, PH=512, PW=1024, C=6, Cf=128, ratios="default", mu=-5.0,
                   min_confidence=0.05, nms_iou_threshold=0.4, post_iou_threshold=0.65,
                   nms_max_output_size=1000, max_k=2, base_size=64),
    # configs[4]: 1080p frames, model at 540x960, paste at 1080x1920
    "cfg5": dict(B=32, H=540, W=960, PH=1080, PW=1920, C=6, Cf=128, ratios="default", mu=-5.8,
                 min_confidence=0.05, nms_iou_threshold=0.4, post_iou_threshold=0.65,
                 nms_max_output_size=100, max_k=2, base_size=36),
    # small variant for quick checks
    "tiny": dict(B=4, H=128, W=256, PH=128, PW=256, C=5, Cf=32, ratios="default", mu=-5.0,
                 min_confidence=0.05, nms_iou_threshold=0.4, post_iou_threshold=0.6,
                 nms_max_output_size=50, max_k=2, base_size=36),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--e2e-steps", type=int, default=8)
    ap.add_argument("--streams", type=int, default=6,
                    help="independent pipelines/streams per GPU; consecutive batches alternate between them")
    ap.add_argument("--input-sets", type=int, default=3,
                    help="distinct input sets per stream, rotated step by step: no tensor is read by two batches in "
                         "flight and none is re-read before everything else has passed through L2")
    ap.add_argument("--prefill", type=int, default=0,
                    help="1: CropAndPadMask's zero background is streamed out on a second stream from the moment "
                         "NMS has produced M (mlp_paste_prefill), the paste kernel writes the boxes only")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="skip the CUDA-graph replay leg")
    ap.add_argument("--no-graph-main", dest="graph_main", action="store_false",
                    help="eager launches in the timed region instead of one captured CUDA graph per stream")
    ap.add_argument("--no-summary", action="store_true",
                    help="skip the SummaryOutput legs (SURVEY 8(f) rank 1: the consumer of the masks)")
    ap.add_argument("--cpu-frames", type=int, default=0, help="frames in the CPU baseline sample (0 = auto)")
    ap.add_argument("--frames", type=int, default=0,
                    help="streaming job (BASELINE configs[4]): process this many frames in total, sharded by image "
                         "over the ranks in batches of B, detections gathered at the end; overrides --steps "
                         "(strong scaling)")
    return ap.parse_args()


def kwargs_of(wl):
    return {k: wl[k] for k in ("min_confidence", "nms_iou_threshold", "post_iou_threshold",
                               "nms_max_output_size", "max_k", "base_size")}


def make_inputs(wl, B, seed):
    import synth
    cfgp = synth.prior_config()
    N = synth.num_anchors(cfgp, wl["H"], wl["W"])
    loc, cls = synth.head_tensors(B, N, wl["C"], mu=wl["mu"], seed=seed)
    fmaps = synth.fpn_maps(B, wl["H"], wl["W"], wl["Cf"], seed=seed + 1000)
    return cfgp, N, loc, cls, fmaps


# ------------------------------------------------------------------ clocks -----
class ClockSampler:
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:          # pragma: no cover
            self.nv = None
            self.err = repr(e)

    def _run(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.0005)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thread:
            self._thread.join()

    def summary(self):
        if self.nv is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "note": "NVML unavailable"}
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# --------------------------------------------------------------- CPU baseline --
class CpuSample:
    """A bounded sample of the workload for the CPU arm: `frames` single-frame inputs generated up
    front (outside the timed region), run through the C restatement of the reference path
    (oracle/c/masklab_oracle.c: "port" - restated CPU path, TensorFlow is not installable here).
    Threads parallelise over frames; ctypes releases the GIL during the C calls."""

    def __init__(self, wl, frames, bits=False):
        import synth
        from oracle import c_oracle as co
        self.wl, self.co, self.bits = wl, co, bits
        co.lib()
        self.inputs = [make_inputs(wl, 1, 7000 + i) for i in range(frames)]
        rcap = (wl["max_k"] + 1) * wl["nms_max_output_size"]
        self.mask_pool = synth.mask_probs(1, rcap, wl["C"], seed=99)      # mask-head stand-in

    def one(self, i):
        wl = self.wl
        cfgp, N, loc, cls, fmaps = self.inputs[i]
        out = self.co.full_path(loc, cls, fmaps, lambda f, b: self.mask_pool[:, :b.shape[1]], cfgp,
                                (wl["H"], wl["W"]), (wl["PH"], wl["PW"]), binary=True, bits=self.bits,
                                **kwargs_of(wl))
        return int(out["bits" if self.bits else "binary"].shape[1])

    def run(self, threads, repeat=1):
        n = len(self.inputs)
        idx = [i % n for i in range(n * repeat)]
        t0 = time.perf_counter()
        if threads <= 1:
            for i in idx:
                self.one(i)
        else:
            from concurrent.futures import ThreadPoolExecutor
            with ThreadPoolExecutor(threads) as ex:
                list(ex.map(self.one, idx))
        wall = time.perf_counter() - t0
        return len(idx) / wall, wall


def run_reference(args, wl):
    """--impl reference: the reference's own implementation is Python over TensorFlow 1.x CPU
    kernels, which cannot be installed here; this arm times the C restatement of
    that path on all host cores.  W warm-up steps, then exactly K timed steps; a step is a bounded
    sample of the workload (one batch; fewer frames when K steps of a whole batch would not
    end within a few minutes)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    threads = max(1, min(cores, 64))
    frames_per_step = max(threads, wl["B"])
    sample = CpuSample(wl, frames_per_step)
    _, w1 = sample.run(threads)                         # page faults, thread pool (untimed, before the W warm-ups)
    budget_s = 150.0
    if w1 * (args.steps + args.warmup) > budget_s:      # bound the whole run: fewer frames per step
        frames_per_step = max(1, int(frames_per_step * budget_s / (w1 * (args.steps + args.warmup))))
        sample.inputs = sample.inputs[:frames_per_step]
    for _ in range(args.warmup):
        sample.run(threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        sample.run(threads)
    wall = time.perf_counter() - t0
    steps = args.steps
    fps = steps * frames_per_step / wall
    fps_bits = None
    if wl["PW"] % 8 == 0:
        sample_b = CpuSample(wl, frames_per_step, bits=True)
        sample_b.run(threads)
        fps_bits = sample_b.run(threads, repeat=max(1, min(steps, 5)))[0]
    desc = (f"{steps} step(s) x {frames_per_step} frames of workload {args.workload} through "
            f"oracle/c (uint8 paste), {threads} threads")
    line = {
        "impl": "reference", "metric": "frames/sec decode+NMS+RoIAlign+mask-paste", "value": fps,
        "unit": "frames/s", "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * wall / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, wl),
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port", "sample": desc},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "e2e_bitpacked": {"value": fps_bits, "unit": "frames/s", "note": "same arm writing 1 bit per pixel"},
        "note": "C restatement of the reference path (TF kernels restated); TensorFlow 1.x is absent",
    }
    print(json.dumps(line), flush=True)


def workload_config(args, wl):
    return {"workload": f"{args.workload}: B={wl['B']}/GPU {wl['W']}x{wl['H']} frames, C={wl['C']}, "
                        f"Cf={wl['Cf']}, max_out={wl['nms_max_output_size']}, paste {wl['PW']}x{wl['PH']} uint8",
            "batch_per_gpu": wl["B"], "parallelism": f"image-sharded x{args.gpus}",
            "stream_frames": (args.frames or None), "cuda_graphs": bool(getattr(args, "graph_main", False)) and not args.frames,
            "streams": args.streams, "input_sets_per_stream": args.input_sets, "prefill": bool(args.prefill),
            "l2_policy": f"{args.input_sets} distinct input sets per stream rotated step by step (no tensor is read by "
                         "two batches in flight); one set (>= 365 MB at cfg2) and one batch of outputs (>= 1.6 GB) each "
                         "exceed the 126 MB L2",
            "score_mu": wl["mu"]}


def source_hash():
    """sha256 over the CUDA sources + the header: ties a committed ncu capture to the build that produced it."""
    import hashlib
    import masklab_b200.build as b
    h = hashlib.sha256()
    for f in sorted(b.sources() + [os.path.join(b.CSRC, "common.cuh"), os.path.join(b.CSRC, "paste_common.cuh"),
                                   os.path.join(b.INCLUDE, "masklab_b200.h")]):
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


# -------------------------------------------------------------------- ours -----
def run_ours(args, wl):
    import torch
    import torch.distributed as dist
    import synth
    import masklab_b200 as ml
    from masklab_b200 import dist as mdist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    B, C = wl["B"], wl["C"]
    H, W, PH, PW = wl["H"], wl["W"], wl["PH"], wl["PW"]
    if args.frames:
        # a stream of `frames` frames sharded by image: every rank owns frames/world of them and runs
        # them in batches of B (masklab_b200.dist.shard_frames / chunks); one gather at the end
        shard = mdist.shard_frames(args.frames, world, rank)
        args.steps = max(1, len(mdist.chunks(shard.padded, B)))
    cfgp, N, loc, cls, fmaps = make_inputs(wl, B, seed=100 + rank)
    cfg = ml.DetectionConfig(paste_output="uint8", prefill=bool(args.prefill), **kwargs_of(wl))
    pipe = ml.PostProcessPipeline(cfgp, (H, W), (PH, PW), C, wl["Cf"], B, cfg, device=local, private_context=True)
    K = pipe.K

    # host (pinned) copies for the end-to-end leg, device copies for the kernel-only leg
    pin = lambda a: torch.from_numpy(a).pin_memory()
    h_loc, h_cls, h_fmaps = pin(loc), pin(cls), [pin(f) for f in fmaps]
    d_loc, d_cls = h_loc.cuda(non_blocking=True), h_cls.cuda(non_blocking=True)
    d_fmaps = [f.cuda(non_blocking=True) for f in h_fmaps]

    # first pass: discover R (rows the mask head would produce) and make its synthetic output
    l0 = pipe.ctx.launch_count()
    rois = pipe.detect_and_align(d_loc, d_cls, d_fmaps)
    mf, R = rois.shapes()
    h_masks = pin(synth.mask_probs(B, R, C, seed=300 + rank))
    d_masks = h_masks.cuda()
    pipe.trim_and_paste(rois, d_masks)
    launches_per_step = pipe.ctx.launch_count() - l0                 # this library's kernels in one pass of the path
    M = int(pipe.trim_m.item())
    counts = rois.counts.cpu().numpy()
    torch.cuda.synchronize()

    # independent per-GPU streams (north star): batch i runs on pipeline/stream i % S, so the latency-bound NMS
    # kernels of one batch overlap the bandwidth-bound RoIAlign/paste of another.  S pipelines and S*NSETS input
    # sets must fit: both are trimmed to 80 % of the free memory (the stress workload holds 27 GB per pipeline).
    S = max(1, args.streams)
    NSETS = max(1, args.input_sets)
    set_bytes = sum(t.numel() * t.element_size() for t in [d_loc, d_cls, d_masks] + d_fmaps)
    free_b = torch.cuda.mem_get_info(local)[0]
    pipe_bytes = pipe.device_bytes()
    while S > 1 and (S - 1) * pipe_bytes + S * NSETS * set_bytes > 0.8 * free_b:
        if NSETS > 2:
            NSETS -= 1
        else:
            S -= 1
    pipes = [pipe] + [ml.PostProcessPipeline(cfgp, (H, W), (PH, PW), C, wl["Cf"], B, cfg, device=local,
                                             private_context=True) for _ in range(S - 1)]
    streams = [torch.cuda.Stream(priority=-1) for _ in range(S)]     # above the pipelines' background-fill streams
    # input set q = set 0 rolled by q frames along the batch axis: its own buffers (what L2 sees), the same
    # statistics, detections that are the rolled detections of set 0 (same M, R)
    def rolled(q):
        if q == 0:
            return dict(loc=d_loc, cls=d_cls, fmaps=d_fmaps, masks=d_masks)
        r = lambda t: torch.roll(t, q % max(B, 1), dims=0) if B > 1 else t.clone()
        return dict(loc=r(d_loc), cls=r(d_cls), fmaps=[r(f) for f in d_fmaps], masks=r(d_masks))
    sets = [[rolled(k * NSETS + q) for q in range(NSETS)] for k in range(S)]

    state = {"i": 0}
    # streaming job: the detection records of every batch of the shard are kept for the final gather
    words = pipe.record.numel()
    stream_rec = torch.empty((args.steps, words), dtype=torch.int32, device="cuda") if args.frames else None

    graphs = None
    if args.graph_main and not args.frames:
        # the path is sync-free, so a whole batch replays as ONE graph launch per (stream, input set)
        try:
            graphs = [[pipes[k].capture(z["loc"], z["cls"], z["fmaps"], z["masks"])[0] for z in sets[k]]
                      for k in range(S)]
        except Exception as exc:                     # pragma: no cover - fall back to eager launches
            print(f"CUDA graph capture failed, eager launches instead: {exc!r}", file=sys.stderr)
            graphs = None

    def step():
        i = state["i"]
        k = i % S
        q = (i // S) % NSETS
        state["i"] += 1
        with torch.cuda.stream(streams[k]):
            if graphs is not None:
                graphs[k][q].replay()
                return
            z = sets[k][q]
            r = pipes[k].detect_and_align(z["loc"], z["cls"], z["fmaps"])
            pipes[k].trim_and_paste(r, z["masks"])
            if stream_rec is not None:
                stream_rec[i % args.steps].copy_(pipes[k].record, non_blocking=True)

    def join_streams():
        for s_ in streams:
            torch.cuda.current_stream().wait_stream(s_)

    def fork_streams():
        for s_ in streams:
            s_.wait_stream(torch.cuda.current_stream())

    # the one collective of the path: ONE all_gather_into_tensor of the packed records (det + counts in one buffer
    # that the cross-class NMS kernel wrote in place)
    g_src = stream_rec.view(-1) if args.frames else pipe.record
    gathered = torch.empty((world, g_src.numel()), dtype=torch.int32, device="cuda") if world > 1 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    warm = max(args.warmup, 3)
    with ClockSampler(local) as clocks:                # clocks are sampled over warm-up + timed region
        fork_streams()
        for _ in range(warm * S):
            step()
        join_streams()
        if world > 1:                                  # the collective is warmed like everything else
            mdist.gather_records(g_src, out=gathered)
        barrier()
        state["i"] = 0
        ev0, evg, ev1 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        barrier()
        ev0.record()
        fork_streams()
        for _ in range(args.steps):
            step()
        join_streams()
        evg.record()
        if world > 1:
            mdist.gather_records(g_src, out=gathered)
        ev1.record()
        barrier()
    ms = ev0.elapsed_time(ev1)
    compute_ms, gather_ms = ev0.elapsed_time(evg), evg.elapsed_time(ev1)
    launches = launches_per_step * args.steps           # counted on an eager pass; a graph replays the same kernels
    per_rank = None
    gather_verified = oracle_verified = None
    if world > 1:
        t = torch.tensor([ms, compute_ms, gather_ms], device="cuda")
        allt = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        per_rank = {"total_ms": [float(x[0]) for x in allt], "compute_ms": [float(x[1]) for x in allt],
                    "gather_ms": [float(x[2]) for x in allt]}
        ms = max(per_rank["total_ms"])
        # what NCCL delivered in the timed region is what the ranks sent (own slice bit-equal, 64-bit checksums
        # of every slice exchanged); then a second, untimed gather of set 0's records, of which rank 0 checks
        # frame 0 of EVERY rank against the C restatement of the reference path
        gather_verified = mdist.verify_gather(gathered, g_src)
        if not args.frames:
            r0 = pipe.detect_and_align(d_loc, d_cls, d_fmaps, prefill=False)
            torch.cuda.synchronize()
            g2 = mdist.gather_records(pipe.record)
            gather_verified = gather_verified and mdist.verify_gather(g2, pipe.record)
            if rank == 0:
                oracle_verified = verify_against_oracle(wl, g2.cpu(), B, K, world)
    fps = world * B * args.steps / (ms * 1e-3)

    # ---- isolated pass: the same step on ONE stream with eager launches, for per-kernel times free of overlap
    # (prefill off: every kernel of the path alone on the GPU, brackets = CUDA events on the launching stream)
    def timed(fn, n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    iso_steps = max(10, min(50, args.steps))
    pipe.ctx.profile(True)

    def eager_plain():
        r_ = pipe.detect_and_align(d_loc, d_cls, d_fmaps, prefill=False)
        pipe.trim_and_paste(r_, d_masks)
    step_plain_ms = timed(eager_plain, iso_steps)
    iso = {k_: v_[0] / v_[1] for k_, v_ in pipe.ctx.profile_read().items()}
    iso["_step_ms"] = step_plain_ms
    pipe.ctx.profile(False)
    # the fill alone (one stream, nothing beside it) and the boxes-only paste behind a fill
    fill_ms = box_ms = None
    if pipe._can_prefill():
        fill_ms = timed(pipe.prefill_background, iso_steps)
        pipe.ctx.profile(True)

        def eager_prefilled():
            r_ = pipe.detect_and_align(d_loc, d_cls, d_fmaps, prefill=True)
            pipe.trim_and_paste(r_, d_masks)
        iso["_step_prefill_ms"] = timed(eager_prefilled, iso_steps)
        pf = pipe.ctx.profile_read()
        pipe.ctx.profile(False)
        box_ms = pf["paste"][0] / pf["paste"][1] if "paste" in pf else None

    # ---- the same step as ONE CUDA graph launch (single stream): what the launch gaps cost, with the
    # background fill forked inside the graph and without it
    graph_ms = graph_plain_ms = None
    if not args.no_graph:
        try:
            for pf_on in (True, False):
                if pf_on and not pipe._can_prefill():
                    continue
                g_, _ = pipe.capture(d_loc, d_cls, d_fmaps, d_masks, prefill=pf_on)
                v_ = timed(g_.replay, max(10, min(100, args.steps)))
                if pf_on:
                    graph_ms = v_
                else:
                    graph_plain_ms = v_
                del g_
        except Exception as exc:                     # pragma: no cover
            graph_ms = f"capture failed: {exc!r}"
    # the headline single-stream figure is the configuration of the timed region (--prefill); the other one beside it
    graph_prefill_ms = graph_ms
    graph_ms = graph_prefill_ms if (args.prefill and graph_prefill_ms is not None) else graph_plain_ms

    # ---- roofline of the dominant kernel: the one that writes the [B,M,PH,PW] uint8 masks.  With prefill that
    # is paste_fill_kernel (the zero background: every byte of the output once), timed ALONE on one stream;
    # without it paste_kernel<uint8> (background + boxes in one kernel), timed alone in the eager pass above.
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_kind = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    paste_bytes = B * M * PH * PW
    use_fill = bool(args.prefill) and fill_ms is not None
    launch_ms = fill_ms if use_fill else iso.get("paste")
    kernel = ("paste_fill_kernel (CropAndPadMask's zero background, mlp_paste_prefill)" if use_fill
              else "paste_kernel<uint8> (CropAndPadMask + >0.5)")
    achieved = (paste_bytes / (launch_ms * 1e-3) / 1e9) if launch_ms else None
    step_bytes = algorithmic_bytes(wl, N, M)
    # DRAM bytes of that kernel from the committed ncu --set full capture - only when the capture was taken from
    # THIS build (hash of the CUDA sources) and this workload; otherwise null
    traffic = traffic_src = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic_r02.json")) as f:
            tj = json.load(f)
        if tj.get("source_hash") == source_hash() and tj.get("workload") == args.workload:
            traffic = tj["dram_bytes_per_launch"].get("paste_fill_kernel" if use_fill else "paste_kernel<1>")
            traffic_src = tj.get("source")
    except Exception:
        pass
    write_peak = 7232.0                               # cudaMemsetAsync on this pool (profiles/write_bw_r01.txt)
    roofline = {"bound": "hbm", "kernel": kernel,
                "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": (achieved / peak) if achieved else None, "traffic": traffic, "traffic_source": traffic_src,
                "source_hash": source_hash(),
                "peak_kind": peak_kind, "bytes_per_launch": paste_bytes,
                "peak_note": "the measured peak is a read+write COPY figure; a write-only stream (this kernel) "
                             "reaches 7232 GB/s with cudaMemsetAsync on the same pool (profiles/write_bw_r01.txt), "
                             "so frac can exceed 1; frac_of_write_only_peak is the stricter figure",
                "frac_of_write_only_peak": (achieved / write_peak) if achieved else None,
                "frac_of_spec_8000": (achieved / 8000.0) if achieved else None,       # nominal ~8 TB/s too
                "avg_launch_ms": launch_ms,
                "how": "CUDA events on the launching stream around the launch, kernel alone on the GPU "
                       "(single-stream pass of this process, after the timed region)",
                "paste_kernel_full_ms": iso.get("paste"), "paste_fill_ms": fill_ms, "paste_boxes_only_ms": box_ms,
                "whole_step": {"algorithmic_bytes": step_bytes,
                               "achieved": step_bytes / (ms / args.steps * 1e-3) / 1e9,
                               "frac": step_bytes / (ms / args.steps * 1e-3) / 1e9 / peak,
                               "frac_of_write_only_peak": step_bytes / (ms / args.steps * 1e-3) / 1e9 / write_peak,
                               "streams": S, "input_sets_per_stream": NSETS},
                "cuda_graph_single_stream_ms_per_step": graph_ms,
                "cuda_graph_single_stream_no_prefill_ms_per_step": graph_plain_ms,
                "cuda_graph_single_stream_prefill_ms_per_step": graph_prefill_ms,
                "latency_us_per_frame_single_stream_graph": (1e3 * graph_ms / B) if isinstance(graph_ms, float) else None,
                "single_stream": {"ms_per_step": iso.get("_step_ms"), "ms_per_step_prefill": iso.get("_step_prefill_ms"),
                                  "stage_ms": {k_: v_ for k_, v_ in iso.items() if not k_.startswith("_")}}}

    # ---- end to end through the public API with host buffers
    e2e = e2e_bits = e2e_clip = None
    if not args.no_e2e:
        e2e = run_e2e(args, wl, pipe, h_loc, h_cls, h_fmaps, h_masks, world, M, barrier)
        if PW % 8 == 0:
            # same path, masks delivered 1 bit per pixel (lossless, 8x fewer bytes over PCIe)
            del graphs, sets, pipes[1:]
            torch.cuda.empty_cache()
            cfg_b = ml.DetectionConfig(paste_output="bits", prefill=bool(args.prefill), **kwargs_of(wl))
            pipe_b = ml.PostProcessPipeline(cfgp, (H, W), (PH, PW), C, wl["Cf"], B, cfg_b, device=local,
                                            private_context=True)
            e2e_bits = run_e2e(args, wl, pipe_b, h_loc, h_cls, h_fmaps, h_masks, world, M, barrier, bits=True)
            del pipe_b
        # same path, masks delivered as box-clipped bit rows (a few MB per batch; lossless, expand_clipped on the host)
        r_ = pipe.detect_and_align(d_loc, d_cls, d_fmaps, prefill=False)
        used_ = pipe.trim_and_clip(r_, d_masks)[3]
        e2e_clip = run_e2e(args, wl, pipe, h_loc, h_cls, h_fmaps, h_masks, world, M, barrier,
                           clip_cap=max(1, int(used_[0].item())))

    # ---- the consumer of the masks (SummaryOutput) fused behind the tail
    summary_leg = None
    if not args.no_summary:
        h_seg = pin(synth.semantic_map(B, PH, PW, seed=500 + rank))
        d_seg = h_seg.cuda()
        h_img = pin(synth.road_frames(B, PH, PW, seed=600 + rank))
        d_img = h_img.cuda()

        def serve():
            r_ = pipe.detect_and_align(d_loc, d_cls, d_fmaps, prefill=False)       # no masks are written
            pipe.trim_and_summarize(r_, d_masks, d_seg)
            pipe.draw(r_, d_masks, d_img, INST_COLORS[:C], 0.3, seg_outs=d_seg, semantic_colors=SEM_COLORS,
                      semantic_alpha=0.3, boxes=True)
            pipe.encode()

        pipe.ctx.profile(False)
        serve()
        Mo = int(pipe.summary_m.item())
        jpeg_len = pipe.jpeg_len.cpu().numpy()
        torch.cuda.synchronize()
        pipe.ctx.profile(True)
        n_ = max(10, min(50, args.steps))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n_):
            serve()
        e1.record()
        torch.cuda.synchronize()
        st_ = {k_: v_[0] / v_[1] for k_, v_ in pipe.ctx.profile_read().items()}
        pipe.ctx.profile(False)
        ms_s = e0.elapsed_time(e1) / n_
        graph_s_ms = graph_seq_ms = None
        if not args.no_graph:
            g_, _ = pipe.capture_serving(d_loc, d_cls, d_fmaps, d_masks, d_seg, d_img, INST_COLORS[:C], 0.3,
                                         semantic_colors=SEM_COLORS, semantic_alpha=0.3, boxes=True)
            for _ in range(3):
                g_.replay()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(n_):
                g_.replay()
            e1.record()
            torch.cuda.synchronize()
            graph_s_ms = e0.elapsed_time(e1) / n_
            del g_
            g_, _ = pipe.capture_serving(d_loc, d_cls, d_fmaps, d_masks, d_seg, d_img, INST_COLORS[:C], 0.3,
                                         semantic_colors=SEM_COLORS, semantic_alpha=0.3, boxes=True,
                                         parallel_branches=False)
            for _ in range(3):
                g_.replay()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(n_):
                g_.replay()
            e1.record()
            torch.cuda.synchronize()
            graph_seq_ms = e0.elapsed_time(e1) / n_
            del g_
        summary_leg = {
            "what": "decode+NMS+RoIAlign, then the two consumers of the masks in the serving graph - SummaryOutput "
                    "and the DrawBoxes+DrawInstance+DrawSegmentation overlay - straight from the mask tiles "
                    "(trim_and_summarize + draw, no [B,M,PH,PW] tensor), then the JPEG encode of every overlay frame "
                    "(EncodeImageContent, bytes identical to libjpeg); single stream, inputs resident in HBM",
            "value": world * B / (ms_s * 1e-3), "unit": "frames/s", "ms_per_step": ms_s,
            "cuda_graph_ms_per_step": graph_s_ms, "cuda_graph_single_branch_ms_per_step": graph_seq_ms,
            "cuda_graph_value": (world * B / (graph_s_ms * 1e-3)) if graph_s_ms else None,
            "rows_per_image": Mo, "stage_ms": st_,
            "seg_bytes": int(h_seg.numel() * 4), "frame_bytes": int(h_img.numel()),
            "jpeg_bytes_per_frame_mean": float(jpeg_len.mean()), "jpeg_bytes_per_frame_max": int(jpeg_len.max())}
        if not args.no_e2e:
            summary_leg["e2e"] = run_e2e(args, wl, pipe, h_loc, h_cls, h_fmaps, h_masks, world, M, barrier,
                                         h_seg=h_seg, summary_rows=Mo, h_img=h_img, jpeg_cap=int(jpeg_len.max()))

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            frames = args.cpu_frames or 64
            sample = CpuSample(wl, frames)
            v0_, w0_ = sample.run(1)                        # warm-up (page faults, caches)
            rep = max(1, int(round(12.0 / max(w0_, 1e-3))))  # ~12 s of CPU work
            v, wall = sample.run(1, repeat=rep)
            cpu = {"value": v, "unit": "frames/s", "cores": 1, "kind": "port",
                   "sample": f"{frames * rep} frames ({frames} distinct) of workload {args.workload} "
                             f"through oracle/c (single thread, uint8 paste), {wall:.1f} s wall"}
        if summary_leg is not None and not args.no_cpu_baseline and world == 1:
            summary_leg["cpu_port"] = cpu_serving_tail(wl, frames=2)
        line = {
            "metric": "frames/sec decode+NMS+RoIAlign+mask-paste", "value": fps, "unit": "frames/s",
            "n_gpus": world, "steps": args.steps, "warmup": warm,
            "compute_ms": compute_ms, "gather_ms": gather_ms, "per_rank": per_rank,
            "gather_verified": gather_verified, "gather_oracle_verified": oracle_verified,
            "gather": (None if world == 1 else {"collective": "one all_gather_into_tensor (NCCL) of the packed records",
                                                "bytes_per_rank": int(g_src.numel() * 4), "warmed": True}),
            "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if args.frames else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, wl), "roofline": roofline, "cpu_baseline": cpu,
            "e2e": e2e, "e2e_bitpacked": e2e_bits, "e2e_boxclip": e2e_clip, "serving_tail": summary_leg, "gpu_launches": int(launches), "clocks": clocks.summary(),
            "detections": {"M": M, "R": R, "Mf": mf, "kept_per_image_mean": float(counts.mean())},
            "device_bytes": pipe.device_bytes(),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def cpu_serving_tail(wl, frames=2):
    """CPU restatement of the serving tail on a bounded sample: the C restatement of the path up to the
    pasted float32 masks, then the NumPy restatements of SummaryOutput and of the three overlay layers
    (oracle/summary_oracle.py, oracle/draw_oracle.py), one frame at a time on one thread."""
    import synth
    from oracle import c_oracle as co, summary_oracle as so, draw_oracle as do
    t_path = t_cons = t_jpeg = 0.0
    cpu_jpeg(np.zeros((16, 16, 3), dtype=np.uint8))                  # library import outside the timed part
    for i in range(frames):
        cfgp, N, loc, cls, fmaps = make_inputs(wl, 1, 9000 + i)
        pool = synth.mask_probs(1, (wl["max_k"] + 1) * wl["nms_max_output_size"], wl["C"], seed=98)
        seg = synth.semantic_map(1, wl["PH"], wl["PW"], seed=9100 + i)
        img = synth.road_frames(1, wl["PH"], wl["PW"], seed=9200 + i)
        t0 = time.perf_counter()
        out = co.full_path(loc, cls, fmaps, lambda f, b: pool[:, :b.shape[1]], cfgp, (wl["H"], wl["W"]),
                           (wl["PH"], wl["PW"]), binary=False, **kwargs_of(wl))
        t1 = time.perf_counter()
        so.summary_output(out["det_i"], seg, out["pasted"])
        vis = do.draw_boxes(img, out["det_i"])
        vis = do.draw_instance(vis, out["det_i"], out["pasted"], INST_COLORS[:wl["C"]], 0.3)
        vis = do.draw_segmentation(vis, seg, SEM_COLORS, 0.3)
        t2 = time.perf_counter()
        jpeg_kind = cpu_jpeg(vis[0])
        t3 = time.perf_counter()
        t_path += t1 - t0
        t_cons += t2 - t1
        t_jpeg += t3 - t2
    return {"value": frames / (t_path + t_cons + t_jpeg), "unit": "frames/s", "cores": 1, "kind": "port",
            "sample": f"{frames} frames of workload through oracle/c (path, float32 paste) + NumPy restatements "
                      f"of SummaryOutput / DrawBoxes / DrawInstance / DrawSegmentation + JPEG encode ({jpeg_kind})",
            "consumers_ms_per_frame": 1e3 * t_cons / frames, "path_ms_per_frame": 1e3 * t_path / frames,
            "jpeg_ms_per_frame": 1e3 * t_jpeg / frames, "jpeg_kind": jpeg_kind}


def cpu_jpeg(frame):
    """JPEG encode of one frame on the CPU: libjpeg-turbo itself through Pillow when it is importable (the library
    tf.io.encode_jpeg links, same parameters), else the NumPy restatement."""
    try:
        import io
        from PIL import Image
        buf = io.BytesIO()
        Image.fromarray(frame).save(buf, format="JPEG", quality=95, dpi=(300, 300))
        return "libjpeg-turbo via Pillow (the reference's library)"
    except ImportError:
        from oracle import jpeg_oracle as jo
        jo.encode_jpeg(frame)
        return "oracle/jpeg_oracle.py (NumPy restatement)"


def verify_against_oracle(wl, gathered, B, K, world):
    """Rank 0: frame 0 of EVERY rank's gathered detection records against the C restatement of the reference
    path on that rank's own seeded inputs (regenerated here on the host)."""
    from masklab_b200 import dist as mdist
    from oracle import c_oracle as co
    det, counts = mdist.unpack_records(gathered, B, K)
    kw = {k: wl[k] for k in ("min_confidence", "nms_iou_threshold", "post_iou_threshold", "nms_max_output_size")}
    ok = True
    for r in range(world):
        cfgp, N, loc, cls, _ = make_inputs(wl, B, seed=100 + r)
        prior = co.prior_layer(cfgp, wl["H"], wl["W"])
        boxes = co.restore_boxes(loc[:1], prior)
        want = co.detection_proposal(cls[:1], boxes, **kw)               # [1,M0,6], -1 padded to its own count
        n = int(counts[r * B])
        got = det[r * B, :n].numpy()
        m0 = int((want[0, :, 4] != -1).sum())
        ok = ok and n == m0 and bool((got == want[0, :n]).all())
    return bool(ok)


def algorithmic_bytes(wl, N, M):
    """read each required input once, write each required output once."""
    B, C, Cf = wl["B"], wl["C"], wl["Cf"]
    decode = B * N * 16 + B * N * C * 4
    det = B * M * 24
    fm = sum(B * (-(-wl["H"] // s)) * (-(-wl["W"] // s)) * Cf * 4 for s in (8, 16, 32)[:wl["max_k"] + 1])
    crops = B * M * 14 * 14 * Cf * 4
    trim = B * M * 28 * 28 * 4
    paste = B * M * wl["PH"] * wl["PW"]
    return decode + det + fm + crops + trim + paste


INST_COLORS = [[192, 32, 128], [160, 96, 0], [96, 0, 128], [32, 96, 192], [96, 32, 128], [64, 64, 64]]
SEM_COLORS = [[64, 0, 128], [128, 96, 0], [128, 192, 0]]      # engine/config.py:32-42


def run_e2e(args, wl, pipe, h_loc, h_cls, h_fmaps, h_masks, world, M, barrier, bits=False, h_seg=None,
            summary_rows=0, h_img=None, jpeg_cap=0, clip_cap=0):
    """Same metric through PostProcessPipeline with HOST buffers: every step copies the inputs
    from pinned host memory, runs the path and reads detections + binary masks back into pinned
    host memory.  Three streams (copy-in, compute, copy-out) and two sets of device input buffers,
    so that the input copy of step i+1 overlaps the compute and the output copy of step i (PCIe
    is full duplex); every byte of every step still crosses PCIe inside the timed region."""
    import torch
    import torch.distributed as dist
    B, PH, PW = wl["B"], wl["PH"], wl["PW"]
    # two sets of device input buffers: the copy-in of step i+1 runs while step i computes
    NB = 2
    d_in = [dict(loc=torch.empty_like(h_loc, device="cuda"), cls=torch.empty_like(h_cls, device="cuda"),
                 fmaps=[torch.empty_like(f, device="cuda") for f in h_fmaps],
                 masks=torch.empty_like(h_masks, device="cuda"),
                 seg=(torch.empty_like(h_seg, device="cuda") if h_seg is not None else None),
                 img=(torch.empty_like(h_img, device="cuda") if h_img is not None else None),
                 free=None) for _ in range(NB)]
    summary = h_seg is not None
    clip = clip_cap > 0
    out_det = torch.empty((B * (pipe.K if (summary or clip) else M) * 6,), dtype=torch.int32).pin_memory()
    if summary:
        out_masks = torch.empty((B * summary_rows * 11,), dtype=torch.float32).pin_memory()
    elif clip:
        # box-clipped bit rows: a fixed pool bound (1.25 x what the warm-up batch needed, 64 KB granules) is copied
        # without asking the device for the size first; the geometry records and the size travel with it
        clip_cap = (int(clip_cap * 1.25) + 65535) // 65536 * 65536
        out_masks = torch.empty((clip_cap,), dtype=torch.uint8).pin_memory()
        out_geom = torch.empty((B * pipe.K * 8,), dtype=torch.int32).pin_memory()
        out_used = torch.empty((2,), dtype=torch.int64).pin_memory()
    else:
        out_masks = torch.empty((B * M * PH * (PW // 8 if bits else PW),), dtype=torch.uint8).pin_memory()
    # the overlay leaves as JPEG files: a fixed per-frame bound (1.25 x the largest file of the warm-up step, 64 KB
    # granules) is copied without asking the device for the lengths first; the lengths travel with it
    jpeg_cap = (int(jpeg_cap * 1.25) + 65535) // 65536 * 65536 if h_img is not None else 0
    out_vis = torch.empty((B, jpeg_cap), dtype=torch.uint8).pin_memory() if h_img is not None else None
    out_len = torch.empty((B,), dtype=torch.int32).pin_memory() if h_img is not None else None
    h2d = sum(t.numel() * t.element_size() for t in [h_loc, h_cls, h_masks] + h_fmaps + ([h_seg] if summary else [])
              + ([h_img] if h_img is not None else []))
    d2h = out_det.numel() * 4 + out_masks.numel() * out_masks.element_size() + (out_vis.numel() + 4 * B if h_img is not None else 0)
    if clip:
        d2h += out_geom.numel() * 4 + 16
    s_in, s_cmp, s_out = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
    state = {"i": 0, "out_done": None}

    def step():
        buf = d_in[state["i"] % NB]
        state["i"] += 1
        with torch.cuda.stream(s_in):
            if buf["free"] is not None:
                s_in.wait_event(buf["free"])                # the step that used this set has computed
            buf["loc"].copy_(h_loc, non_blocking=True)
            buf["cls"].copy_(h_cls, non_blocking=True)
            for d, h in zip(buf["fmaps"], h_fmaps):
                d.copy_(h, non_blocking=True)
            buf["masks"].copy_(h_masks, non_blocking=True)
            if summary:
                buf["seg"].copy_(h_seg, non_blocking=True)
            if h_img is not None:
                buf["img"].copy_(h_img, non_blocking=True)
            ev_in = torch.cuda.Event()
            ev_in.record(s_in)
        with torch.cuda.stream(s_cmp):
            s_cmp.wait_event(ev_in)
            if state["out_done"] is not None:
                s_cmp.wait_event(state["out_done"])         # previous results copied out
            r = pipe.detect_and_align(buf["loc"], buf["cls"], buf["fmaps"], prefill=(False if (summary or clip) else None))
            vis = None
            if summary:
                det_i32, pasted, _ = pipe.trim_and_summarize(r, buf["masks"], buf["seg"])
                if h_img is not None:
                    pipe.draw(r, buf["masks"], buf["img"], INST_COLORS[:wl["C"]], 0.3, seg_outs=buf["seg"],
                              semantic_colors=SEM_COLORS, semantic_alpha=0.3, boxes=True)
                    vis, vis_len = pipe.encode()
            elif clip:
                det_i32, geom, pasted, used = pipe.trim_and_clip(r, buf["masks"], pool_bytes=clip_cap)
            else:
                det_i32, pasted, _ = pipe.trim_and_paste(r, buf["masks"])
            cmp_done = torch.cuda.Event()
            cmp_done.record(s_cmp)
            buf["free"] = cmp_done
        with torch.cuda.stream(s_out):
            s_out.wait_event(cmp_done)
            out_det.copy_(det_i32[:out_det.numel()], non_blocking=True)
            out_masks.copy_(pasted[:out_masks.numel()], non_blocking=True)
            if clip:
                out_geom.copy_(geom.view(-1), non_blocking=True)
                out_used.copy_(used, non_blocking=True)
            if vis is not None:
                out_vis.copy_(vis[:, :jpeg_cap], non_blocking=True)
                out_len.copy_(vis_len, non_blocking=True)
            state["out_done"] = torch.cuda.Event()
            state["out_done"].record(s_out)

    torch.cuda.synchronize()
    step()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(s_in)
    for _ in range(args.e2e_steps):
        step()
    s_out.wait_stream(s_cmp)
    ev1.record(s_out)
    barrier()
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    step_ms = ms / args.e2e_steps
    return {"value": world * B * args.e2e_steps / (ms * 1e-3), "unit": "frames/s",
            "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": args.e2e_steps,
            "ms_per_step": step_ms,
            # both directions run concurrently (PCIe is full duplex): the busier one bounds the step
            "pcie_gbs_busier_direction": max(h2d, d2h) / (step_ms * 1e-3) / 1e9,
            "bound": "PCIe (host<->device copies of every step), not the kernels",
            "pool_bytes_used": (int(out_used[0]) if clip else None),
            "api": ("PostProcessPipeline.detect_and_align + trim_and_clip; pinned host inputs in, int32 detections + "
                    "box-clipped bit rows (geometry records + byte pool, expand_clipped rebuilds the dense masks) out"
                    if clip else
                    "PostProcessPipeline.detect_and_align + trim_and_summarize" + (" + draw + encode" if h_img is not None else "")
                    + "; pinned host inputs (heads, FPN maps, mask-head output, semantic map"
                    + (", frames" if h_img is not None else "") + ") in, int32 detections + [B,M',11] summary"
                    + (" + the JPEG files of the overlay (fixed per-frame bound)" if h_img is not None else "")
                    + " out every step; the [B,M,PH,PW] "
                    "masks are never written" if summary else
                    "PostProcessPipeline.detect_and_align + trim_and_paste; pinned host inputs in, "
                    "int32 detections + " + ("bit-packed (1 bit/pixel)" if bits else "uint8") +
                    " masks out to pinned host memory every step")}


def main():
    args = parse_args()
    wl = WORKLOADS[args.workload]
    # The contract is ONE JSON line on stdout.  Libraries (NCCL prints its version banner on
    # stdout at communicator creation) must not interleave with it: route fd 1 to stderr while
    # the benchmark runs and hand the real stdout back only for the final line.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    out = io.StringIO()
    try:
        with contextlib.redirect_stdout(out):
            if args.impl == "reference":
                run_reference(args, wl)
            else:
                run_ours(args, wl)
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    lines = [l for l in out.getvalue().splitlines() if l.strip()]
    json_lines = [l for l in lines if l.lstrip().startswith("{")]
    for l in lines:
        if l not in json_lines:
            print(l, file=sys.stderr)
    for l in json_lines[-1:]:
        print(l, flush=True)


if __name__ == "__main__":
    main()
