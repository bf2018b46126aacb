"""PostProcessPipeline — the whole post-backbone path as one sync-free device pipeline.

It is the wiring of /root/reference/engine/retinamasklab.py:458-470 (RestoreBoxes ->
DetectionProposal -> MaskDistribute -> PyramidRoiAlign), :615-616 (TrimInstances),
:635-636 (UpSampleOutput) and /root/reference/road_project/setup/serving.py:30
(CropAndPadMask), but on fixed-capacity device buffers: the data-dependent sizes
(M kept boxes, Mf RoIs per level, R = sum Mf) stay on the device as int32 scalars that
the next kernel reads, so a batch is enqueued without any host round trip.  The mask
head (MaskSubNet, dense convolutions) is not part of this path: the caller runs it
between `detect_and_align` and `trim_and_paste`.
"""
import ctypes
import os
from dataclasses import dataclass, field

import torch

from . import runtime as rt
from .prior import PriorBoxes


@dataclass
class DetectionConfig:
    """Hyper-parameters of the path; defaults are the reference's layer defaults
    (engine/layers/detection.py:469-473, engine/layers/instance.py:47,104)."""
    min_confidence: float = 0.05
    nms_iou_threshold: float = 0.4
    post_iou_threshold: float = 0.65
    nms_max_output_size: int = 1000
    max_k: int = 2
    base_size: float = 64
    crop_size: tuple = (14, 14)
    mask_size: tuple = (28, 28)
    padding: str = "same"
    strict_batch: bool = True
    paste_output: str = "uint8"          # 'uint8' (binary, > 0.5 fused), 'bits' (8 px/byte) or 'float32' (drop-in)
    fused: bool = True                   # fused halves (6 kernels) vs the chain of stage kernels (13)
    # Zero the background of the [B,M,PH,PW] masks on a second stream right after the NMS kernels, beside
    # RoIAlign / the mask head, and let the paste kernel write the boxes only (mlp_paste_prefill).  Only
    # worth it when trim_and_paste follows; detect_and_align(prefill=...) overrides it per call.
    prefill: bool = False
    # layout of the mask head output handed to the tail: 'interleaved' = the reference's [B,R,mh,mw,C];
    # 'planar' = [B,R,C,mh,mw] (channels-first head): the tail then reads 1/C of the bytes
    mask_layout: str = "interleaved"


@dataclass
class AlignedRois:
    """Device-side result of the first half (everything MaskSubNet needs)."""
    det: torch.Tensor             # [B,K,6] f32, -1 padded (K = nms_max_output_size capacity)
    keep: torch.Tensor            # [B,K,2] i32 (anchor index, class), -1 padded
    counts: torch.Tensor          # [B] i32
    m_dev: torch.Tensor           # [1] i32  M = max(1, max counts)
    dist: torch.Tensor            # [B,K,7] f32
    level_counts: torch.Tensor    # [L,B] i32
    level_m: torch.Tensor         # [L+1] i32: Mf per level, then R = sum Mf
    crops: list = field(default_factory=list)   # per level flat capacity buffers, valid prefix [B,Mf,ch,cw,Cf]
    roi_boxes: torch.Tensor = None               # flat capacity buffer, valid prefix [B,R,6]

    def shapes(self):
        """One D2H of L+1 ints: (Mf list, R)."""
        v = self.level_m.tolist()
        return v[:-1], v[-1]


class PostProcessPipeline:
    def __init__(self, prior, image_hw, frame_hw, num_classes, fpn_channels, batch,
                 config=None, device=None, private_context=False):
        self.cfg = config or DetectionConfig()
        self.prior = prior if isinstance(prior, PriorBoxes) else PriorBoxes(**prior)
        self.image_hw = (int(image_hw[0]), int(image_hw[1]))
        self.frame_hw = (int(frame_hw[0]), int(frame_hw[1]))
        self.C = int(num_classes)
        self.Cf = int(fpn_channels)
        self.B = int(batch)
        # a context owns scratch and must not be shared between streams: pipelines that run
        # concurrently on different streams each take a private one
        if private_context:
            dev = rt.Context.get(device).device
            self.ctx = rt.Context(dev)
        else:
            self.ctx = rt.Context.get(device)
        self.lib = self.ctx.lib
        self.L = int(self.cfg.max_k) + 1
        self.K = int(self.cfg.nms_max_output_size)
        self.prior_c = self.prior.to_c(self.cfg.padding)
        self.N = int(self.lib.mlp_prior_count(ctypes.byref(self.prior_c), *self.image_hw))
        if self.N < 0:
            rt.check(self.N)
        groups = self.prior.grouped()
        if self.L > len(groups):
            raise rt.InvalidArgumentError(rt.MLP_EINVAL, "max_k + 1 exceeds the number of pyramid levels")
        same = self.cfg.padding == "same"
        H, W = self.image_hw
        self.fmap_hw = [((H + s - 1) // s if same else H // s, (W + s - 1) // s if same else W // s)
                        for s, _ in groups[:self.L]]
        if any(h < 1 or w < 1 for h, w in self.fmap_hw):
            raise rt.InvalidArgumentError(
                rt.MLP_EINVAL, f"pyramid level without a single cell for a {H}x{W} image with padding="
                f"{self.cfg.padding!r}: {self.fmap_hw} (tf.image.crop_and_resize rejects an empty image too)")
        self.params = rt.DetectionParamsC(
            float(self.cfg.min_confidence), float(self.cfg.nms_iou_threshold),
            float(self.cfg.post_iou_threshold), self.K, 1 if self.cfg.strict_batch else 0)
        if self.cfg.mask_layout not in ("interleaved", "planar"):
            raise rt.InvalidArgumentError(rt.MLP_EINVAL, f"unknown mask_layout {self.cfg.mask_layout!r}")
        if self.cfg.mask_layout == "planar" and not self.cfg.fused:
            raise rt.InvalidArgumentError(rt.MLP_EINVAL, "mask_layout='planar' needs the fused tail (fused=True)")
        self._tail_flags = rt.MLP_MASKS_PLANAR if self.cfg.mask_layout == "planar" else 0
        self._side = self._main = None   # internal streams of the overlapped background fill (created on first use)
        self._fill_done = None           # event: the fill of the batch in flight has been enqueued / finished
        self._alloc()

    # ------------------------------------------------------------------ buffers
    def _alloc(self):
        c, B, K, L = self.ctx, self.B, self.K, self.L
        ch, cw = self.cfg.crop_size
        mh, mw = self.cfg.mask_size
        f32, i32 = torch.float32, torch.int32
        # det + counts live in ONE buffer (masklab_b200.dist.record_*): the cross-class NMS kernel writes the
        # multi-GPU gather's send buffer directly, no staging copy before the collective
        from . import dist as mdist
        self.record = c.empty((mdist.record_words(B, K)[1],), i32)
        self.record.zero_()
        self.det, self.counts = mdist.record_views(self.record, B, K)
        self.keep = c.empty((B, K, 2), i32)
        self.m_dev = c.empty((1,), i32)
        self.dist = c.empty((B, K, 7), f32)
        self.level_counts = c.empty((L, B), i32)
        self.level_m = c.empty((L + 1,), i32)
        self.crops = [c.empty((B * K * ch * cw * self.Cf,), f32) for _ in range(L)]
        self.roi_boxes = c.empty((B * L * K * 6,), f32)
        self.trim_counts = c.empty((B,), i32)
        self.trim_m = c.empty((1,), i32)
        self.trim_m.zero_()              # M of the previous batch: sizes the speculative background fill
        self.trim_boxes = c.empty((B * K * 6,), f32)
        self.trim_masks = c.empty((B * K * mh * mw,), f32)
        self.det_i32 = c.empty((B * K * 6,), i32)
        self.masks_i32 = c.empty((B * K * mh * mw,), i32)
        PH, PW = self.frame_hw
        mode = self.cfg.paste_output
        if mode not in ("uint8", "float32", "bits"):
            raise rt.InvalidArgumentError(rt.MLP_EINVAL, f"unknown paste_output {mode!r}")
        if mode == "bits" and PW % 8:
            raise rt.InvalidArgumentError(rt.MLP_EINVAL, "bit-packed output needs frame width % 8 == 0")
        self.paste_mode = {"float32": rt.MLP_PASTE_F32, "uint8": rt.MLP_PASTE_U8, "bits": rt.MLP_PASTE_BITS}[mode]
        self.paste_row = PW // 8 if mode == "bits" else PW          # elements per frame row of `pasted`
        self.pasted = c.empty((B * K * PH * self.paste_row,), f32 if mode == "float32" else torch.uint8)
        self._crop_ptrs = (ctypes.c_void_p * L)(*[c.view(t).value for t in self.crops])
        self._fh = (ctypes.c_int32 * L)(*[h for h, _ in self.fmap_hw])
        self._fw = (ctypes.c_int32 * L)(*[w for _, w in self.fmap_hw])

    def device_bytes(self):
        ts = [self.det, self.keep, self.dist, self.roi_boxes, self.trim_boxes, self.trim_masks,
              self.det_i32, self.masks_i32, self.pasted] + self.crops
        return sum(t.numel() * t.element_size() for t in ts) + self.ctx.scratch_bytes()

    # --------------------------------------------------------------- first half
    def _can_prefill(self):
        vec = {rt.MLP_PASTE_F32: 4, rt.MLP_PASTE_U8: 16, rt.MLP_PASTE_BITS: 128}[self.paste_mode]
        return self.cfg.fused and self.frame_hw[1] % vec == 0

    def detect_and_align(self, loc_pred, cls_pred, fmaps, prefill=None):
        """loc_pred [B,N,4], cls_pred [B,N,C], fmaps: L tensors [B,Hf,Wf,Cf] (NHWC), all f32
        CUDA.  Enqueues a2-a10; returns AlignedRois (device tensors, no host sync).  With prefill
        (default: config.prefill) the zero background of the masks is streamed out on a second
        stream from the moment the NMS kernels have produced M; trim_and_paste joins it."""
        c, lib, B, K, L = self.ctx, self.lib, self.B, self.K, self.L
        st = c.stream()
        prefill = (self.cfg.prefill if prefill is None else bool(prefill)) and self._can_prefill()
        self._fill_done = None
        if tuple(loc_pred.shape) != (B, self.N, 4) or tuple(cls_pred.shape) != (B, self.N, self.C):
            raise rt.InvalidArgumentError(
                rt.MLP_EINVAL, f"expected loc {(B, self.N, 4)} / cls {(B, self.N, self.C)}, got "
                f"{tuple(loc_pred.shape)} / {tuple(cls_pred.shape)}")
        if len(fmaps) < L:
            raise rt.InvalidArgumentError(rt.MLP_EINVAL, f"need {L} FPN maps, got {len(fmaps)}")
        for f, (h, w) in zip(fmaps[:L], self.fmap_hw):
            if tuple(f.shape) != (B, h, w, self.Cf):
                raise rt.InvalidArgumentError(
                    rt.MLP_EINVAL, f"FPN map {tuple(f.shape)} != {(B, h, w, self.Cf)}")
        if self.cfg.fused and prefill:
            fmap_ptrs = (ctypes.c_void_p * L)(*[c.view(f, torch.float32).value for f in fmaps[:L]])
            ch, cw = self.cfg.crop_size
            # Two internal streams.  `side` (least priority) streams the zero background of the masks out:
            # first SPECULATIVELY, sized by the previous batch's M (a longer prefix than needed is only
            # wasted work), beside the latency-bound NMS kernels, then the rest once this batch's counts
            # exist.  `main` (high priority) runs the chain itself, so that its CTAs are placed ahead of the
            # fill's 25,600 short ones whatever the priority of the caller's stream.
            cur = torch.cuda.current_stream(c.device)
            if self._side is None:
                self._side = torch.cuda.Stream(device=c.device, priority=0)
                self._main = torch.cuda.Stream(device=c.device, priority=-1)
            side, main = self._side, self._main
            start = torch.cuda.Event()
            start.record(cur)
            side.wait_event(start)
            main.wait_event(start)
            null = ctypes.c_void_p(None)
            with torch.cuda.stream(side):
                rt.check(lib.mlp_paste_prefill(
                    c.handle, null, null, c.view(self.trim_m), B, K, self.frame_hw[0], self.frame_hw[1],
                    self.paste_mode, c.view(self.pasted), ctypes.c_void_p(side.cuda_stream)))
            with torch.cuda.stream(main):
                rt.check(lib.mlp_detect_plan(
                    c.handle, ctypes.byref(self.prior_c), c.view(loc_pred, torch.float32),
                    c.view(cls_pred, torch.float32), B, self.image_hw[0], self.image_hw[1], self.C,
                    ctypes.byref(self.params), int(self.cfg.max_k), float(self.cfg.base_size),
                    c.view(self.det), c.view(self.keep), c.view(self.counts), c.view(self.m_dev),
                    c.view(self.dist), c.view(self.level_counts), c.view(self.level_m),
                    ctypes.c_void_p(main.cuda_stream)))
                nms_done = torch.cuda.Event()
                nms_done.record(main)
            side.wait_event(nms_done)
            with torch.cuda.stream(side):
                rt.check(lib.mlp_paste_prefill(
                    c.handle, c.view(self.trim_m), c.view(self.counts), null, B, K, self.frame_hw[0],
                    self.frame_hw[1], self.paste_mode, c.view(self.pasted), ctypes.c_void_p(side.cuda_stream)))
                self._fill_done = torch.cuda.Event()
                self._fill_done.record(side)
            with torch.cuda.stream(main):
                rt.check(lib.mlp_roi_align_run(
                    c.handle, fmap_ptrs, self._fh, self._fw, L, self.Cf, c.view(self.dist), B, K, K,
                    c.view(self.m_dev), float(self.image_hw[0]), float(self.image_hw[1]), int(ch), int(cw),
                    c.view(self.level_counts), c.view(self.level_m), self._crop_ptrs, c.view(self.roi_boxes),
                    ctypes.c_void_p(main.cuda_stream)))
                aligned = torch.cuda.Event()
                aligned.record(main)
            cur.wait_event(aligned)
            return AlignedRois(self.det, self.keep, self.counts, self.m_dev, self.dist,
                               self.level_counts, self.level_m, self.crops, self.roi_boxes)
        if self.cfg.fused:
            fmap_ptrs = (ctypes.c_void_p * L)(*[c.view(f, torch.float32).value for f in fmaps[:L]])
            ch, cw = self.cfg.crop_size
            rt.check(lib.mlp_detect_align(
                c.handle, ctypes.byref(self.prior_c), c.view(loc_pred, torch.float32),
                c.view(cls_pred, torch.float32), B, self.image_hw[0], self.image_hw[1], self.C,
                ctypes.byref(self.params), int(self.cfg.max_k), float(self.cfg.base_size), fmap_ptrs,
                self._fh, self._fw, self.Cf, int(ch), int(cw), c.view(self.det), c.view(self.keep),
                c.view(self.counts), c.view(self.m_dev), c.view(self.dist), c.view(self.level_counts),
                c.view(self.level_m), self._crop_ptrs, c.view(self.roi_boxes), st))
            return AlignedRois(self.det, self.keep, self.counts, self.m_dev, self.dist,
                               self.level_counts, self.level_m, self.crops, self.roi_boxes)
        rt.check(lib.mlp_detect_from_heads(
            c.handle, ctypes.byref(self.prior_c), c.view(loc_pred, torch.float32),
            c.view(cls_pred, torch.float32), B, self.image_hw[0], self.image_hw[1], self.C,
            ctypes.byref(self.params), c.view(self.det), c.view(self.keep), c.view(self.counts),
            c.view(self.m_dev), st))
        rt.check(lib.mlp_mask_distribute(c.handle, c.view(self.det), B * K, int(self.cfg.max_k),
                                         float(self.cfg.base_size), c.view(self.dist), st))
        rt.check(lib.mlp_roi_align_plan(c.handle, c.view(self.dist), B, K, K, c.view(self.m_dev), L,
                                        c.view(self.level_counts), c.view(self.level_m), st))
        fmap_ptrs = (ctypes.c_void_p * L)(*[c.view(f, torch.float32).value for f in fmaps[:L]])
        ch, cw = self.cfg.crop_size
        rt.check(lib.mlp_roi_align_run(
            c.handle, fmap_ptrs, self._fh, self._fw, L, self.Cf, c.view(self.dist), B, K, K,
            c.view(self.m_dev), float(self.image_hw[0]), float(self.image_hw[1]), int(ch), int(cw),
            c.view(self.level_counts), c.view(self.level_m), self._crop_ptrs, c.view(self.roi_boxes),
            st))
        return AlignedRois(self.det, self.keep, self.counts, self.m_dev, self.dist, self.level_counts,
                           self.level_m, self.crops, self.roi_boxes)

    def prefill_background(self):
        """Enqueue CropAndPadMask's zero background for the current detections on the CURRENT stream (no fork);
        detect_and_align(prefill=True) does this on a second stream.  For measuring the fill alone."""
        c = self.ctx
        null = ctypes.c_void_p(None)
        rt.check(self.lib.mlp_paste_prefill(c.handle, null, c.view(self.counts), null, self.B, self.K,
                                            self.frame_hw[0], self.frame_hw[1], self.paste_mode, c.view(self.pasted),
                                            c.stream()))

    def roi_views(self, rois):
        """Reference-shaped views ([B,Mf,ch,cw,Cf] per level, [B,R,6]) — one small D2H."""
        mf, R = rois.shapes()
        ch, cw = self.cfg.crop_size
        crops = [t[:self.B * m * ch * cw * self.Cf].view(self.B, m, ch, cw, self.Cf)
                 for t, m in zip(rois.crops, mf)]
        boxes = rois.roi_boxes[:self.B * R * 6].view(self.B, R, 6)
        return crops, boxes

    # -------------------------------------------------------------- second half
    def trim_and_paste(self, rois, roi_masks):
        """roi_masks: mask-head output, f32, dense [B,R,mh,mw,C] (R = rois.level_m[-1]; a flat
        buffer whose prefix has that layout is fine).  Enqueues a11-a14.  Returns
        (det_i32 flat, pasted flat, m_dev): pasted has the valid prefix [B,M,PH,PW]; det_i32 is
        [B,K,6] (capacity rows, fused path) or has the valid prefix [B,M,6] (stage path) - use
        result_views() for reference-shaped tensors."""
        c, lib, B, K, L = self.ctx, self.lib, self.B, self.K, self.L
        st = c.stream()
        mh, mw = self.cfg.mask_size
        r_cap = L * K
        r_dev = ctypes.c_void_p(c.view(rois.level_m).value + 4 * L)
        masks_ptr = c.view(roi_masks, torch.float32)
        ratio = (torch.tensor([float(self.frame_hw[0]), float(self.frame_hw[1])], dtype=torch.float32)
                 / torch.tensor([float(self.image_hw[0]), float(self.image_hw[1])], dtype=torch.float32))
        mode = self.paste_mode
        self._compact_det = not self.cfg.fused
        if self.cfg.fused:
            mode |= self._tail_flags
            if self._fill_done is not None:                  # join the background fill of this batch
                torch.cuda.current_stream(c.device).wait_event(self._fill_done)
                self._fill_done = None
                mode |= rt.MLP_PASTE_PREFILLED
            rt.check(lib.mlp_trim_paste(
                c.handle, c.view(rois.roi_boxes), masks_ptr, B, r_cap, r_dev, mh, mw, self.C,
                float(ratio[0]), float(ratio[1]), K, self.frame_hw[0], self.frame_hw[1], mode,
                c.view(self.det_i32), c.view(self.trim_counts), c.view(self.trim_m), c.view(self.pasted),
                st))
            return self.det_i32, self.pasted, self.trim_m
        rt.check(lib.mlp_trim_plan(c.handle, c.view(rois.roi_boxes), B, r_cap, r_dev,
                                   c.view(self.trim_counts), c.view(self.trim_m), st))
        rt.check(lib.mlp_trim_run(c.handle, c.view(rois.roi_boxes), masks_ptr, B, r_cap, r_dev, mh, mw,
                                  self.C, c.view(self.trim_m), c.view(self.trim_boxes),
                                  c.view(self.trim_masks), st))
        rt.check(lib.mlp_upsample_output(
            c.handle, c.view(self.trim_boxes), B * K, float(ratio[0]), float(ratio[1]),
            c.view(self.det_i32), c.view(self.trim_masks), B * K * mh * mw, c.view(self.masks_i32), st))
        rt.check(lib.mlp_crop_and_pad_mask(
            c.handle, c.view(self.det_i32), c.view(self.masks_i32), B, K, 0, c.view(self.trim_m), mh, mw,
            self.frame_hw[0], self.frame_hw[1], mode, c.view(self.pasted), st))
        return self.det_i32, self.pasted, self.trim_m

    def trim_and_clip(self, rois, roi_masks, pool_bytes=None):
        """The second half with the masks in box-clipped form (csrc/clip.cu): TrimInstances + UpSampleOutput,
        then per instance its clipped box and the bit rows of (CropAndPadMask value > 0.5) INSIDE that box - the
        tf.pad zeros of misc.py:393-394 are never materialised.  Returns (det_i32 [B,K,6], geom int32 [B,K,8],
        pool uint8 [capacity], used int64 [2] on the device): cfg-2 needs 2.4 MB of pool where the dense masks
        are 1.68 GB.  `expand_clipped` rebuilds the dense [B,M,PH,PW] uint8 tensor bit for bit.  pool_bytes:
        capacity of the pool (default: a bound that cannot overflow, B*K*PH*ceil(PW/8)); used[0] > capacity
        reports an overflow (rows past the capacity were dropped)."""
        c, lib, B, K, L = self.ctx, self.lib, self.B, self.K, self.L
        if not self.cfg.fused:
            raise rt.InvalidArgumentError(rt.MLP_EINVAL, "trim_and_clip needs the fused tail (fused=True)")
        mh, mw = self.cfg.mask_size
        PH, PW = self.frame_hw
        r_cap = L * K
        r_dev = ctypes.c_void_p(c.view(rois.level_m).value + 4 * L)
        masks_ptr = c.view(roi_masks, torch.float32)
        ratio = (torch.tensor([float(PH), float(PW)], dtype=torch.float32)
                 / torch.tensor([float(self.image_hw[0]), float(self.image_hw[1])], dtype=torch.float32))
        cap = int(pool_bytes) if pool_bytes is not None else int(lib.mlp_clip_pool_bound(B, K, PH, PW))
        if getattr(self, "clip_cap", None) != cap:
            self.clip_pool = c.empty(((cap + 15) // 16 * 16,), torch.uint8)
            self.clip_geom = c.empty((B, K, 8), torch.int32)
            self.clip_used = c.empty((2,), torch.int64)
            self.clip_cap = cap
        if self._fill_done is not None:                      # a background fill of this batch is in flight: join it
            torch.cuda.current_stream(c.device).wait_event(self._fill_done)
            self._fill_done = None
        rt.check(lib.mlp_trim_paste(
            c.handle, c.view(rois.roi_boxes), masks_ptr, B, r_cap, r_dev, mh, mw, self.C, float(ratio[0]),
            float(ratio[1]), K, PH, PW, rt.MLP_PASTE_NONE | self._tail_flags, c.view(self.det_i32),
            c.view(self.trim_counts), c.view(self.trim_m), ctypes.c_void_p(None), c.stream()))
        rt.check(lib.mlp_clip_masks(
            c.handle, c.view(self.det_i32), masks_ptr, r_cap, r_dev, self.C, c.view(self.trim_counts), B, K, mh, mw,
            PH, PW, c.view(self.clip_geom), c.view(self.clip_pool), cap, c.view(self.clip_used), c.stream()))
        self._compact_det = False
        return self.det_i32, self.clip_geom, self.clip_pool, self.clip_used

    def trim_and_summarize(self, rois, roi_masks, seg_outs, default_road_size=3.25, threshold=0.1,
                           paste=False, split=False):
        """SURVEY 8(f) rank 1 fused behind the tail: TrimInstances + UpSampleOutput, then
        SummaryOutput (road_project/setup/serving.py:45-48) evaluated straight from the mask tiles.
        With paste=False the [B,M,PH,PW] masks are never written (the serving graph only reduces
        them); paste=True also writes them in the configured output mode.  seg_outs: int32
        [B,PH,PW,S] (UpSampleOutput's thresholded semantic map).  Returns (det_i32 [B,K,6],
        summary flat with the valid prefix [B,M',11], m_out [1] = M' on the device)."""
        from .layers import summary as ls
        c, lib, B, K, L = self.ctx, self.lib, self.B, self.K, self.L
        if not self.cfg.fused:
            raise rt.InvalidArgumentError(rt.MLP_EINVAL, "trim_and_summarize needs the fused tail (fused=True)")
        st = c.stream()
        mh, mw = self.cfg.mask_size
        PH, PW = self.frame_hw
        if tuple(seg_outs.shape[:3]) != (B, PH, PW) or seg_outs.dtype != torch.int32:
            raise rt.InvalidArgumentError(rt.MLP_EINVAL, f"seg_outs must be int32 [{B},{PH},{PW},S]")
        S = int(seg_outs.shape[3])
        r_cap = L * K
        r_dev = ctypes.c_void_p(c.view(rois.level_m).value + 4 * L)
        masks_ptr = c.view(roi_masks, torch.float32)
        ratio = (torch.tensor([float(PH), float(PW)], dtype=torch.float32)
                 / torch.tensor([float(self.image_hw[0]), float(self.image_hw[1])], dtype=torch.float32))
        if not hasattr(self, "summary"):
            self.summary = c.empty((B * (K + 1) * 11,), torch.float32)
            self.summary_m = c.empty((1,), torch.int32)
            self.road_unit = c.empty((B, PH), torch.float32)
            self.road_bits = c.empty((B, PH, (PW + 31) // 32), torch.int32)
            self.crack_bits = c.empty((B, PH, (PW + 31) // 32), torch.int32)
            self.crack_box = c.empty((4 + 8 * B,), torch.int32)
        def tail():
            flags = 0
            if self._fill_done is not None:                  # a background fill of this batch is in flight: join it
                torch.cuda.current_stream(c.device).wait_event(self._fill_done)
                self._fill_done = None
                flags = rt.MLP_PASTE_PREFILLED if paste else 0
            rt.check(lib.mlp_trim_paste(
                c.handle, c.view(rois.roi_boxes), masks_ptr, B, r_cap, r_dev, mh, mw, self.C,
                float(ratio[0]), float(ratio[1]), K, PH, PW,
                (self.paste_mode if paste else rt.MLP_PASTE_NONE) | self._tail_flags | flags,
                c.view(self.det_i32), c.view(self.trim_counts), c.view(self.trim_m),
                c.view(self.pasted) if paste else ctypes.c_void_p(None), c.stream()))

        def road():                    # needs only the semantic map
            rt.check(lib.mlp_road_scan(
                c.handle, c.view(seg_outs), B, PH, PW, S, ls.ROAD_CHANNEL, ls.CRACK_CHANNEL,
                float(default_road_size), c.view(self.road_unit), c.view(self.road_bits), c.view(self.crack_bits),
                c.view(self.crack_box), c.stream()))

        def summarize():               # needs the tail and the road scan
            rt.check(lib.mlp_tile_summary(
                c.handle, c.view(self.det_i32), ctypes.c_void_p(None), masks_ptr, r_cap, r_dev, self.C,
                c.view(self.trim_counts), B, K, K, ctypes.c_void_p(None), mh, mw,
                c.view(self.road_unit), c.view(self.road_bits), c.view(self.crack_box),
                PH, PW, float(threshold), c.view(self.summary), c.view(self.summary_m), c.view(self.trim_m), c.stream()))

        self._compact_det = False
        if split:                      # the caller enqueues the parts itself (capture_serving: on two streams)
            return tail, road, summarize
        tail()
        road()
        summarize()
        self._compact_det = False
        return self.det_i32, self.summary, self.summary_m

    def draw(self, rois, roi_masks, images, instance_colors, instance_alpha=.3, seg_outs=None,
             semantic_colors=None, semantic_alpha=.3, boxes=False):
        """SURVEY 8(f) rank 2 behind the tail: the visualisation branch of
        road_project/setup/serving.py:34-40 - DrawBoxes (boxes=True), DrawInstance and, when seg_outs is
        given, DrawSegmentation - drawn straight from the mask tiles.  Needs the tail prepared by
        trim_and_paste / trim_and_summarize of the same batch.  images: uint8 or float32 [B,PH,PW,3].
        Returns uint8 [B,PH,PW,3]."""
        c, lib, B, K, L = self.ctx, self.lib, self.B, self.K, self.L
        if not self.cfg.fused:
            raise rt.InvalidArgumentError(rt.MLP_EINVAL, "draw needs the fused tail (fused=True)")
        mh, mw = self.cfg.mask_size
        PH, PW = self.frame_hw
        if tuple(images.shape) != (B, PH, PW, 3) or images.dtype not in (torch.uint8, torch.float32):
            raise rt.InvalidArgumentError(rt.MLP_EINVAL, f"images must be uint8/float32 [{B},{PH},{PW},3]")
        col = rt.DrawColorsC.make(instance_colors, instance_alpha)
        sem, seg_ptr, seg_t = None, ctypes.c_void_p(None), rt.MLP_I32
        if seg_outs is not None:
            sem = rt.DrawColorsC.make(semantic_colors, semantic_alpha)
            if tuple(seg_outs.shape) != (B, PH, PW, sem.num_classes) or seg_outs.dtype not in (torch.int32, torch.float32):
                raise rt.InvalidArgumentError(rt.MLP_EINVAL, "seg_outs must be int32/float32 [B,PH,PW,len(semantic_colors)]")
            seg_ptr = c.view(seg_outs)
            seg_t = rt.MLP_I32 if seg_outs.dtype == torch.int32 else rt.MLP_F32
        if not hasattr(self, "vis"):
            self.vis = c.empty((B, PH, PW, 3), torch.uint8)
        r_dev = ctypes.c_void_p(c.view(rois.level_m).value + 4 * L)
        # boxes=True: DrawBoxes' rectangles ride along as a one-bit-per-pixel map (no copy of the frame)
        fn = lib.mlp_draw_tiles_boxes if boxes else lib.mlp_draw_tiles
        rt.check(fn(
            c.handle, c.view(images), rt.MLP_U8 if images.dtype == torch.uint8 else rt.MLP_F32,
            c.view(self.det_i32), ctypes.c_void_p(None), c.view(roi_masks, torch.float32), L * K, r_dev, self.C,
            c.view(self.trim_counts), B, K, K, ctypes.c_void_p(None), mh, mw, PH, PW, ctypes.byref(col),
            seg_ptr, seg_t, ctypes.byref(sem) if sem is not None else None, c.view(self.vis), c.stream()))
        return self.vis

    def encode(self, images=None, quality=95):
        """Last step of the serving graph (road_project/setup/serving.py:41, EncodeImageContent): the JPEG files
        of `images` (default: the overlay of the last draw) as (files uint8 [B,stride], lengths int32 [B]) on
        the device - every frame of the batch, bytes identical to tf.io.encode_jpeg / libjpeg.  A negative
        length reports a file that did not fit the stride (3 bytes per pixel)."""
        c, B = self.ctx, self.B
        images = self.vis if images is None else images
        PH, PW = int(images.shape[1]), int(images.shape[2])
        if tuple(images.shape) != (B, PH, PW, 3) or images.dtype != torch.uint8:
            raise rt.InvalidArgumentError(rt.MLP_EINVAL, f"images must be uint8 [{B},H,W,3]")
        if getattr(self, "jpeg_hw", None) != (PH, PW):
            stride = min(int(self.lib.mlp_jpeg_max_bytes(PH, PW)), 1024 + 3 * PH * PW)
            self.jpeg_files = c.empty((B, (stride + 15) // 16 * 16), torch.uint8)
            self.jpeg_len = c.empty((B,), torch.int32)
            self.jpeg_hw = (PH, PW)
        rt.check(self.lib.mlp_jpeg_encode(c.handle, c.view(images), B, PH, PW, int(quality), c.view(self.jpeg_files),
                                          int(self.jpeg_files.shape[1]), c.view(self.jpeg_len), c.stream()))
        return self.jpeg_files, self.jpeg_len

    def summary_view(self):
        """Reference-shaped [B,M',11] view of the last trim_and_summarize (one D2H of M')."""
        Mo = int(self.summary_m.item())
        return self.summary[:self.B * Mo * 11].view(self.B, Mo, 11)

    # ---------------------------------------------------------------- CUDA graph
    def _own_context(self):
        """A captured graph bakes the ctx's scratch pointers in.  The per-device singleton ctx is shared
        with every stand-alone layer call (which may regrow an arena for a larger shape), so a pipeline
        that is about to be captured moves to a private ctx first; capture then freezes it."""
        if self.ctx.shared:
            self.ctx = rt.Context(self.ctx.device)

    def capture(self, loc_pred, cls_pred, fmaps, roi_masks, prefill=None):
        """Capture one whole batch (detect_and_align + trim_and_paste over THESE buffers) into a CUDA
        graph: the path is sync-free - every data-dependent size stays on the device - so the six
        kernels and their memsets replay as one launch.  Refill the same input tensors and call
        `.replay()` on the returned graph; results land in the pipeline's buffers as usual.  The
        mask head is not part of the graph: roi_masks must already hold its output for the RoIs of
        this batch when the graph is replayed (for a real model capture the two halves separately)."""
        self._own_context()
        torch.cuda.synchronize(self.ctx.device)
        side = torch.cuda.Stream(device=self.ctx.device)
        side.wait_stream(torch.cuda.current_stream(self.ctx.device))
        with torch.cuda.stream(side):                       # warm-up: scratch growth, function attributes
            for _ in range(2):
                rois = self.detect_and_align(loc_pred, cls_pred, fmaps, prefill=prefill)
                self.trim_and_paste(rois, roi_masks)
        torch.cuda.current_stream(self.ctx.device).wait_stream(side)
        torch.cuda.synchronize(self.ctx.device)
        self.ctx.freeze(True)                               # the graph references the arenas from here on
        graph = torch.cuda.CUDAGraph()
        # captured on a high-priority stream: kernel nodes inherit it, the forked background fill stays at the
        # least priority and only fills the SM slots the main chain leaves free
        with torch.cuda.graph(graph, stream=torch.cuda.Stream(device=self.ctx.device, priority=-1)):
            rois = self.detect_and_align(loc_pred, cls_pred, fmaps, prefill=prefill)
            self.trim_and_paste(rois, roi_masks)
        graph._mlp_owner = self                             # the graph replays into this pipeline's buffers and scratch
        return graph, rois

    def capture_serving(self, loc_pred, cls_pred, fmaps, roi_masks, seg_outs, images, instance_colors,
                        instance_alpha=.3, semantic_colors=None, semantic_alpha=.3, boxes=True, quality=95,
                        parallel_branches=True):
        """The serving tail of one batch as ONE CUDA graph: detect_and_align -> trim_and_summarize -> draw ->
        encode (SummaryOutput, the three overlays and the JPEG files; road_project/setup/serving.py:29-48).  All of
        it is sync-free, so the ~25 kernels replay as one launch; with parallel_branches the graph forks after the
        tail (road scan + summary beside overlays + JPEG).  Refill the same input tensors and call
        `.replay()`; summary, overlay and files land in the pipeline's buffers (summary_view(), vis, jpeg_files,
        jpeg_len).  As with capture(), the mask head is not part of the graph."""
        # the summary branch at the least priority, the graph itself captured on a high-priority stream (kernel nodes
        # inherit it): the overlay -> JPEG chain is the critical path, the summary fills what it leaves free
        prio = os.environ.get("MLP_SERVING_PRIORITIES", "1") != "0"
        branch = torch.cuda.Stream(device=self.ctx.device, priority=0) if prio else torch.cuda.Stream(device=self.ctx.device)

        def run():
            # two branches (different scratch arenas, no shared state), joined at the end: road scan from the start and
            # the summary after the tail on one, detection, tail, overlays and JPEG on the other
            cur = torch.cuda.current_stream(self.ctx.device)
            if parallel_branches:
                start = torch.cuda.Event()
                start.record(cur)
            rois = self.detect_and_align(loc_pred, cls_pred, fmaps, prefill=False)    # no masks are written
            tail, road, summarize = self.trim_and_summarize(rois, roi_masks, seg_outs, split=True)
            if parallel_branches:
                with torch.cuda.stream(branch):             # the road scan only reads the semantic map: it runs beside
                    branch.wait_event(start)                # the latency-bound NMS kernels
                    road()
            tail()
            if parallel_branches:
                fork = torch.cuda.Event()
                fork.record(cur)
                with torch.cuda.stream(branch):
                    branch.wait_event(fork)
                    summarize()
            else:
                road()
                summarize()
            self.draw(rois, roi_masks, images, instance_colors, instance_alpha, seg_outs=seg_outs,
                      semantic_colors=semantic_colors, semantic_alpha=semantic_alpha, boxes=boxes)
            self.encode(quality=quality)
            if parallel_branches:
                cur.wait_stream(branch)
            return rois
        self._own_context()
        torch.cuda.synchronize(self.ctx.device)
        side = torch.cuda.Stream(device=self.ctx.device)
        side.wait_stream(torch.cuda.current_stream(self.ctx.device))
        with torch.cuda.stream(side):                       # warm-up: buffers, scratch growth, function attributes
            for _ in range(2):
                run()
        torch.cuda.current_stream(self.ctx.device).wait_stream(side)
        torch.cuda.synchronize(self.ctx.device)
        self.ctx.freeze(True)                               # the graph references the arenas from here on
        graph = torch.cuda.CUDAGraph()
        cap = torch.cuda.Stream(device=self.ctx.device, priority=-1) if prio else None
        with torch.cuda.graph(graph, stream=cap):
            rois = run()
        graph._mlp_owner = self                             # the graph replays into this pipeline's buffers and scratch
        return graph, rois

    def result_views(self):
        """Reference-shaped views of the last trim_and_paste (one D2H of M)."""
        M = int(self.trim_m.item())
        PH, PW = self.frame_hw
        if getattr(self, "_compact_det", True):
            det = self.det_i32[:self.B * M * 6].view(self.B, M, 6)
        else:
            det = self.det_i32.view(self.B, self.K, 6)[:, :M]
        masks = self.pasted[:self.B * M * PH * self.paste_row].view(self.B, M, PH, self.paste_row)
        return det, masks


def expand_clipped(geom, pool, frame_hw, m_rows=None):
    """Host-side inverse of PostProcessPipeline.trim_and_clip: geom int32 [B,K,8] and the byte pool (NumPy arrays
    or CPU tensors, e.g. what a client received) -> the dense uint8 {0,1} masks [B,M,PH,PW] CropAndPadMask + `> 0.5`
    define (M = m_rows, default K).  Zero-fill, then OR every instance's bit rows into its clipped box."""
    import numpy as np
    geom = np.asarray(geom)
    pool = np.asarray(pool, dtype=np.uint8)
    B, K = geom.shape[:2]
    M = K if m_rows is None else int(m_rows)
    PH, PW = frame_hw
    out = np.zeros((B, M, PH, PW), dtype=np.uint8)
    for b in range(B):
        for j in range(M):
            xmin, ymin, w, h, lo, hi = (int(v) for v in geom[b, j, :6])
            if w <= 0 or h <= 0:
                continue
            off = (lo & 0xffffffff) | ((hi & 0xffffffff) << 32)
            rb = (w + 7) // 8
            rows = pool[off:off + h * rb].reshape(h, rb)
            out[b, j, ymin:ymin + h, xmin:xmin + w] = np.unpackbits(rows, axis=1, bitorder="little")[:, :w]
    return out
