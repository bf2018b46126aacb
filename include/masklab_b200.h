/* masklab_b200.h — C ABI of the B200-native MaskLab post-backbone hot path.
 *
 * One shared library (libmasklab_b200.so, hand-written CUDA for sm_100a) replaces
 * the TensorFlow op chain behind the reference's Keras custom layers
 * (/root/reference/engine/layers/{detection,instance,misc}.py).  The reference is
 * pure Python, so "the FFI a maintainer would bind" is ctypes; every entry point
 * below names the reference interface (file:line) it replaces.  INTEGRATION.md
 * shows the ctypes stubs.
 *
 * Conventions
 *   - plain C: device pointers, sizes, scalars; no torch / TF types.
 *   - every pointer named *_dev is DEVICE memory on the ctx's GPU, 16-byte aligned,
 *     dense row-major ("C contiguous"); tensors are BORROWED for the duration of
 *     the enqueued work, never freed by the library.
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); no hidden
 *     synchronisation except in functions documented as "[syncs]".
 *   - return 0 (MLP_OK) or a negative MLP_E* code; mlp_last_error() gives the
 *     thread-local message.
 *   - data-dependent shapes (M = kept boxes per image, Mf = RoIs per level) are
 *     produced ON THE DEVICE as int32 scalars and consumed by later stages from
 *     device memory; the host reads them only when it has to materialise the
 *     reference's dynamically shaped tensors.
 *   - a ctx owns scratch (candidate lists, NMS work space); one ctx per
 *     (GPU, stream); not thread-safe.
 *   - padding sentinel is -1 everywhere, at least one slot per image
 *     (engine/layers/misc.py:235-236, 275-286).
 */
#ifndef MASKLAB_B200_H_
#define MASKLAB_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MLP_VERSION 100          /* 0.1.0 */

#define MLP_OK          0
#define MLP_EINVAL     -1        /* bad argument (shape, alignment, limits)          */
#define MLP_ECUDA      -2        /* CUDA runtime error (message has the CUDA string) */
#define MLP_ENOMEM     -3        /* scratch allocation failed                        */
#define MLP_EDLPACK    -4        /* DLPack tensor rejected (dtype/device/strides)    */
#define MLP_EBATCH     -5        /* B > 32: tf.dynamic_partition(...,32), misc.py:275 */
#define MLP_EFROZEN    -6        /* scratch would have to grow while a CUDA graph references it */

#define MLP_MAX_LEVELS   8
#define MLP_MAX_ANCHORS  32
#define MLP_MAX_BATCH    32      /* engine/layers/misc.py:275                        */
#define MLP_MAX_KEEP     2048    /* largest nms_max_output_size the NMS kernels hold */

typedef struct mlp_ctx mlp_ctx;
typedef void* mlp_stream_t;      /* cudaStream_t */

/* ---- library / context ------------------------------------------------------ */
int         mlp_version(void);
const char* mlp_last_error(void);
int         mlp_ctx_create(int device, mlp_ctx** out);
void        mlp_ctx_destroy(mlp_ctx* ctx);
int         mlp_ctx_device(const mlp_ctx* ctx);
int         mlp_ctx_sm_count(const mlp_ctx* ctx);
/* Bytes of scratch currently owned by the ctx. */
int64_t     mlp_ctx_scratch_bytes(const mlp_ctx* ctx);
/* Number of kernels this ctx has launched since creation (bench `gpu_launches`). */
int64_t     mlp_ctx_launch_count(const mlp_ctx* ctx);
/* Scratch arenas are grow-only and regrowth frees the old block.  A captured CUDA graph keeps
 * pointers into them, so whoever captures calls mlp_ctx_freeze_scratch(ctx, 1) after the warm-up
 * run: from then on a call that would need more scratch returns MLP_EFROZEN (nothing is freed)
 * until the graph is dropped and the ctx thawed with freeze = 0.  Growth attempted while a
 * stream capture is active fails with MLP_ECUDA and a message that says so.                 */
int         mlp_ctx_freeze_scratch(mlp_ctx* ctx, int freeze);

/* ---- per-stage timing (bench.py roofline) -----------------------------------
 * When enabled, every stage function brackets its kernels with a pair of CUDA events on
 * the launching stream (no extra synchronisation).  mlp_ctx_profile_read [syncs] sums
 * the elapsed milliseconds and the number of bracketed calls per stage.               */
#define MLP_NUM_STAGES 24
const char* mlp_stage_name(int stage);
int mlp_ctx_profile_enable(mlp_ctx* ctx, int enable);
int mlp_ctx_profile_read(mlp_ctx* ctx, double* ms_out /*[MLP_NUM_STAGES]*/,
                         int64_t* count_out /*[MLP_NUM_STAGES]*/);

/* ---- DLPack handoff ---------------------------------------------------------
 * Python passes PyCapsule("dltensor") -> DLManagedTensor*; this validates it and
 * extracts the plain pointer/shape the stage functions take.  Layout of
 * DLManagedTensor follows dlpack.h (v0.x legacy struct, what
 * torch.utils.dlpack.to_dlpack produces).  The tensor stays owned by the caller.  */
typedef struct {
    void*   data;            /* device pointer incl. byte_offset */
    int32_t device_id;
    int32_t ndim;
    int32_t dtype_code;      /* 0 int, 1 uint, 2 float */
    int32_t dtype_bits;
    int64_t shape[8];
    int64_t numel;
} mlp_tensor_view;

#define MLP_F32 0
#define MLP_I32 1
#define MLP_U8  2
#define MLP_I64 3
/* expect_dtype: one of MLP_F32/I32/U8/I64, or -1 to accept any.  Rejects non-CUDA
 * devices, a device other than the ctx's, non-dense strides, ndim > 8 and
 * pointers not aligned to 16 bytes. */
int mlp_dlpack_view(const mlp_ctx* ctx, const void* dl_managed_tensor, int expect_dtype,
                    mlp_tensor_view* out);

/* ---- a1/a2: prior (anchor) configuration ------------------------------------
 * Output of PriorBoxes.setup (engine/prior.py:55-67) grouped by stride ascending
 * as PriorLayer.__init__ does (engine/layers/detection.py:260-262).               */
typedef struct {
    int32_t num_levels;
    int32_t padding_same;                               /* 1: ceil(H/s) ('same'); 0: floor */
    int32_t stride[MLP_MAX_LEVELS];
    int32_t num_anchors[MLP_MAX_LEVELS];
    int32_t anchor_w[MLP_MAX_LEVELS][MLP_MAX_ANCHORS];
    int32_t anchor_h[MLP_MAX_LEVELS][MLP_MAX_ANCHORS];
} mlp_prior_config;

/* N = sum_l Hf_l*Wf_l*A_l for an H x W image; negative on bad config. */
int64_t mlp_prior_count(const mlp_prior_config* prior, int height, int width);

/* PriorLayer.call (engine/layers/detection.py:269-298): out_dev int32 [B,N,4]
 * (cx,cy,w,h), anchor order (stride asc, y, x, anchor), tiled over the batch.     */
int mlp_prior_layer(mlp_ctx* ctx, const mlp_prior_config* prior, int batch, int height,
                    int width, int32_t* out_dev, mlp_stream_t stream);

/* ---- a3: RestoreBoxes.call (engine/layers/detection.py:325-344) --------------
 * loc_dev f32 [rows,4], prior_dev [rows,4] int32 (prior_is_f32=0) or f32 (=1)
 * -> out_dev f32 [rows,4] (cx,cy,w,h).  exp() is the correctly rounded f32 value. */
int mlp_restore_boxes(mlp_ctx* ctx, const float* loc_dev, const void* prior_dev,
                      int prior_is_f32, int64_t rows, float* out_dev, mlp_stream_t stream);
/* Same, anchors generated in-kernel from the prior config (no [B,N,4] prior tensor). */
int mlp_restore_boxes_from_prior(mlp_ctx* ctx, const mlp_prior_config* prior, const float* loc_dev,
                                 int batch, int height, int width, float* out_dev,
                                 mlp_stream_t stream);

/* ---- a4: NormalizeBoxes.call (engine/layers/detection.py:360-375) ------------
 * boxes_dev f32 [rows,row_stride] (first 4 columns cx,cy,w,h) -> out_dev f32
 * [rows,4] (y1,x1,y2,x2)/(H,W).                                                   */
int mlp_normalize_boxes(mlp_ctx* ctx, const float* boxes_dev, int64_t rows, int row_stride,
                        float image_h, float image_w, float* out_dev, mlp_stream_t stream);

/* ---- a5-a8: DetectionProposal.call (engine/layers/detection.py:482-567) ------ */
typedef struct {
    float   min_confidence;        /* cls_pred >= min_confidence                     */
    float   nms_iou_threshold;     /* per (image,class) NMS: suppress iff IoU > thr  */
    float   post_iou_threshold;    /* per image cross-class NMS                      */
    int32_t nms_max_output_size;   /* cap on kept boxes per NMS call, <= MLP_MAX_KEEP */
    int32_t strict_batch;          /* 1: reject B > 32 like misc.py:275              */
} mlp_detection_params;

/* cls_dev f32 [B,N,C]; boxes_dev f32 [B,N,4] (cx,cy,w,h) as RestoreBoxes returns.
 * Outputs (capacity K = nms_max_output_size rows per image):
 *   det_dev    f32 [B,K,6]  (cx,cy,w,h,class,score), rows >= count[b] are -1
 *   keep_dev   i32 [B,K,2]  (anchor index n, class c) of each kept row, -1 padded
 *                           (may be NULL)
 *   counts_dev i32 [B]      kept boxes per image
 *   m_dev      i32 [1]      M = max(1, max_b counts[b])  -> reference output is
 *                           det[:, :M, :] (MoldBatch, engine/layers/misc.py:231-286) */
int mlp_detection_proposal(mlp_ctx* ctx, const float* cls_dev, const float* boxes_dev,
                           int batch, int64_t num_boxes, int num_classes,
                           const mlp_detection_params* params, float* det_dev,
                           int32_t* keep_dev, int32_t* counts_dev, int32_t* m_dev,
                           mlp_stream_t stream);

/* Fused a2+a3+a4+a5-a8: boxes are decoded from loc_dev [B,N,4] and the prior
 * config only for candidates that pass the score threshold; no [B,N,4] prior or
 * restored-box tensor is materialised.  Same outputs as mlp_detection_proposal.    */
int mlp_detect_from_heads(mlp_ctx* ctx, const mlp_prior_config* prior, const float* loc_dev,
                          const float* cls_dev, int batch, int height, int width,
                          int num_classes, const mlp_detection_params* params,
                          float* det_dev, int32_t* keep_dev, int32_t* counts_dev,
                          int32_t* m_dev, mlp_stream_t stream);

/* ---- a9: MaskDistribute.call (engine/layers/instance.py:52-66) ---------------
 * det_dev f32 [rows,6] -> out_dev f32 [rows,7] (k,cx,cy,w,h,class,score).         */
int mlp_mask_distribute(mlp_ctx* ctx, const float* det_dev, int64_t rows, int max_k,
                        float base_size, float* out_dev, mlp_stream_t stream);

/* ---- a10: PyramidRoiAlign.call (engine/layers/instance.py:109-139) -----------
 * Plan: per level f and image b count rows of dist_dev [B,M,7] with k == f;
 *   level_counts_dev i32 [L,B];  level_m_dev i32 [L+1]: [f] = Mf = max(1, max_b count),
 *   [L] = sum_f Mf (R, written by mlp_roi_align_run; TrimInstances reads it as r_dev).
 * dist row stride is m_stride rows per image (the capacity K when dist_dev is a
 * capacity buffer); only the first *m_dev rows (or m_rows when m_dev is NULL) of
 * each image are looked at.                                                       */
int mlp_roi_align_plan(mlp_ctx* ctx, const float* dist_dev, int batch, int m_rows, int m_stride,
                       const int32_t* m_dev, int num_levels, int32_t* level_counts_dev,
                       int32_t* level_m_dev, mlp_stream_t stream);

/* Run: fmaps_dev[f] f32 [B,fh[f],fw[f],Cf] (NHWC).  For each level f writes
 *   crops_dev[f]   f32 [B,Mf,ch,cw,Cf] with Mf = level_m_dev[f] read ON DEVICE
 *                  (caller allocates capacity >= B*m_rows slots, or exactly B*Mf
 *                  after reading level_m_dev), -1 in padded slots;
 *   roi_boxes_dev  f32 [B,sum_f Mf,6] (cx,cy,w,h,class,score), level-major, -1 pad.
 * crop_and_resize arithmetic of TF (bilinear, extrapolation 0), boxes normalised
 * by the model-input image size (image_h,image_w) at every level.                  */
int mlp_roi_align_run(mlp_ctx* ctx, const float* const* fmaps_dev, const int32_t* fh,
                      const int32_t* fw, int num_levels, int channels, const float* dist_dev,
                      int batch, int m_rows, int m_stride, const int32_t* m_dev,
                      float image_h, float image_w, int crop_h, int crop_w,
                      const int32_t* level_counts_dev, int32_t* level_m_dev,
                      float* const* crops_dev, float* roi_boxes_dev, mlp_stream_t stream);

/* ---- a11: TrimInstances.call (engine/layers/instance.py:258-277) -------------
 * roi_boxes_dev f32 [B,R,6], roi_masks_dev f32 [B,R,mh,mw,C].  R is read from
 * r_dev (i32 [1], e.g. the sum of level_m_dev) when not NULL, else r_rows.
 * Plan writes counts_dev i32 [B] (rows with class != -1) and m_dev i32 [1].
 * Run writes out_boxes_dev f32 [B,M,6] and out_masks_dev f32 [B,M,mh,mw] (the class
 * channel of each valid row), M read from m_dev on device, -1 padded.             */
int mlp_trim_plan(mlp_ctx* ctx, const float* roi_boxes_dev, int batch, int r_rows,
                  const int32_t* r_dev, int32_t* counts_dev, int32_t* m_dev, mlp_stream_t stream);
int mlp_trim_run(mlp_ctx* ctx, const float* roi_boxes_dev, const float* roi_masks_dev, int batch,
                 int r_rows, const int32_t* r_dev, int mask_h, int mask_w, int num_classes,
                 const int32_t* m_dev, float* out_boxes_dev, float* out_masks_dev,
                 mlp_stream_t stream);

/* ---- a12: UpSampleOutput.call, instance part (engine/layers/misc.py:169-188) --
 * det_dev f32 [rows,6] -> det_i32_dev [rows,6]; masks_dev f32 [mask_elems] ->
 * masks_i32_dev (mask > 0.5).  ratio_h = PH/h_s, ratio_w = PW/w_s; cx,w use
 * ratio_h and cy,h use ratio_w exactly as misc.py:180-183 does.                   */
int mlp_upsample_output(mlp_ctx* ctx, const float* det_dev, int64_t rows, float ratio_h,
                        float ratio_w, int32_t* det_i32_dev, const float* masks_dev,
                        int64_t mask_elems, int32_t* masks_i32_dev, mlp_stream_t stream);

/* ---- a13/a14: CropAndPadMask.call (engine/layers/misc.py:358-401) ------------ */
#define MLP_PASTE_F32 0   /* drop-in: f32 bilinear values, [B,M,PH,PW]                  */
#define MLP_PASTE_U8  1   /* canonical binary mask (value > 0.5, misc.py:457,:611-615) */
#define MLP_PASTE_BITS 2  /* the same binary mask, 1 bit per pixel: [B,M,PH,PW/8] uint8, bit k of byte i
                             = pixel 8*i+k (numpy packbits bitorder='little'); PW % 8 == 0       */
#define MLP_PASTE_NONE 3  /* mlp_trim_paste only: prepare the tail (int boxes, tile table), write no masks;
                           * out_dev may be NULL and M is left to mlp_tile_summary                    */
/* flags OR-ed into mlp_trim_paste's out_mode */
#define MLP_PASTE_PREFILLED 0x100 /* out_dev's [B,M,PH,PW] prefix was zeroed by mlp_paste_prefill: write the boxes only */
#define MLP_MASKS_PLANAR    0x200 /* roi_masks_dev is [B,R,C,mh,mw] (one plane per class, what a channels-first mask
                                   * head produces) instead of the reference's [B,R,mh,mw,C]: the tail then reads
                                   * only the instance's own class plane (1/C of the bytes)                        */
/* det_i32_dev [B,M,6], masks_i32_dev i32 [B,M,mh,mw] -> out_dev [B,M,PH,PW].
 * M is read from m_dev (i32 [1]) when not NULL, else m_rows; m_stride is the row
 * stride of det/masks per image.  A box clipped to zero area gives an all-zero mask
 * (TF raises InvalidArgument there).                                              */
int mlp_crop_and_pad_mask(mlp_ctx* ctx, const int32_t* det_i32_dev, const int32_t* masks_i32_dev,
                          int batch, int m_rows, int m_stride, const int32_t* m_dev, int mask_h,
                          int mask_w, int frame_h, int frame_w, int out_mode, void* out_dev,
                          mlp_stream_t stream);

/* ---- fused halves of the path (what PostProcessPipeline enqueues) -------------------------
 * The mask head (MaskSubNet, dense convolutions) sits between them and is not part of this
 * library.  Same results as the chain of stage functions above, fewer kernels and no
 * intermediate tensors:
 *
 * mlp_detect_align = a2-a10, the wiring of engine/retinamasklab.py:458-469: boxes decoded from
 *   loc_dev + the prior config for candidates only; MaskDistribute and the RoIAlign plan are folded
 *   into the cross-class NMS epilogue; then the RoIAlign run.  4 kernel launches.
 *   Outputs are capacity buffers (K = nms_max_output_size rows per image): det_dev [B,K,6],
 *   keep_dev [B,K,2] (may be NULL), counts_dev [B], m_dev [1], dist_dev [B,K,7],
 *   level_counts_dev [L,B], level_m_dev [L+1], crops_dev[f] (prefix [B,Mf,ch,cw,Cf] of a
 *   B*K*ch*cw*Cf buffer), roi_boxes_dev (prefix [B,R,6] of a B*L*K*6 buffer); L = max_k+1.
 *
 * mlp_trim_paste = a11-a14, engine/retinamasklab.py:615-616, :635-636 and
 *   road_project/setup/serving.py:30: valid rows of roi_boxes_dev [B,R,6] ranked per image, int32
 *   boxes of UpSampleOutput written to det_i32_dev [B,K,6] (capacity rows, padding rows as the
 *   reference produces them), and the paste kernel reads every instance's class channel of
 *   roi_masks_dev [B,R,mh,mw,C] directly (> 0.5 on the fly).  R is read from r_dev (i32 [1], e.g.
 *   level_m_dev + L) when not NULL, else r_rows.  counts_dev [B], m_dev [1] = M;
 *   out_dev [B,M,PH,PW] (uint8 or f32, M on device).  2 kernel launches.                     */
int mlp_detect_align(mlp_ctx* ctx, const mlp_prior_config* prior, const float* loc_dev,
                     const float* cls_dev, int batch, int height, int width, int num_classes,
                     const mlp_detection_params* params, int max_k, float base_size,
                     const float* const* fmaps_dev, const int32_t* fh, const int32_t* fw, int channels,
                     int crop_h, int crop_w, float* det_dev, int32_t* keep_dev, int32_t* counts_dev,
                     int32_t* m_dev, float* dist_dev, int32_t* level_counts_dev, int32_t* level_m_dev,
                     float* const* crops_dev, float* roi_boxes_dev, mlp_stream_t stream);
int mlp_trim_paste(mlp_ctx* ctx, const float* roi_boxes_dev, const float* roi_masks_dev, int batch,
                   int r_rows, const int32_t* r_dev, int mask_h, int mask_w, int num_classes,
                   float ratio_h, float ratio_w, int k_rows, int frame_h, int frame_w, int out_mode,
                   int32_t* det_i32_dev, int32_t* counts_dev, int32_t* m_dev, void* out_dev,
                   mlp_stream_t stream);
/* mlp_detect_plan = mlp_detect_align without its last kernel (the RoIAlign run): a2-a9 and the RoIAlign
 *   plan.  It exists so that the caller can fork after the NMS kernels: mlp_paste_prefill (below) on a
 *   second stream beside mlp_roi_align_run / the mask head on the first.
 * mlp_paste_prefill: CropAndPadMask's background ahead of time.  97 % of the [B,M,PH,PW] output is
 *   zeros that depend on M only (engine/layers/misc.py:393-399 pads every resized mask to the frame);
 *   this zeroes instance rows [m_from, m_to) of out_dev's flat prefix: m_to = max(1, max_b counts_dev[b])
 *   (misc.py:235-236) when counts_dev is given, else *m_to_dev; m_from = *m_from_dev (0 when NULL); all
 *   read on the device.  Zeroing a longer prefix than the final M needs is wasted work, never wrong, so a
 *   first call may run BEFORE the NMS kernels with m_to_dev = the M of the previous batch and a second one
 *   after them completes [that M, this batch's M).  Join the stream before
 *   mlp_trim_paste(.., out_mode | MLP_PASTE_PREFILLED, ..), which then writes only the 16-byte segments
 *   that intersect a box.  frame_w must be a multiple of the store width (16 uint8 / 4 float32 / 128
 *   bit-packed pixels).                                                                          */
int mlp_detect_plan(mlp_ctx* ctx, const mlp_prior_config* prior, const float* loc_dev,
                    const float* cls_dev, int batch, int height, int width, int num_classes,
                    const mlp_detection_params* params, int max_k, float base_size, float* det_dev,
                    int32_t* keep_dev, int32_t* counts_dev, int32_t* m_dev, float* dist_dev,
                    int32_t* level_counts_dev, int32_t* level_m_dev, mlp_stream_t stream);
int mlp_paste_prefill(mlp_ctx* ctx, const int32_t* m_from_dev, const int32_t* counts_dev,
                      const int32_t* m_to_dev, int batch, int k_rows, int frame_h, int frame_w, int out_mode,
                      void* out_dev, mlp_stream_t stream);

/* ---- a13/a14 in box-clipped form: CropAndPadMask (engine/layers/misc.py:358-401) + the consumers' > 0.5 without
 * the tf.pad zeros (misc.py:393-394), for clients on the far side of PCIe / NCCL.  Runs behind
 * mlp_trim_paste(.., MLP_PASTE_NONE, ..) of the same batch.  geom_dev int32 [B,K,8]: per slot of the capacity grid
 * { xmin, ymin, w, h, offset_lo, offset_hi, class, conf }; pool_dev: h rows of ceil(w/8) bytes per instance at
 * `offset`, bit k of byte i of row r = pixel (ymin+r, xmin+8i+k) of the instance's frame-sized binary mask
 * (w = h = 0: nothing pasted).  used_dev int64 [2] = bytes the batch needs (> pool_capacity: the pool was too small,
 * rows past it were dropped) and the number of work items.  mlp_clip_pool_bound: a capacity that cannot overflow. */
int64_t mlp_clip_pool_bound(int batch, int k_rows, int frame_h, int frame_w);
int mlp_clip_masks(mlp_ctx* ctx, const int32_t* det_i32_dev, const float* roi_masks_dev, int r_rows,
                   const int32_t* r_dev, int num_classes, const int32_t* counts_dev, int batch, int k_rows,
                   int mask_h, int mask_w, int frame_h, int frame_w, int32_t* geom_dev, uint8_t* pool_dev,
                   int64_t pool_capacity, int64_t* used_dev, mlp_stream_t stream);

/* ---- a8: MoldBatch.call (engine/layers/misc.py:231-286) as a standalone operator ----
 * x_dev [K,row_elems] of 4-byte elements, batch_idx_dev i32 [K] (image id of each row).
 * Plan: counts_dev i32 [B], m_dev i32 [1] = max(1, max_b count).  Run: out_dev
 * [B,M,row_elems] (M read on device), input order kept per image, padded with -1
 * (-1.0f when pad_is_float, else int -1).  Rows whose id is outside [0,B) are dropped
 * (tf.dynamic_partition raises there).                                              */
int mlp_mold_batch_plan(mlp_ctx* ctx, const int32_t* batch_idx_dev, int64_t rows, int batch,
                        int32_t* counts_dev, int32_t* m_dev, mlp_stream_t stream);
int mlp_mold_batch_run(mlp_ctx* ctx, const void* x_dev, const int32_t* counts_dev, int64_t rows,
                       int64_t row_elems, int batch, int pad_is_float, const int32_t* m_dev,
                       void* out_dev, mlp_stream_t stream);

/* ---- the first consumer of the pasted masks ---------------------------
 * SummaryOutput.call (engine/layers/misc.py:554-583) with CrackToInstance (:515-537),
 * CalculateInstanceSize (:633-724) and IncludeMyRoad (:603-617); wiring
 * road_project/setup/serving.py:45-48.  Floating-point contract (oracle/summary_oracle.py): sums
 * are accumulated in float64 and rounded to float32 once; the road-border fit is the closed-form
 * least squares in float64.  The semantic map is UpSampleOutput's {0,1} int32 tensor.
 *
 * mlp_road_scan: one pass over seg_dev i32 [B,PH,PW,S]:
 *   unit_dev f32 [B,PH] = metres per pixel on each frame row
 *   (CalculateInstanceSize._calculate_road_size_by_vertical_per_batch, misc.py:660-678),
 *   road_bits_dev u32 [B,PH,ceil(PW/32)] = my_road bitmap (bit i of word k = pixel 32k+i), and, when
 *   crack_channel >= 0, crack_bits_dev (same layout) and crack_box_dev i32 [4 + 8*B] (16-byte
 *   aligned): words 0-3 = (ymin,xmin,ymax,xmax) of the non-zero crack pixels of the whole batch
 *   ((INT_MAX,INT_MAX,-1,-1) if none) for CrackToInstance, then 8 words per image with the
 *   reductions of its crack pseudo-instance (opaque; consumed by the two functions below).
 *   2 kernel launches, 3 with a crack channel.
 *
 * mlp_summary_output: det_i32_dev [B,m_stride,6] + masks_dev [B,M,PH,PW] (MLP_F32 as CropAndPadMask
 *   returns them, or MLP_U8 binary) -> out_dev f32 [B,M',11] =
 *   (class,cx,cy,w,h,conf,pixel_counts,instance_size,horizontal_size,vertical_size,include_my_road).
 *   M from m_dev (i32 [1]) when not NULL, else m_rows; m_stride 0 = compact.  With crack_box_dev
 *   (as filled by mlp_road_scan) not NULL the crack pseudo-instance (mask = the crack bitmap) is
 *   appended when its box has positive area: M' = M + 1, written to m_out_dev; out_dev needs
 *   B*(m_rows+1)*11 floats.
 *
 * mlp_tile_summary: the same [B,M',11] summary WITHOUT the [B,M,PH,PW] tensor - every instance's
 *   float32 paste values are evaluated inside its clipped box straight from its 28x28 tile (exactly
 *   the values mlp_crop_and_pad_mask(MLP_PASTE_F32) would write) and reduced on the fly, so the
 *   largest tensor of the path is neither written nor re-read.  Two sources:
 *   masks_i32_dev != NULL : int32 tiles [B,m_stride,mh,mw] + det_i32_dev [B,m_stride,6], M from
 *                           m_dev / m_rows (the standalone CropAndPadMask inputs);
 *   masks_i32_dev == NULL : the fused tail - call mlp_trim_paste (any out_mode, MLP_PASTE_NONE to skip
 *                           the masks altogether) with the same ctx, shapes and stream first; tiles come
 *                           from roi_masks_dev [B,R,mh,mw,C] through the tail's slot table, m_rows =
 *                           m_stride = k_rows, counts_dev from that call; M is written to m_dev_out.  */
int mlp_road_scan(mlp_ctx* ctx, const int32_t* seg_dev, int batch, int frame_h, int frame_w,
                  int channels, int road_channel, int crack_channel, float default_road_size,
                  float* unit_dev, uint32_t* road_bits_dev, uint32_t* crack_bits_dev,
                  int32_t* crack_box_dev, mlp_stream_t stream);
int mlp_summary_output(mlp_ctx* ctx, const int32_t* det_i32_dev, const void* masks_dev, int mask_dtype,
                       const float* unit_dev, const uint32_t* road_bits_dev,
                       const int32_t* crack_box_dev, int batch, int m_rows, int m_stride,
                       const int32_t* m_dev, int frame_h, int frame_w, float include_threshold,
                       float* out_dev, int32_t* m_out_dev, mlp_stream_t stream);
int mlp_tile_summary(mlp_ctx* ctx, const int32_t* det_i32_dev, const int32_t* masks_i32_dev,
                     const float* roi_masks_dev, int r_rows, const int32_t* r_dev, int num_classes,
                     const int32_t* counts_dev, int batch, int m_rows, int m_stride, const int32_t* m_dev,
                     int mask_h, int mask_w, const float* unit_dev, const uint32_t* road_bits_dev,
                     const int32_t* crack_box_dev, int frame_h, int frame_w,
                     float include_threshold, float* out_dev, int32_t* m_out_dev, int32_t* m_dev_out,
                     mlp_stream_t stream);

/* ---- SURVEY 8(f) rank 3 (resize part): the bilinear resizes either side of the path -----------
 * tf.compat.v1.image.resize_bilinear(align_corners=True) of DownSampleInput.call
 * (engine/layers/misc.py:143-154) and of the semantic half of UpSampleOutput.call (:190-195).
 * Also ResizeLike.call (misc.py:302-307).  in_dev [B,in_h,in_w,S] (MLP_F32, MLP_U8 or MLP_I32, cast to
 * float32) -> out_dev [B,out_h,out_w,S]: float32 values, or with MLP_RESIZE_THRESHOLD int32
 * (value > 0.5) (misc.py:194).  MLP_RESIZE_NO_ALIGN_CORNERS selects the legacy align_corners=False
 * scale in/out (ResizeLike(align_corners=False)).                                               */
#define MLP_RESIZE_THRESHOLD        1
#define MLP_RESIZE_NO_ALIGN_CORNERS 2
int mlp_resize_bilinear(mlp_ctx* ctx, const void* in_dev, int in_dtype, int batch, int in_h, int in_w,
                        int channels, int out_h, int out_w, int flags, void* out_dev, mlp_stream_t stream);
/* SemanticSmoothing.call (engine/layers/semantic.py:270-284): tf.nn.erosion2d then tf.nn.dilation2d
 * with a flat kernel_size x kernel_size kernel, strides/rates 1, padding SAME, times weight;
 * kernel_size <= 0: only the weight.  in_dev/out_dev f32 [B,h,w,S]; in_dev != out_dev.          */
int mlp_semantic_smoothing(mlp_ctx* ctx, const float* in_dev, int batch, int height, int width, int channels,
                           int kernel_size, float weight, float* out_dev, mlp_stream_t stream);

/* ---- the overlay layers of the serving graph ------------------------------
 * DrawSegmentation.call (engine/layers/misc.py:412-421) and DrawInstance.call (:440-463), wired in
 * road_project/setup/serving.py:34-40.  images_dev [B,PH,PW,3] (MLP_U8 or MLP_F32), out_dev uint8
 * [B,PH,PW,3] = cast(clip(images + (sum_c colors[c] * seg[..., c]) * alpha, 0, 255)).
 *
 * mlp_draw_segmentation: seg_dev [B,PH,PW,C] (MLP_I32 or MLP_F32), C = colors->num_classes.
 * mlp_draw_instance: seg[..., c] = (sum of masks_dev[b, j] over the instances with det class == c)
 *   > 0.5, masks_dev [B,M,PH,PW] (MLP_F32 as CropAndPadMask returns them, or MLP_U8 binary), summed
 *   in instance order; M from m_dev / m_rows, m_stride 0 = compact.
 * mlp_draw_tiles: the same image WITHOUT the [B,M,PH,PW] tensor: every instance's float32 paste
 *   values are evaluated from its tile only where a pixel block touches its clipped box.  Tile
 *   sources as for mlp_tile_summary (int32 tiles, or the fused tail after mlp_trim_paste); mask rows
 *   of at most 32 columns.  With seg_dev / sem_colors not NULL, DrawSegmentation over the result
 *   (serving.py:38-40) is applied in the same pass.                                             */
/* mlp_draw_boxes: DrawBoxes.call (engine/layers/misc.py:481-503, serving.py:34): out_dev uint8
 *   [B,PH,PW,3] = the frame (float32 frames clipped to [0,255] and truncated) with the one-pixel white
 *   rectangle of tf.image.draw_bounding_boxes for every row of det_i32_dev [B,m_stride,6] (negative
 *   coordinates clamped to 0 first, so padding rows mark the pixel at the origin like the reference).
 *   images_dev == out_dev (uint8) draws in place.                                               */
int mlp_draw_boxes(mlp_ctx* ctx, const void* images_dev, int image_dtype, const int32_t* det_i32_dev, int batch,
                   int m_rows, int m_stride, const int32_t* m_dev, int frame_h, int frame_w, uint8_t* out_dev,
                   mlp_stream_t stream);
#define MLP_MAX_DRAW_CLASSES 16
typedef struct {
    int32_t num_classes;
    float   alpha;
    float   rgb[MLP_MAX_DRAW_CLASSES][3];
} mlp_draw_colors;
int mlp_draw_segmentation(mlp_ctx* ctx, const void* images_dev, int image_dtype, const void* seg_dev,
                          int seg_dtype, int batch, int frame_h, int frame_w, const mlp_draw_colors* colors,
                          uint8_t* out_dev, mlp_stream_t stream);
int mlp_draw_instance(mlp_ctx* ctx, const void* images_dev, int image_dtype, const int32_t* det_i32_dev,
                      const void* masks_dev, int mask_dtype, int batch, int m_rows, int m_stride,
                      const int32_t* m_dev, int frame_h, int frame_w, const mlp_draw_colors* colors,
                      uint8_t* out_dev, mlp_stream_t stream);
int mlp_draw_tiles(mlp_ctx* ctx, const void* images_dev, int image_dtype, const int32_t* det_i32_dev,
                   const int32_t* masks_i32_dev, const float* roi_masks_dev, int r_rows, const int32_t* r_dev,
                   int num_classes, const int32_t* counts_dev, int batch, int m_rows, int m_stride,
                   const int32_t* m_dev, int mask_h, int mask_w, int frame_h, int frame_w,
                   const mlp_draw_colors* inst_colors, const void* seg_dev, int seg_dtype,
                   const mlp_draw_colors* sem_colors, uint8_t* out_dev, mlp_stream_t stream);
/* mlp_draw_tiles_boxes: mlp_draw_boxes followed by mlp_draw_tiles in one pass over the frame (the whole
 *   visualisation branch of serving.py:34-40): the rectangles go into a one-bit-per-pixel map first, so the
 *   frame is neither copied nor written before the overlay pass.  Same arguments and result as drawing the
 *   boxes into a copy of the frame and handing that to mlp_draw_tiles.                           */
int mlp_draw_tiles_boxes(mlp_ctx* ctx, const void* images_dev, int image_dtype, const int32_t* det_i32_dev,
                   const int32_t* masks_i32_dev, const float* roi_masks_dev, int r_rows, const int32_t* r_dev,
                   int num_classes, const int32_t* counts_dev, int batch, int m_rows, int m_stride,
                   const int32_t* m_dev, int mask_h, int mask_w, int frame_h, int frame_w,
                   const mlp_draw_colors* inst_colors, const void* seg_dev, int seg_dtype,
                   const mlp_draw_colors* sem_colors, uint8_t* out_dev, mlp_stream_t stream);

/* ---- training-side target assignment -------------------------------------
 * CalculateIOU.call (engine/layers/detection.py:391-422): aa_dev [Na, aa_stride >= 4], bb_dev
 *   [Nb, bb_stride >= 4] (cx,cy,w,h) -> out_dev f32 [Na,Nb] = inter / (union + 1e-5).
 * AssignBoxes.call (detection.py:619-690): gt_boxes_dev f32 [B,G,6] (-1 padded), pr_boxes_dev
 *   [B,N,4] (MLP_F32 or MLP_I32 as PriorLayer returns them) -> cls_true_dev [B,N,C] one-hot,
 *   loc_true_dev [B,N,4], assign_mask_dev [B,N,1] (1 background, 0 assigned, -1 ignore band
 *   0.4 <= IoU < 0.5).  Repeated scatter indices behave like TensorFlow's CPU kernels: the last class
 *   update wins, regression targets of repeated matches add up.
 * AssignMasks.call (engine/layers/instance.py:330-380): roi_boxes_dev [B,R,6], gt_boxes_dev [B,G,6],
 *   gt_masks_dev f32 [B,G,H,W] -> out_dev i32 [B,R,mh,mw]: class id where the best ground truth's mask
 *   cropped to the RoI is > 0.5, num_classes elsewhere.
 * DetectionIOUMetric.call (engine/metrics.py:117-160): pred_boxes_dev [B,P,6], gt_boxes_dev [B,G,6]
 *   -> out_dev f32 [B,3] = (precision, recall, fmeasure) at IoU > 0.5.                          */
int mlp_calculate_iou(mlp_ctx* ctx, const float* aa_dev, int num_aa, int aa_stride, const float* bb_dev, int num_bb,
                      int bb_stride, float* out_dev, mlp_stream_t stream);
int mlp_assign_boxes(mlp_ctx* ctx, const float* gt_boxes_dev, const void* pr_boxes_dev, int pr_dtype, int batch,
                     int num_gt, int num_priors, int num_classes, float* cls_true_dev, float* loc_true_dev,
                     float* assign_mask_dev, mlp_stream_t stream);
int mlp_assign_masks(mlp_ctx* ctx, const float* roi_boxes_dev, int num_rois, const float* gt_boxes_dev, int num_gt,
                     const float* gt_masks_dev, int batch, int height, int width, int mask_h, int mask_w,
                     int num_classes, float match_iou_threshold, int32_t* out_dev, mlp_stream_t stream);
int mlp_detection_iou_metric(mlp_ctx* ctx, const float* pred_boxes_dev, int num_pred, const float* gt_boxes_dev,
                             int num_gt, int batch, float* out_dev, mlp_stream_t stream);

/* ---- SURVEY 8(f) rank 2, last step of the serving graph: the JPEG encode -----------------------
 * EncodeImageContent.call (engine/layers/misc.py:343-351, road_project/setup/serving.py:41) =
 * tf.io.encode_jpeg(image) with default attributes: libjpeg baseline, quality 95, YCbCr 4:2:0,
 * JDCT_ISLOW, Annex K Huffman tables, JFIF 300x300 dpi.  images_dev uint8 [B,H,W,3] ->
 * out_dev[b * out_stride ...] = the complete JPEG file of frame b, len_dev[b] = its size in bytes.
 * The bytes equal libjpeg(-turbo)'s for the same parameters.  A frame whose file would exceed
 * out_stride writes nothing and reports -(bytes needed); mlp_jpeg_max_bytes is the bound that can
 * never be exceeded (about 10 bytes per pixel; typical files take 0.3-1).  The reference encodes
 * frame 0 only; batch > 1 encodes every frame.
 * mlp_jpeg_header: the 623 header bytes (SOI .. SOS) for a frame size, into host memory.        */
int64_t mlp_jpeg_max_bytes(int frame_h, int frame_w);
int mlp_jpeg_header(int frame_h, int frame_w, int quality, uint8_t* out_host, int capacity);
int mlp_jpeg_encode(mlp_ctx* ctx, const uint8_t* images_dev, int batch, int frame_h, int frame_w, int quality,
                    uint8_t* out_dev, int64_t out_stride, int32_t* len_dev, mlp_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MASKLAB_B200_H_ */
