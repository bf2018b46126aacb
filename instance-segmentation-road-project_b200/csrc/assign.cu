// assign.cu — training-side target assignment: CalculateIOU, AssignBoxes,
// AssignMasks and DetectionIOUMetric.
//
// Reference: /root/reference/engine/layers/detection.py:378-422 (CalculateIOU), :589-697 (AssignBoxes),
// /root/reference/engine/layers/instance.py:296-386 (AssignMasks), /root/reference/engine/metrics.py:109-165
// (DetectionIOUMetric); restated in oracle/training_oracle.py, whose header fixes the two
// order-dependent TensorFlow scatters (tensor_scatter_nd_update: the last update of the index list
// wins; scatter_nd: repeated updates are added in list order).
//
// The reference materialises IoU matrices ([B*G, N] with N = 163,680 priors) and index lists; here
// nothing of that size exists: AssignBoxes is one thread per (image, prior) that walks the image's
// few ground-truth boxes twice - first the IoU >= 0.5 matches, then the "best prior of a ground
// truth" matches - which is exactly the order in which the reference's concatenated index list hits
// that prior, so "last update wins" and "repeated updates add up" fall out of program order.
#include "common.cuh"

namespace {

constexpr int kAssignThreads = 256;
constexpr int kMaxGt = 1024;              // ground-truth rows per image held in shared memory

// CalculateIOU.call for one pair (a = row box, b = column box), (cx,cy,w,h) each, split into the
// per-box part (corners, area) and the per-pair part.  Disjoint pairs - almost all of them - skip
// the division: 0 / (union + 1e-5) is exactly 0.
struct Corners { float x1, y1, x2, y2, area; };

__device__ __forceinline__ Corners corners_of(const float4 b) {
    const float hw = __fdiv_rn(b.z, 2.0f), hh = __fdiv_rn(b.w, 2.0f);
    Corners c;
    c.x1 = __fsub_rn(b.x, hw); c.y1 = __fsub_rn(b.y, hh);
    c.x2 = __fadd_rn(b.x, hw); c.y2 = __fadd_rn(b.y, hh);
    c.area = __fmul_rn(b.z, b.w);
    return c;
}
__device__ __forceinline__ float iou_corners(const Corners& a, const Corners& b) {
    const float iw = fmaxf(0.0f, __fsub_rn(fminf(b.x2, a.x2), fmaxf(b.x1, a.x1)));
    const float ih = fmaxf(0.0f, __fsub_rn(fminf(b.y2, a.y2), fmaxf(b.y1, a.y1)));
    const float inter = __fmul_rn(iw, ih);
    if (!(inter > 0.0f) && inter == inter) return 0.0f;                 // NaN falls through to the division
    return __fdiv_rn(inter, __fadd_rn(__fsub_rn(__fadd_rn(b.area, a.area), inter), 1e-5f));
}
__device__ __forceinline__ float iou_cxcywh(const float4 a, const float4 b) {
    return iou_corners(corners_of(a), corners_of(b));
}

__device__ __forceinline__ float4 box_at(const void* p, int is_i32, int64_t row, int stride) {
    if (is_i32) {
        const int32_t* q = static_cast<const int32_t*>(p) + row * stride;
        return make_float4((float)q[0], (float)q[1], (float)q[2], (float)q[3]);
    }
    const float* q = static_cast<const float*>(p) + row * stride;
    return make_float4(q[0], q[1], q[2], q[3]);
}

__global__ void __launch_bounds__(kAssignThreads)
calculate_iou_kernel(const float* __restrict__ aa, int na, int sa, const float* __restrict__ bb, int nb, int sb,
                     float* __restrict__ out) {
    const int64_t total = (int64_t)na * nb;
    for (int64_t i = (int64_t)blockIdx.x * kAssignThreads + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * kAssignThreads) {
        const int r = (int)(i / nb), c = (int)(i - (int64_t)r * nb);
        out[i] = iou_cxcywh(box_at(aa, 0, r, sa), box_at(bb, 0, c, sb));
    }
}

// ---- AssignBoxes ------------------------------------------------------------------------------
// best[b,g] = argmax_n iou(gt[b,g], pr[0,n]) * mask  (first maximum, detection.py:631-634)
__global__ void __launch_bounds__(kAssignThreads)
gt_best_prior_kernel(const float* __restrict__ gt, int G, const void* __restrict__ pr, int pr_i32, int N,
                     int32_t* __restrict__ best) {
    __shared__ float s_v[kAssignThreads / 32];
    __shared__ int s_i[kAssignThreads / 32];
    const int g = blockIdx.x, b = blockIdx.y;
    const float* row = gt + ((int64_t)b * G + g) * 6;
    const Corners gb = corners_of(make_float4(row[0], row[1], row[2], row[3]));
    const float mask = row[0] != -1.0f ? 1.0f : 0.0f;
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int n = threadIdx.x; n < N; n += kAssignThreads) {
        const float v = __fmul_rn(iou_corners(gb, corners_of(box_at(pr, pr_i32, n, 4))), mask);
        if (v > bv) { bv = v; bi = n; }                  // ascending n per thread: the first maximum stays
    }
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { s_v[threadIdx.x >> 5] = bv; s_i[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < kAssignThreads / 32; ++w)
            if (s_v[w] > bv || (s_v[w] == bv && s_i[w] < bi)) { bv = s_v[w]; bi = s_i[w]; }
        best[b * G + g] = bi == 0x7fffffff ? 0 : bi;
    }
}

__global__ void __launch_bounds__(kAssignThreads)
assign_boxes_kernel(const float* __restrict__ gt, int G, const void* __restrict__ pr, int pr_i32, int N, int C,
                    const int32_t* __restrict__ best, float* __restrict__ cls_true, float* __restrict__ loc_true,
                    float* __restrict__ assign_mask) {
    __shared__ float s_gt[kMaxGt * 6];
    __shared__ Corners s_gc[kMaxGt];
    __shared__ int s_best[kMaxGt];
    const int b = blockIdx.y;
    for (int i = threadIdx.x; i < G * 6; i += kAssignThreads) s_gt[i] = gt[(int64_t)b * G * 6 + i];
    for (int i = threadIdx.x; i < G; i += kAssignThreads) {
        const float* r = gt + ((int64_t)b * G + i) * 6;
        s_gc[i] = corners_of(make_float4(r[0], r[1], r[2], r[3]));
        s_best[i] = best[b * G + i];
    }
    __syncthreads();
    const int n = blockIdx.x * kAssignThreads + threadIdx.x;
    if (n >= N) return;
    const Corners p0 = corners_of(box_at(pr, pr_i32, n, 4));                 // IoU uses pr_boxes[0] (:627-628)
    const float4 pb = box_at(pr, pr_i32, (int64_t)b * N + n, 4);             // targets use pr_boxes[b] (:665-667)
    float label = -1.0f;
    float loc[4] = {0.f, 0.f, 0.f, 0.f};
    bool ignore = false;
    auto update = [&](const float* g) {                                      // one entry of match_indices
        label = g[4];                                                        // tensor_scatter_nd_update: last wins
        loc[0] = __fadd_rn(loc[0], __fdiv_rn(__fsub_rn(g[0], pb.x), pb.z));  // scatter_nd: repeated updates add
        loc[1] = __fadd_rn(loc[1], __fdiv_rn(__fsub_rn(g[1], pb.y), pb.w));
        loc[2] = __fadd_rn(loc[2], log_cr(__fdiv_rn(g[2], pb.z)));
        loc[3] = __fadd_rn(loc[3], log_cr(__fdiv_rn(g[3], pb.w)));
    };
    for (int g = 0; g < G; ++g) {                                            // tf.where(iou >= 0.5), ascending g
        const float* r = s_gt + g * 6;
        const float mask = r[0] != -1.0f ? 1.0f : 0.0f;
        const float v = __fmul_rn(iou_corners(s_gc[g], p0), mask);
        if (v >= 0.5f) update(r);
        if (v < 0.5f && v >= 0.4f) ignore = true;                            // :651-660
    }
    for (int g = 0; g < G; ++g) {                                            // best_indices of the valid ground truths
        const float* r = s_gt + g * 6;
        if (r[5] > 0.0f && s_best[g] == n) update(r);
    }
    const float v2 = label != -1.0f ? label : (float)C;                      // :644-646
    const int ci = __float2int_rz(v2);
    float* co = cls_true + ((int64_t)b * N + n) * C;
    for (int c = 0; c < C; ++c) co[c] = c == ci ? 1.0f : 0.0f;               // one_hot(.., C + 1)[..., :C]
    float am = ci == C ? 1.0f : 0.0f;                                        // the background column
    if (ignore) am = -1.0f;
    assign_mask[(int64_t)b * N + n] = am;
    reinterpret_cast<float4*>(loc_true)[(int64_t)b * N + n] = make_float4(loc[0], loc[1], loc[2], loc[3]);
}

// ---- AssignMasks --------------------------------------------------------------------------------
// One CTA per (image, RoI): best ground truth of the RoI (same class, both rows valid, first
// maximum), then the crop of that ground truth's mask to the RoI (tf.image.crop_and_resize,
// bilinear, extrapolation 0) thresholded into class ids.
__global__ void __launch_bounds__(128)
assign_masks_kernel(const float* __restrict__ roi, int R, const float* __restrict__ gt, int G,
                    const float* __restrict__ gt_masks, int H, int W, int mh, int mw, int C, float thr,
                    int32_t* __restrict__ out) {
    __shared__ float s_bv[4];
    __shared__ int s_bi[4];
    __shared__ int s_gi;
    __shared__ float s_cls;
    const int r = blockIdx.x, b = blockIdx.y;
    const float* rr = roi + ((int64_t)b * R + r) * 6;
    const float4 rb = make_float4(rr[0], rr[1], rr[2], rr[3]);
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int g = threadIdx.x; g < G; g += 128) {
        const float* gr = gt + ((int64_t)b * G + g) * 6;
        const float valid = (gr[5] != -1.0f && rr[5] != -1.0f) ? 1.0f : 0.0f;
        const float same = gr[4] == rr[4] ? 1.0f : 0.0f;
        const float v = __fmul_rn(__fmul_rn(iou_cxcywh(make_float4(gr[0], gr[1], gr[2], gr[3]), rb), valid), same);
        if (v > bv) { bv = v; bi = g; }
    }
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { s_bv[threadIdx.x >> 5] = bv; s_bi[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 4; ++w)
            if (s_bv[w] > bv || (s_bv[w] == bv && s_bi[w] < bi)) { bv = s_bv[w]; bi = s_bi[w]; }
        if (bi == 0x7fffffff) bi = 0;
        s_gi = bi;
        s_cls = bv >= thr ? gt[((int64_t)b * G + bi) * 6 + 4] : (float)C;
    }
    __syncthreads();
    const int gi = s_gi;
    const float cls = s_cls;
    const float* m = gt_masks + ((int64_t)b * G + gi) * H * W;
    // NormalizeBoxes(shape = (H, W)) of the RoI, then crop_and_resize source coordinates
    const float fh = (float)H, fw = (float)W;
    const float hw = __fdiv_rn(rb.z, 2.0f), hh = __fdiv_rn(rb.w, 2.0f);
    const float x1 = __fdiv_rn(__fsub_rn(rb.x, hw), fw), y1 = __fdiv_rn(__fsub_rn(rb.y, hh), fh);
    const float x2 = __fdiv_rn(__fadd_rn(rb.x, hw), fw), y2 = __fdiv_rn(__fadd_rn(rb.y, hh), fh);
    const float hm1 = (float)(H - 1), wm1 = (float)(W - 1);
    const float sy = mh > 1 ? __fdiv_rn(__fmul_rn(__fsub_rn(y2, y1), hm1), (float)(mh - 1)) : 0.0f;
    const float sx = mw > 1 ? __fdiv_rn(__fmul_rn(__fsub_rn(x2, x1), wm1), (float)(mw - 1)) : 0.0f;
    int32_t* o = out + ((int64_t)b * R + r) * mh * mw;
    for (int i = threadIdx.x; i < mh * mw; i += 128) {
        const int y = i / mw, x = i - y * mw;
        const float in_y = mh > 1 ? __fadd_rn(__fmul_rn(y1, hm1), __fmul_rn((float)y, sy))
                                  : __fmul_rn(__fmul_rn(0.5f, __fadd_rn(y1, y2)), hm1);
        const float in_x = mw > 1 ? __fadd_rn(__fmul_rn(x1, wm1), __fmul_rn((float)x, sx))
                                  : __fmul_rn(__fmul_rn(0.5f, __fadd_rn(x1, x2)), wm1);
        float v = 0.0f;                                                      // extrapolation_value
        if (in_y >= 0.0f && in_y <= hm1 && in_x >= 0.0f && in_x <= wm1) {
            const float fy = floorf(in_y), fx = floorf(in_x);
            const int t = (int)fy, bt = (int)ceilf(in_y), l = (int)fx, rt = (int)ceilf(in_x);
            const float ly = __fsub_rn(in_y, fy), lx = __fsub_rn(in_x, fx);
            const float tl = m[(int64_t)t * W + l], tr = m[(int64_t)t * W + rt];
            const float bl = m[(int64_t)bt * W + l], br = m[(int64_t)bt * W + rt];
            const float top = __fadd_rn(tl, __fmul_rn(__fsub_rn(tr, tl), lx));
            const float bot = __fadd_rn(bl, __fmul_rn(__fsub_rn(br, bl), lx));
            v = __fadd_rn(top, __fmul_rn(__fsub_rn(bot, top), ly));
        }
        o[i] = __float2int_rz(v > 0.5f ? cls : (float)C);
    }
}

// ---- DetectionIOUMetric -------------------------------------------------------------------------
// One CTA per image: precision / recall / F-measure at IoU > 0.5 (metrics.py:117-160).
__global__ void __launch_bounds__(kAssignThreads)
detection_metric_kernel(const float* __restrict__ pred, int P, const float* __restrict__ gt, int G,
                        float* __restrict__ out) {
    __shared__ int s_cnt[4];
    const int b = blockIdx.x;
    if (threadIdx.x < 4) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const float* pp = pred + (int64_t)b * P * 6;
    const float* gg = gt + (int64_t)b * G * 6;
    int pos = 0, tru = 0, npred = 0, ngt = 0;
    for (int i = threadIdx.x; i < P; i += kAssignThreads) {                  // rows: reduce_max over ground truths
        const float* r = pp + i * 6;
        const bool pv = r[0] != -1.0f;
        npred += pv;
        float mx = -INFINITY;
        for (int j = 0; j < G; ++j) {
            const float* q = gg + j * 6;
            const float keep = (pv || q[0] != -1.0f) ? 1.0f : 0.0f;
            mx = fmaxf(mx, __fmul_rn(iou_cxcywh(make_float4(r[0], r[1], r[2], r[3]), make_float4(q[0], q[1], q[2], q[3])), keep));
        }
        pos += mx > 0.5f;
    }
    for (int j = threadIdx.x; j < G; j += kAssignThreads) {                  // columns: reduce_max over predictions
        const float* q = gg + j * 6;
        const bool gv = q[0] != -1.0f;
        ngt += gv;
        float mx = -INFINITY;
        for (int i = 0; i < P; ++i) {
            const float* r = pp + i * 6;
            const float keep = (gv || r[0] != -1.0f) ? 1.0f : 0.0f;
            mx = fmaxf(mx, __fmul_rn(iou_cxcywh(make_float4(r[0], r[1], r[2], r[3]), make_float4(q[0], q[1], q[2], q[3])), keep));
        }
        tru += mx > 0.5f;
    }
    atomicAdd(&s_cnt[0], pos); atomicAdd(&s_cnt[1], tru); atomicAdd(&s_cnt[2], npred); atomicAdd(&s_cnt[3], ngt);
    __syncthreads();
    if (threadIdx.x == 0) {
        const float eps = 1e-7f;                                             // K.epsilon()
        const float p = __fdiv_rn((float)s_cnt[0], __fadd_rn((float)s_cnt[2], eps));
        const float r = __fdiv_rn((float)s_cnt[1], __fadd_rn((float)s_cnt[3], eps));
        const float f = __fdiv_rn(__fmul_rn(2.0f, __fmul_rn(p, r)), __fadd_rn(__fadd_rn(p, r), eps));
        out[b * 3] = p; out[b * 3 + 1] = r; out[b * 3 + 2] = f;
    }
}

}  // namespace

// ================================================================ host side ===
extern "C" int mlp_calculate_iou(mlp_ctx* ctx, const float* aa_dev, int num_aa, int aa_stride, const float* bb_dev,
                                 int num_bb, int bb_stride, float* out_dev, mlp_stream_t stream) {
    MLP_CHECK_ARG(ctx && aa_dev && bb_dev && out_dev, "mlp_calculate_iou: NULL argument");
    MLP_CHECK_ARG(num_aa >= 1 && num_bb >= 1 && aa_stride >= 4 && bb_stride >= 4, "mlp_calculate_iou: bad shape");
    DeviceGuard g(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope prof(ctx, MLP_ST_ASSIGN, st);
    const int64_t total = (int64_t)num_aa * num_bb;
    int64_t blocks = (total + kAssignThreads - 1) / kAssignThreads;
    const int64_t cap = (int64_t)ctx->sm_count * 32;
    calculate_iou_kernel<<<(int)(blocks < cap ? blocks : cap), kAssignThreads, 0, st>>>(aa_dev, num_aa, aa_stride, bb_dev,
                                                                                        num_bb, bb_stride, out_dev);
    MLP_LAUNCH_CHECK(ctx);
    return MLP_OK;
}

extern "C" int mlp_assign_boxes(mlp_ctx* ctx, const float* gt_boxes_dev, const void* pr_boxes_dev, int pr_dtype,
                                int batch, int num_gt, int num_priors, int num_classes, float* cls_true_dev,
                                float* loc_true_dev, float* assign_mask_dev, mlp_stream_t stream) {
    MLP_CHECK_ARG(ctx && gt_boxes_dev && pr_boxes_dev && cls_true_dev && loc_true_dev && assign_mask_dev,
                  "mlp_assign_boxes: NULL argument");
    MLP_CHECK_ARG(batch >= 1 && batch <= 65535 && num_gt >= 1 && num_gt <= kMaxGt && num_priors >= 1 && num_classes >= 1,
                  "mlp_assign_boxes: bad shape B=%d G=%d N=%d C=%d (G <= %d)", batch, num_gt, num_priors, num_classes,
                  kMaxGt);
    MLP_CHECK_ARG(pr_dtype == MLP_F32 || pr_dtype == MLP_I32, "mlp_assign_boxes: priors must be f32 or i32");
    MLP_CHECK_ARG(mlp_aligned16(loc_true_dev), "mlp_assign_boxes: loc_true_dev must be 16-byte aligned");
    DeviceGuard g(ctx->device);
    int rc = mlp_ensure_scratch(ctx, MLP_ARENA_ASSIGN, (int64_t)batch * num_gt * 4);
    if (rc) return rc;
    int32_t* best = static_cast<int32_t*>(ctx->arena[MLP_ARENA_ASSIGN]);
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope prof(ctx, MLP_ST_ASSIGN, st);
    gt_best_prior_kernel<<<dim3(num_gt, batch), kAssignThreads, 0, st>>>(gt_boxes_dev, num_gt, pr_boxes_dev,
                                                                         pr_dtype == MLP_I32, num_priors, best);
    MLP_LAUNCH_CHECK(ctx);
    assign_boxes_kernel<<<dim3((num_priors + kAssignThreads - 1) / kAssignThreads, batch), kAssignThreads, 0, st>>>(
        gt_boxes_dev, num_gt, pr_boxes_dev, pr_dtype == MLP_I32, num_priors, num_classes, best, cls_true_dev,
        loc_true_dev, assign_mask_dev);
    MLP_LAUNCH_CHECK(ctx);
    return MLP_OK;
}

extern "C" int mlp_assign_masks(mlp_ctx* ctx, const float* roi_boxes_dev, int num_rois, const float* gt_boxes_dev,
                                int num_gt, const float* gt_masks_dev, int batch, int height, int width, int mask_h,
                                int mask_w, int num_classes, float match_iou_threshold, int32_t* out_dev,
                                mlp_stream_t stream) {
    MLP_CHECK_ARG(ctx && roi_boxes_dev && gt_boxes_dev && gt_masks_dev && out_dev, "mlp_assign_masks: NULL argument");
    MLP_CHECK_ARG(batch >= 1 && batch <= 65535 && num_rois >= 1 && num_gt >= 1 && height >= 1 && width >= 1 &&
                      mask_h >= 1 && mask_w >= 1 && num_classes >= 1,
                  "mlp_assign_masks: bad shape");
    DeviceGuard g(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope prof(ctx, MLP_ST_ASSIGN, st);
    assign_masks_kernel<<<dim3(num_rois, batch), 128, 0, st>>>(roi_boxes_dev, num_rois, gt_boxes_dev, num_gt,
                                                               gt_masks_dev, height, width, mask_h, mask_w, num_classes,
                                                               match_iou_threshold, out_dev);
    MLP_LAUNCH_CHECK(ctx);
    return MLP_OK;
}

extern "C" int mlp_detection_iou_metric(mlp_ctx* ctx, const float* pred_boxes_dev, int num_pred,
                                        const float* gt_boxes_dev, int num_gt, int batch, float* out_dev,
                                        mlp_stream_t stream) {
    MLP_CHECK_ARG(ctx && pred_boxes_dev && gt_boxes_dev && out_dev, "mlp_detection_iou_metric: NULL argument");
    MLP_CHECK_ARG(batch >= 1 && num_pred >= 1 && num_gt >= 1, "mlp_detection_iou_metric: bad shape");
    DeviceGuard g(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope prof(ctx, MLP_ST_ASSIGN, st);
    detection_metric_kernel<<<batch, kAssignThreads, 0, st>>>(pred_boxes_dev, num_pred, gt_boxes_dev, num_gt, out_dev);
    MLP_LAUNCH_CHECK(ctx);
    return MLP_OK;
}
