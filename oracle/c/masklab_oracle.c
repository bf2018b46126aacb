/* masklab_oracle.c — plain-C restatement of the reference's post-backbone path.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): a second, independently written CPU
 * restatement used (a) to cross-check the NumPy oracle and (b) as the multi-threaded CPU
 * baseline of bench.py.  PARITY UNPINNED: the reference has no golden vectors for this
 * path and its arithmetic lives in TensorFlow 1.14/1.15 kernels that are not vendored; their
 * published algorithms are restated here with the reference call site next to each function
 * (paths relative to /root/reference).
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared  (no FMA contraction, no fast-math: every
 * float operation rounds once, like the TF CPU kernels and NumPy).  Single-threaded C; the
 * image has no OpenMP runtime, so callers parallelise over frames with threads through ctypes
 * (frames are independent, the GIL is released during the call).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MLO_MAX_LEVELS 8
#define MLO_MAX_ANCHORS 32

typedef struct {
    int32_t num_levels, padding_same;
    int32_t stride[MLO_MAX_LEVELS], num_anchors[MLO_MAX_LEVELS];
    int32_t anchor_w[MLO_MAX_LEVELS][MLO_MAX_ANCHORS], anchor_h[MLO_MAX_LEVELS][MLO_MAX_ANCHORS];
} mlo_prior;

static float exp_cr(float x) { return (float)exp((double)x); }   /* correctly rounded f32 */
static float log_cr(float x) { return (float)log((double)x); }

/* ---- a2: PriorLayer.call, engine/layers/detection.py:269-298 ------------------------- */
int64_t mlo_prior_count(const mlo_prior* p, int H, int W) {
    int64_t n = 0;
    for (int l = 0; l < p->num_levels; ++l) {
        int s = p->stride[l];
        int hf = p->padding_same ? (H + s - 1) / s : H / s, wf = p->padding_same ? (W + s - 1) / s : W / s;
        n += (int64_t)hf * wf * p->num_anchors[l];
    }
    return n;
}

void mlo_prior_layer(const mlo_prior* p, int H, int W, int32_t* out /* [N,4] cx,cy,w,h */) {
    int64_t n = 0;
    for (int l = 0; l < p->num_levels; ++l) {
        int s = p->stride[l];
        int hf = p->padding_same ? (H + s - 1) / s : H / s, wf = p->padding_same ? (W + s - 1) / s : W / s;
        for (int y = 0; y < hf; ++y)
            for (int x = 0; x < wf; ++x)
                for (int a = 0; a < p->num_anchors[l]; ++a, ++n) {
                    out[n * 4 + 0] = s / 2 + x * s;
                    out[n * 4 + 1] = s / 2 + y * s;
                    out[n * 4 + 2] = p->anchor_w[l][a];
                    out[n * 4 + 3] = p->anchor_h[l][a];
                }
    }
}

/* ---- a3: RestoreBoxes.call, detection.py:325-344 -------------------------------------- */
void mlo_restore_boxes(const float* loc, const int32_t* prior, int64_t rows, float* out) {
    for (int64_t i = 0; i < rows; ++i) {
        float px = (float)prior[i * 4], py = (float)prior[i * 4 + 1], pw = (float)prior[i * 4 + 2],
              ph = (float)prior[i * 4 + 3];
        out[i * 4 + 0] = loc[i * 4 + 0] * pw + px;
        out[i * 4 + 1] = loc[i * 4 + 1] * ph + py;
        out[i * 4 + 2] = exp_cr(loc[i * 4 + 2]) * pw;
        out[i * 4 + 3] = exp_cr(loc[i * 4 + 3]) * ph;
    }
}

/* ---- tf.image.non_max_suppression (NonMaxSuppressionV3), literal pop-and-check loop ---- */
typedef struct { float score; int idx; } cand_t;
static int cand_cmp(const void* a, const void* b) {
    const cand_t *x = a, *y = b;
    if (x->score > y->score) return -1;
    if (x->score < y->score) return 1;
    return x->idx - y->idx;                       /* ties: lower index first */
}
static float iou_tf(const float* a, const float* b) {   /* boxes (y1,x1,y2,x2) */
    float ymin_i = fminf(a[0], a[2]), xmin_i = fminf(a[1], a[3]), ymax_i = fmaxf(a[0], a[2]), xmax_i = fmaxf(a[1], a[3]);
    float ymin_j = fminf(b[0], b[2]), xmin_j = fminf(b[1], b[3]), ymax_j = fmaxf(b[0], b[2]), xmax_j = fmaxf(b[1], b[3]);
    float area_i = (ymax_i - ymin_i) * (xmax_i - xmin_i), area_j = (ymax_j - ymin_j) * (xmax_j - xmin_j);
    if (area_i <= 0 || area_j <= 0) return 0.0f;
    float ih = fmaxf(fminf(ymax_i, ymax_j) - fmaxf(ymin_i, ymin_j), 0.0f);
    float iw = fmaxf(fminf(xmax_i, xmax_j) - fmaxf(xmin_i, xmin_j), 0.0f);
    float inter = ih * iw;
    return inter / (area_i + area_j - inter);
}
/* boxes [n,4], scores [n] -> selected indices (selection order); returns count */
static int nms_tf(const float* boxes, const float* scores, int n, int max_out, float thr, int* sel, cand_t* work) {
    for (int i = 0; i < n; ++i) { work[i].score = scores[i]; work[i].idx = i; }
    qsort(work, (size_t)n, sizeof(cand_t), cand_cmp);
    int k = 0;
    for (int r = 0; r < n && k < max_out; ++r) {
        int i = work[r].idx, keep = 1;
        for (int q = k - 1; q >= 0; --q)
            if (iou_tf(boxes + 4 * (int64_t)i, boxes + 4 * (int64_t)sel[q]) > thr) { keep = 0; break; }
        if (keep) sel[k++] = i;
    }
    return k;
}

/* ---- a4-a8: DetectionProposal.call, detection.py:482-567 (+ MoldBatch layout) ---------
 * cls [B,N,C], boxes [B,N,4] cxcywh -> det [B,K,6] (-1 padded), keep [B,K,2] (n,c), counts [B];
 * returns M = max(1, max count).  K = max_out capacity. */
int mlo_detection_proposal(const float* cls, const float* boxes, int B, int N, int C, float min_conf,
                           float nms_thr, float post_thr, int max_out, float* det, int32_t* keep,
                           int32_t* counts) {
    int M = 1;
    for (int b = 0; b < B; ++b) {
        const float* cl = cls + (int64_t)b * N * C;
        const float* bx = boxes + (int64_t)b * N * 4;
        /* candidates in (n,c) scan order, grouped by class; group order = first appearance */
        int* ccount = calloc((size_t)C, sizeof(int));
        int* first = malloc((size_t)C * sizeof(int));
        int total = 0;
        for (int c = 0; c < C; ++c) first[c] = -1;
        for (int n = 0; n < N; ++n)
            for (int c = 0; c < C; ++c)
                if (cl[(int64_t)n * C + c] >= min_conf) {
                    if (first[c] < 0) first[c] = total;
                    ccount[c]++; total++;
                }
        int* order = malloc((size_t)C * sizeof(int));
        int ng = 0;
        for (int c = 0; c < C; ++c) if (ccount[c] > 0) order[ng++] = c;
        for (int i = 1; i < ng; ++i) {            /* insertion sort by first appearance */
            int c = order[i], j = i;
            while (j > 0 && first[order[j - 1]] > first[c]) { order[j] = order[j - 1]; --j; }
            order[j] = c;
        }
        /* per class NMS */
        int cap = ng * max_out + 1;
        float* pc_box = malloc((size_t)cap * 4 * sizeof(float));     /* normalised (y1,x1,y2,x2) */
        float* pc_score = malloc((size_t)cap * sizeof(float));
        int* pc_n = malloc((size_t)cap * sizeof(int));
        int* pc_c = malloc((size_t)cap * sizeof(int));
        int pc_total = 0;
        for (int gi = 0; gi < ng; ++gi) {
            int c = order[gi], m = ccount[c];
            float* gb = malloc((size_t)m * 4 * sizeof(float));
            float* gs = malloc((size_t)m * sizeof(float));
            int* gn = malloc((size_t)m * sizeof(int));
            int* sel = malloc((size_t)(max_out > 0 ? max_out : 1) * sizeof(int));
            cand_t* work = malloc((size_t)m * sizeof(cand_t));
            int k = 0;
            for (int n = 0; n < N; ++n) {
                float s = cl[(int64_t)n * C + c];
                if (s >= min_conf) {
                    const float* q = bx + (int64_t)n * 4;             /* NormalizeBoxes, shape = ones */
                    gb[k * 4 + 0] = (q[1] - q[3] / 2.0f) / 1.0f;
                    gb[k * 4 + 1] = (q[0] - q[2] / 2.0f) / 1.0f;
                    gb[k * 4 + 2] = (q[1] + q[3] / 2.0f) / 1.0f;
                    gb[k * 4 + 3] = (q[0] + q[2] / 2.0f) / 1.0f;
                    gs[k] = s; gn[k] = n; ++k;
                }
            }
            int ns = nms_tf(gb, gs, m, max_out, nms_thr, sel, work);
            for (int i = 0; i < ns; ++i, ++pc_total) {
                memcpy(pc_box + 4 * (int64_t)pc_total, gb + 4 * (int64_t)sel[i], 4 * sizeof(float));
                pc_score[pc_total] = gs[sel[i]]; pc_n[pc_total] = gn[sel[i]]; pc_c[pc_total] = c;
            }
            free(gb); free(gs); free(gn); free(sel); free(work);
        }
        /* cross-class NMS over the concatenated survivors */
        int* sel = malloc((size_t)(max_out > 0 ? max_out : 1) * sizeof(int));
        cand_t* work = malloc((size_t)(pc_total + 1) * sizeof(cand_t));
        int ns = nms_tf(pc_box, pc_score, pc_total, max_out, post_thr, sel, work);
        float* d = det + (int64_t)b * max_out * 6;
        int32_t* kp = keep + (int64_t)b * max_out * 2;
        for (int i = 0; i < max_out; ++i) {
            if (i < ns) {
                int p = sel[i], n = pc_n[p];
                memcpy(d + i * 6, bx + (int64_t)n * 4, 4 * sizeof(float));
                d[i * 6 + 4] = (float)pc_c[p];
                d[i * 6 + 5] = pc_score[p];
                kp[i * 2] = n; kp[i * 2 + 1] = pc_c[p];
            } else {
                for (int q = 0; q < 6; ++q) d[i * 6 + q] = -1.0f;
                kp[i * 2] = kp[i * 2 + 1] = -1;
            }
        }
        counts[b] = ns;
        { if (ns > M) M = ns; }
        free(sel); free(work); free(pc_box); free(pc_score); free(pc_n); free(pc_c);
        free(order); free(first); free(ccount);
    }
    return M;
}

/* ---- a9: MaskDistribute.call, engine/layers/instance.py:52-66 ------------------------- */
void mlo_mask_distribute(const float* det, int64_t rows, int max_k, float base_size, float* out) {
    float base_eps = (float)((double)base_size + 1e-7);
    for (int64_t i = 0; i < rows; ++i) {
        const float* r = det + i * 6;
        float size = sqrtf(r[2] * r[3]);
        float dk = log_cr((size + 1e-7f) / base_eps) / log_cr(2.0f);
        float k = floorf(dk);
        k = fminf(fmaxf(k, 0.0f), (float)max_k);
        out[i * 7] = (r[0] == -1.0f) ? r[0] : k;
        memcpy(out + i * 7 + 1, r, 6 * sizeof(float));
    }
}

/* ---- tf.image.crop_and_resize, one box (bilinear, extrapolation 0) --------------------- */
static void crop_and_resize_one(const float* img, int Hf, int Wf, int D, float y1, float x1, float y2,
                                float x2, int ch, int cw, float* out) {
    float hs = (ch > 1) ? (y2 - y1) * (float)(Hf - 1) / (float)(ch - 1) : 0.0f;
    float ws = (cw > 1) ? (x2 - x1) * (float)(Wf - 1) / (float)(cw - 1) : 0.0f;
    for (int y = 0; y < ch; ++y) {
        float in_y = (ch > 1) ? y1 * (float)(Hf - 1) + (float)y * hs : 0.5f * (y1 + y2) * (float)(Hf - 1);
        float* orow = out + (int64_t)y * cw * D;
        if (!(in_y >= 0 && in_y <= (float)(Hf - 1))) { memset(orow, 0, (size_t)cw * D * sizeof(float)); continue; }
        int top = (int)floorf(in_y), bot = (int)ceilf(in_y);
        float ly = in_y - floorf(in_y);
        for (int x = 0; x < cw; ++x) {
            float in_x = (cw > 1) ? x1 * (float)(Wf - 1) + (float)x * ws : 0.5f * (x1 + x2) * (float)(Wf - 1);
            float* o = orow + (int64_t)x * D;
            if (!(in_x >= 0 && in_x <= (float)(Wf - 1))) { memset(o, 0, (size_t)D * sizeof(float)); continue; }
            int left = (int)floorf(in_x), right = (int)ceilf(in_x);
            float lx = in_x - floorf(in_x);
            const float* tl = img + ((int64_t)top * Wf + left) * D;
            const float* tr = img + ((int64_t)top * Wf + right) * D;
            const float* bl = img + ((int64_t)bot * Wf + left) * D;
            const float* br = img + ((int64_t)bot * Wf + right) * D;
            for (int d = 0; d < D; ++d) {
                float t = tl[d] + (tr[d] - tl[d]) * lx;
                float bb = bl[d] + (br[d] - bl[d]) * lx;
                o[d] = t + (bb - t) * ly;
            }
        }
    }
}

/* ---- a10: PyramidRoiAlign.call, instance.py:109-139 -----------------------------------
 * Two passes like the CUDA path: plan (counts per level/image, Mf) then run.  dist [B,M,7]. */
void mlo_roi_plan(const float* dist, int B, int M, int L, int32_t* level_counts /*[L,B]*/, int32_t* level_m /*[L+1]*/) {
    int R = 0;
    for (int f = 0; f < L; ++f) {
        int mx = 1;
        for (int b = 0; b < B; ++b) {
            int c = 0;
            for (int j = 0; j < M; ++j) c += dist[((int64_t)b * M + j) * 7] == (float)f;
            level_counts[f * B + b] = c;
            if (c > mx) mx = c;
        }
        level_m[f] = mx; R += mx;
    }
    level_m[L] = R;
}

void mlo_roi_run(const float* const* fmaps, const int32_t* fh, const int32_t* fw, int L, int Cf,
                 const float* dist, int B, int M, float image_h, float image_w, int ch, int cw,
                 const int32_t* level_m, float* const* crops, float* roi_boxes) {
    int R = level_m[L];
    int64_t crop_elems = (int64_t)ch * cw * Cf;
    for (int b = 0; b < B; ++b) {
        int off = 0;
        for (int f = 0; f < L; ++f) {
            int mf = level_m[f], slot = 0;
            const float* img = fmaps[f] + (int64_t)b * fh[f] * fw[f] * Cf;
            for (int j = 0; j < M; ++j) {
                const float* r = dist + ((int64_t)b * M + j) * 7;
                if (r[0] != (float)f) continue;
                float cx = r[1], cy = r[2], w = r[3], h = r[4];
                float x1 = (cx - w / 2.0f) / image_w, y1 = (cy - h / 2.0f) / image_h;
                float x2 = (cx + w / 2.0f) / image_w, y2 = (cy + h / 2.0f) / image_h;
                crop_and_resize_one(img, fh[f], fw[f], Cf, y1, x1, y2, x2, ch, cw,
                                    crops[f] + ((int64_t)b * mf + slot) * crop_elems);
                memcpy(roi_boxes + ((int64_t)b * R + off + slot) * 6, r + 1, 6 * sizeof(float));
                ++slot;
            }
            for (; slot < mf; ++slot) {                    /* MoldBatch padding */
                float* c = crops[f] + ((int64_t)b * mf + slot) * crop_elems;
                for (int64_t i = 0; i < crop_elems; ++i) c[i] = -1.0f;
                for (int q = 0; q < 6; ++q) roi_boxes[((int64_t)b * R + off + slot) * 6 + q] = -1.0f;
            }
            off += mf;
        }
    }
}

/* ---- a11: TrimInstances.call, instance.py:258-277 -------------------------------------- */
int mlo_trim_plan(const float* roi_boxes, int B, int R, int32_t* counts) {
    int M = 1;
    for (int b = 0; b < B; ++b) {
        int c = 0;
        for (int j = 0; j < R; ++j) c += roi_boxes[((int64_t)b * R + j) * 6 + 4] != -1.0f;
        counts[b] = c;
        if (c > M) M = c;
    }
    return M;
}

void mlo_trim_run(const float* roi_boxes, const float* roi_masks, int B, int R, int mh, int mw, int C,
                  int M, float* out_boxes, float* out_masks) {
    int px = mh * mw;
    for (int b = 0; b < B; ++b) {
        int slot = 0;
        for (int j = 0; j < R; ++j) {
            const float* r = roi_boxes + ((int64_t)b * R + j) * 6;
            if (r[4] == -1.0f) continue;
            int cls = (int)r[4];
            memcpy(out_boxes + ((int64_t)b * M + slot) * 6, r, 6 * sizeof(float));
            const float* m = roi_masks + ((int64_t)b * R + j) * px * C;
            float* o = out_masks + ((int64_t)b * M + slot) * px;
            for (int p = 0; p < px; ++p) o[p] = m[(int64_t)p * C + cls];
            ++slot;
        }
        for (; slot < M; ++slot) {
            for (int q = 0; q < 6; ++q) out_boxes[((int64_t)b * M + slot) * 6 + q] = -1.0f;
            for (int p = 0; p < px; ++p) out_masks[((int64_t)b * M + slot) * px + p] = -1.0f;
        }
    }
}

/* ---- a12: UpSampleOutput.call (instance part), engine/layers/misc.py:169-188 ----------- */
void mlo_upsample(const float* det, int64_t rows, float ratio_h, float ratio_w, int32_t* det_i,
                  const float* masks, int64_t n, int32_t* masks_i) {
    for (int64_t i = 0; i < rows; ++i) {
        const float* r = det + i * 6;
        det_i[i * 6 + 0] = (int32_t)(r[0] * ratio_h);     /* cx * ratio[0]  (sic, misc.py:180) */
        det_i[i * 6 + 1] = (int32_t)(r[1] * ratio_w);
        det_i[i * 6 + 2] = (int32_t)(r[2] * ratio_h);
        det_i[i * 6 + 3] = (int32_t)(r[3] * ratio_w);
        det_i[i * 6 + 4] = (int32_t)r[4];
        det_i[i * 6 + 5] = (int32_t)(r[5] * 100.0f);
    }
    for (int64_t i = 0; i < n; ++i) masks_i[i] = masks[i] > 0.5f;
}

/* ---- a13/a14: CropAndPadMask.call, misc.py:358-401; legacy ResizeBilinear align_corners --
 * out_u8 != NULL: binary masks (pasted > 0.5, misc.py:457); out_f32 != NULL: raw values. */
static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

void mlo_paste(const int32_t* det, const int32_t* masks, int B, int M, int mh, int mw, int PH, int PW,
               float* out_f32, uint8_t* out_u8, uint8_t* out_bits /* [B,M,PH,PW/8], bit k of byte i = pixel 8i+k */) {
    int mx = INT32_MIN;
    for (int64_t i = 0; i < (int64_t)B * M; ++i) if (det[i * 6 + 5] > mx) mx = det[i * 6 + 5];
    int thr = (mx > 50) ? 50 : -100;
    int64_t frame = (int64_t)PH * PW;
    for (int64_t inst = 0; inst < (int64_t)B * M; ++inst) {
        const int32_t* r = det + inst * 6;
        if (out_f32) memset(out_f32 + inst * frame, 0, (size_t)frame * sizeof(float));
        if (out_u8) memset(out_u8 + inst * frame, 0, (size_t)frame);
        if (out_bits) memset(out_bits + inst * (frame / 8), 0, (size_t)(frame / 8));
        if (r[5] < thr) continue;
        float cx = (float)(r[0] > 1 ? r[0] : 1), cy = (float)(r[1] > 1 ? r[1] : 1);
        float w = (float)(r[2] > 1 ? r[2] : 1), h = (float)(r[3] > 1 ? r[3] : 1);
        int xmin = clampi((int)ceilf(cx - w / 2.0f), 0, PW), xmax = clampi((int)ceilf(cx + w / 2.0f), 0, PW);
        int ymin = clampi((int)ceilf(cy - h / 2.0f), 0, PH), ymax = clampi((int)ceilf(cy + h / 2.0f), 0, PH);
        int oh = ymax - ymin, ow = xmax - xmin;
        if (oh <= 0 || ow <= 0) continue;             /* TF raises InvalidArgument; defined as zeros */
        float sy = (oh > 1) ? (float)(mh - 1) / (float)(oh - 1) : (float)mh / (float)oh;
        float sx = (ow > 1) ? (float)(mw - 1) / (float)(ow - 1) : (float)mw / (float)ow;
        const int32_t* m = masks + inst * mh * mw;
        for (int y = 0; y < oh; ++y) {
            float py = (float)y * sy, fy = floorf(py);
            int ylo = (int)fy > 0 ? (int)fy : 0, yhi = (int)ceilf(py) < mh - 1 ? (int)ceilf(py) : mh - 1;
            float ly = py - fy;
            for (int x = 0; x < ow; ++x) {
                float px = (float)x * sx, fx = floorf(px);
                int xlo = (int)fx > 0 ? (int)fx : 0, xhi = (int)ceilf(px) < mw - 1 ? (int)ceilf(px) : mw - 1;
                float lx = px - fx;
                float tl = (float)m[ylo * mw + xlo], tr = (float)m[ylo * mw + xhi];
                float bl = (float)m[yhi * mw + xlo], br = (float)m[yhi * mw + xhi];
                float t = tl + (tr - tl) * lx, bb = bl + (br - bl) * lx;
                float v = t + (bb - t) * ly;
                int64_t o = inst * frame + (int64_t)(ymin + y) * PW + (xmin + x);
                if (out_f32) out_f32[o] = v;
                if (out_u8) out_u8[o] = v > 0.5f;
                if (out_bits && v > 0.5f) out_bits[o >> 3] |= (uint8_t)(1u << (o & 7));
            }
        }
    }
}

