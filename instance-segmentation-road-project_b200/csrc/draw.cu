// draw.cu — the overlay layers of the serving graph: DrawSegmentation and
// DrawInstance.
//
// Reference: /root/reference/engine/layers/misc.py:404-429 (DrawSegmentation: clip(images +
// (sum_c colors[c] * seg[..., c]) * alpha, 0, 255) -> uint8), :432-475 (DrawInstance: per image and
// class, (sum of the pasted masks of that class) > 0.5, then DrawSegmentation), wired in
// /root/reference/road_project/setup/serving.py:34-40.  Restated in oracle/draw_oracle.py.
//
//   draw_segmentation_kernel  element-wise, one thread per pixel.
//   draw_instance_kernel      drop-in: reads the [B,M,PH,PW] masks once (float32 or uint8), four
//                             pixels per thread, instances grouped per class in shared memory so
//                             that the per-class sums live in registers (HBM read-bound).
//   pack_tiles_kernel +       fused: every instance's 28x28 tile as 28 bit rows plus its clipped-box
//   draw_tiles_kernel         geometry; then one CTA per 16x64 pixel block finds the instances whose
//                             box touches the block and evaluates their float32 paste values from
//                             the bit rows (the values CropAndPadMask would write) only there.  The
//                             [B,M,PH,PW] tensor is never written or read; the semantic overlay of
//                             serving.py:38-40 can ride along in the same pass.
#include "paste_common.cuh"

namespace {

constexpr int kDrawThreads = 256;
constexpr int kMaxDrawInst = 2048;        // instances per image the class lists hold (MLP_MAX_KEEP)

__device__ __forceinline__ float blend(float img, float csum, float alpha) {
    // clip_by_value(images + color * alpha, 0, 255); the uint8 cast truncates
    return fminf(fmaxf(__fadd_rn(img, __fmul_rn(csum, alpha)), 0.0f), 255.0f);
}
__device__ __forceinline__ float px_f32(const uint8_t* p) { return (float)__ldg(p); }
__device__ __forceinline__ float px_f32(const float* p) { return __ldg(p); }
__device__ __forceinline__ float px_f32(const int32_t* p) { return (float)__ldg(p); }

// ---- DrawSegmentation -----------------------------------------------------------------------
template <typename ImgT, typename SegT>
__global__ void __launch_bounds__(kDrawThreads)
draw_segmentation_kernel(const ImgT* __restrict__ images, const SegT* __restrict__ seg, int64_t npix,
                         const mlp_draw_colors col, uint8_t* __restrict__ out) {
    const int64_t p = (int64_t)blockIdx.x * kDrawThreads + threadIdx.x;
    if (p >= npix) return;
    const int C = col.num_classes;
    float cs[3] = {0.f, 0.f, 0.f};
    const SegT* s = seg + p * C;
    for (int c = 0; c < C; ++c) {                          // reduce_sum(colors * seg[..., None], axis=-2)
        const float v = px_f32(s + c);
#pragma unroll
        for (int k = 0; k < 3; ++k) cs[k] = __fadd_rn(cs[k], __fmul_rn(col.rgb[c][k], v));
    }
#pragma unroll
    for (int k = 0; k < 3; ++k)
        out[p * 3 + k] = (uint8_t)__float2uint_rz(blend(px_f32(images + p * 3 + k), cs[k], col.alpha));
}

// ---- DrawInstance over the materialised masks -----------------------------------------------
// One thread per 16 bytes of the frame (4 float32 / 16 uint8 mask pixels, MaskVec in
// paste_common.cuh): every mask of the image is read once with 128-bit loads.
template <typename ImgT, typename MaskT>
__global__ void __launch_bounds__(kDrawThreads)
draw_instance_kernel(const ImgT* __restrict__ images, const int32_t* __restrict__ det, const MaskT* __restrict__ masks,
                     int m_rows, int m_stride, const int32_t* __restrict__ m_dev, int npx,
                     const mlp_draw_colors col, uint8_t* __restrict__ out) {
    using Vec = MaskVec<MaskT>;
    constexpr int kPx = Vec::kPx;
    __shared__ unsigned short s_list[kMaxDrawInst];        // instances grouped by class, order j inside
    __shared__ int s_beg[MLP_MAX_DRAW_CLASSES + 1];
    const int b = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31;
    int M = m_dev ? *m_dev : m_rows;
    if (M > m_rows) M = m_rows;
    if (m_stride == 0) m_stride = M;
    const int C = col.num_classes;
    if (tid < 32) {                                        // tf.where(det[..., -2] == class_id) per class
        int n = 0;
        for (int c = 0; c < C; ++c) {
            if (lane == 0) s_beg[c] = n;
            for (int j0 = 0; j0 < M; j0 += 32) {
                const int j = j0 + lane;
                const bool hit = j < M && det[((int64_t)b * m_stride + j) * 6 + 4] == c;
                const unsigned m = __ballot_sync(0xffffffffu, hit);
                if (hit) s_list[n + __popc(m & ((1u << lane) - 1u))] = (unsigned short)j;
                n += __popc(m);
            }
        }
        if (lane == 0) s_beg[C] = n;
    }
    __syncthreads();
    const int p = (blockIdx.x * kDrawThreads + tid) * kPx; // kPx consecutive pixels of the frame
    if (p >= npx) return;
    const int n = min(kPx, npx - p);
    const bool vec = n == kPx && (npx % kPx) == 0;         // every mask plane starts 16-byte aligned
    const MaskT* mb = masks + (int64_t)b * M * npx + p;
    float cs[kPx][3];
#pragma unroll
    for (int q = 0; q < kPx; ++q) cs[q][0] = cs[q][1] = cs[q][2] = 0.0f;
    for (int c = 0; c < C; ++c) {
        float acc[kPx];                                    // reduce_sum of the class's masks, order j
#pragma unroll
        for (int q = 0; q < kPx; ++q) acc[q] = 0.0f;
        const int i1 = s_beg[c + 1];
        for (int i = s_beg[c]; i < i1; i += 4) {           // four mask loads in flight
            Vec v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const MaskT* src = mb + (int64_t)s_list[min(i + u, i1 - 1)] * npx;
                if (vec) v[u].load(src); else v[u].load_tail(src, n);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (i + u < i1 && v[u].any())              // adding zeros is exact: skip them
#pragma unroll
                    for (int q = 0; q < kPx; ++q) acc[q] = __fadd_rn(acc[q], v[u].at(q));
        }
#pragma unroll
        for (int q = 0; q < kPx; ++q)
            if (acc[q] > 0.5f)
#pragma unroll
                for (int k = 0; k < 3; ++k) cs[q][k] = __fadd_rn(cs[q][k], col.rgb[c][k]);
    }
    const int64_t o = ((int64_t)b * npx + p) * 3;
#pragma unroll
    for (int q = 0; q < kPx; ++q)
        if (q < n)
#pragma unroll
            for (int k = 0; k < 3; ++k)
                out[o + q * 3 + k] = (uint8_t)__float2uint_rz(blend(px_f32(images + o + q * 3 + k), cs[q][k], col.alpha));
}

// ---- DrawInstance straight from the mask tiles ----------------------------------------------
struct DrawGeom {                         // clipped box of an instance in the frame, 32 bytes
    int xmin, xmax, ymin, ymax;
    float sx, sy;
    int cls;                              // -1: inactive (filtered, empty box) or class outside the colour table
    int pad;
};

// One warp per instance: bit rows of its {0,1} tile (mask_w <= 32) and its geometry.
__global__ void __launch_bounds__(kDrawThreads)
pack_tiles_kernel(const int32_t* __restrict__ det, const PasteSrc S, int B, int m_rows, int m_stride, int mh, int mw,
                  int PH, int PW, int num_colors, DrawGeom* __restrict__ geom, uint32_t* __restrict__ bits,
                  int32_t* __restrict__ m_used) {
    int M, thr;
    paste_scalars(S, B, m_rows, M, thr);
    if (m_stride == 0) m_stride = M;
    if (blockIdx.x == 0 && threadIdx.x == 0) *m_used = M;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t inst = (int64_t)blockIdx.x * (kDrawThreads / 32) + warp;
    if (inst >= (int64_t)B * M) return;
    const int b = (int)(inst / M), j = (int)(inst - (int64_t)b * M);
    const int32_t* row = det + ((int64_t)b * m_stride + j) * 6;
    const PasteGeom g = paste_geometry(row, thr, mh, mw, PH, PW);
    const int cls = row[4];
    const bool draw = g.active && cls >= 0 && cls < num_colors;
    const int64_t slot = (int64_t)b * m_rows + j;
    if (lane == 0) {
        DrawGeom d;
        d.xmin = g.xmin; d.xmax = g.xmax; d.ymin = g.ymin; d.ymax = g.ymax; d.sx = g.sx; d.sy = g.sy;
        d.cls = draw ? cls : -1; d.pad = 0;
        geom[slot] = d;
    }
    if (!draw) return;
    const TileRef tref = tile_ref(S, b, j, m_stride, mh * mw, cls, mh, mw);
    for (int y = 0; y < mh; ++y) {
        const int v = lane < mw ? tref.at(y * mw + lane) : 0;
        const unsigned w = __ballot_sync(0xffffffffu, v != 0);
        if (lane == 0) bits[slot * mh + y] = w;
    }
}

constexpr int kBlkH = 16, kBlkW = 64;     // pixel block of a CTA: 16 rows x 16 threads x 4 pixels
constexpr int kMaxCand = 1024;            // instances touching one block kept in shared memory (indices)
constexpr int kGeomCache = 32;            // ... of which the first ones with their geometry

struct DrawTilesArgs {
    const void* images;
    const DrawGeom* geom;      // [B, m_rows]
    const uint32_t* bits;      // [B, m_rows, mh]
    const int32_t* m_used;     // [1]
    const void* seg;           // [B, PH, PW, Cs] or NULL
    int B, m_rows, mh, mw, PH, PW;
    mlp_draw_colors inst, sem;
    uint8_t* out;
    const uint32_t* line_bits; // [B, PH, line_words] pixels on DrawBoxes' rectangles (1 bit each), or NULL
    int line_words;
};

// 4 consecutive channel-interleaved pixels (12 values) of a frame row as float
__device__ __forceinline__ void load_px12(const uint8_t* p, bool vec, int n, float* v) {
    if (vec) {
        const uint32_t* w = reinterpret_cast<const uint32_t*>(p);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const uint32_t u = __ldg(w + k);
#pragma unroll
            for (int i = 0; i < 4; ++i) v[k * 4 + i] = (float)((u >> (8 * i)) & 0xffu);
        }
    } else {
#pragma unroll
        for (int i = 0; i < 12; ++i) v[i] = i < 3 * n ? (float)__ldg(p + i) : 0.0f;
    }
}
__device__ __forceinline__ void load_px12(const float* p, bool vec, int n, float* v) {
    if (vec) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float4 f = __ldg(reinterpret_cast<const float4*>(p) + k);
            v[k * 4] = f.x; v[k * 4 + 1] = f.y; v[k * 4 + 2] = f.z; v[k * 4 + 3] = f.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < 12; ++i) v[i] = i < 3 * n ? __ldg(p + i) : 0.0f;
    }
}
__device__ __forceinline__ float seg_f32(const int32_t* p) { return (float)__ldg(p); }
__device__ __forceinline__ float seg_f32(const float* p) { return __ldg(p); }
template <typename T> __device__ __forceinline__ float seg_word(uint32_t w);          // one element of a 128-bit load
template <> __device__ __forceinline__ float seg_word<float>(uint32_t w) { return __uint_as_float(w); }
template <> __device__ __forceinline__ float seg_word<int32_t>(uint32_t w) { return (float)(int32_t)w; }

// kCs: semantic classes known at compile time (0 = no semantic overlay, 3 = the serving graph's three
// classes with their colours in registers and 128-bit map loads, -1 = any number, generic loop).
template <typename ImgT, typename SegT, int kCs>
__global__ void __launch_bounds__(kDrawThreads)
draw_tiles_kernel(const DrawTilesArgs A) {
    __shared__ unsigned short s_cand[kMaxCand];            // instances touching the block, instance order
    __shared__ signed char s_cls[kMaxCand];
    __shared__ DrawGeom s_geom[kGeomCache];
    __shared__ int s_wcnt[kDrawThreads / 32];
    __shared__ unsigned char s_ord[kGeomCache];
    __shared__ int s_start[MLP_MAX_DRAW_CLASSES + 1];
    const int b = blockIdx.z;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int M = min(*A.m_used, A.m_rows);
    const int by0 = blockIdx.y * kBlkH, bx0 = blockIdx.x * kBlkW;
    const DrawGeom* G = A.geom + (int64_t)b * A.m_rows;
    const int C = A.inst.num_classes;
    // instances whose clipped box touches this block, in instance order: 256 instances per round,
    // one per thread, ordered compaction with warp ballots + a prefix over the 8 warp counts
    int ncand = 0;
    for (int base = 0; base < M; base += kDrawThreads) {
        const int j = base + tid;
        bool hit = false;
        DrawGeom d;
        if (j < M) {
            d = G[j];
            hit = d.cls >= 0 && d.ymin < by0 + kBlkH && d.ymax > by0 && d.xmin < bx0 + kBlkW && d.xmax > bx0;
        }
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (lane == 0) s_wcnt[warp] = __popc(m);
        __syncthreads();
        int pre = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < kDrawThreads / 32; ++w) {
            const int v = s_wcnt[w];
            if (w < warp) pre += v;
            tot += v;
        }
        const int k = ncand + pre + __popc(m & ((1u << lane) - 1u));
        if (hit && k < kMaxCand) {
            s_cand[k] = (unsigned short)j;
            s_cls[k] = (signed char)d.cls;
            if (k < kGeomCache) s_geom[k] = d;
        }
        ncand += tot;
        __syncthreads();
    }
    // At most 32 candidates (= the geometry cache): warp 0 groups them by class with ballots, order kept, so that
    // the per-class loops below touch only their own candidates instead of scanning the list once per class.
    const bool grouped = ncand > 0 && ncand <= kGeomCache && C <= MLP_MAX_DRAW_CLASSES;
    if (grouped) {
        if (warp == 0) {
            const int mine = lane < ncand ? (int)s_cls[lane] : -1;
            int start = 0;
            for (int c = 0; c < C; ++c) {
                const unsigned m = __ballot_sync(0xffffffffu, mine == c);
                if (mine == c) s_ord[start + __popc(m & ((1u << lane) - 1u))] = (unsigned char)lane;
                if (lane == 0) s_start[c] = start;
                start += __popc(m);
            }
            if (lane == 0) s_start[C] = start;
        }
        __syncthreads();
    }
    const int oy = by0 + (tid >> 4), ox = bx0 + (tid & 15) * 4;
    if (oy >= A.PH || ox >= A.PW) return;
    const int mh = A.mh, mw = A.mw;
    float cs[4][3];
#pragma unroll
    for (int q = 0; q < 4; ++q) cs[q][0] = cs[q][1] = cs[q][2] = 0.0f;
    if (ncand > 0) {
        // the float32 value CropAndPadMask writes at (oy, ox+q) for instance j: two-stage lerp of the {0,1} tile
        auto eval = [&](int j, const DrawGeom& d, float (&acc)[4]) {
            const float py = __fmul_rn((float)(oy - d.ymin), d.sy);
            const float fy = floorf(py);
            const float ly = __fsub_rn(py, fy);
            const uint32_t* tb = A.bits + ((int64_t)b * A.m_rows + j) * mh;
            const uint32_t w0 = __ldg(tb + max((int)fy, 0)), w1 = __ldg(tb + min((int)ceilf(py), mh - 1));
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int x = ox + q;
                if (x < d.xmin || x >= d.xmax) continue;
                const float p = __fmul_rn((float)(x - d.xmin), d.sx);
                const float fl = floorf(p);
                const int xlo = max((int)fl, 0), xhi = min((int)ceilf(p), mw - 1);
                const float lx = __fsub_rn(p, fl);
                const float tl = (float)((w0 >> xlo) & 1u), tr = (float)((w0 >> xhi) & 1u);
                const float bl = (float)((w1 >> xlo) & 1u), br = (float)((w1 >> xhi) & 1u);
                const float t = __fadd_rn(tl, __fmul_rn(__fsub_rn(tr, tl), lx));
                const float bo = __fadd_rn(bl, __fmul_rn(__fsub_rn(br, bl), lx));
                acc[q] = __fadd_rn(acc[q], __fadd_rn(t, __fmul_rn(__fsub_rn(bo, t), ly)));
            }
        };
        for (int c = 0; c < C; ++c) {
            if (grouped && s_start[c] == s_start[c + 1]) continue;      // no instance of this class near the block
            float acc[4] = {0.f, 0.f, 0.f, 0.f};           // reduce_sum of the class's masks, order j
            if (grouped) {
                // the usual case (at most 32 candidates): only this class's candidates, still in instance order
                for (int k = s_start[c]; k < s_start[c + 1]; ++k) {
                    const int i = s_ord[k];
                    const DrawGeom d = s_geom[i];
                    if (oy < d.ymin || oy >= d.ymax || ox + 3 < d.xmin || ox >= d.xmax) continue;
                    eval((int)s_cand[i], d, acc);
                }
            } else {
                // list overflow: walk all instances of the image instead (same order, same result)
                const bool all = ncand > kMaxCand;
                const int i1 = all ? M : ncand;
                for (int i = 0; i < i1; ++i) {
                    if (!all && s_cls[i] != c) continue;
                    const int j = all ? i : (int)s_cand[i];
                    const DrawGeom d = (!all && i < kGeomCache) ? s_geom[i] : G[j];
                    if (d.cls != c || oy < d.ymin || oy >= d.ymax || ox + 3 < d.xmin || ox >= d.xmax) continue;
                    eval(j, d, acc);
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (acc[q] > 0.5f)
#pragma unroll
                    for (int k = 0; k < 3; ++k) cs[q][k] = __fadd_rn(cs[q][k], A.inst.rgb[c][k]);
        }
    }
    // ---- blend: DrawInstance's uint8, then (optionally) DrawSegmentation over it (serving.py:38-40)
    const int64_t pix0 = ((int64_t)b * A.PH + oy) * A.PW + ox;
    const int n = min(4, A.PW - ox);
    const bool vec = n == 4 && (A.PW & 3) == 0;            // 4-pixel groups are 12-byte / 48-byte aligned
    float img[12];
    load_px12(static_cast<const ImgT*>(A.images) + pix0 * 3, vec, n, img);
    if (A.line_bits) {
        // DrawBoxes in front of the overlays (serving.py:34): its uint8 canvas (clip + truncation for float frames)
        // with the rectangles' pixels at 255; ox is a multiple of 4, so the four pixels share one word of the bitmap
        if (sizeof(ImgT) != 1) {                           // uint8 frames are their own canvas
#pragma unroll
            for (int i = 0; i < 12; ++i) img[i] = (float)__float2uint_rz(fminf(fmaxf(img[i], 0.0f), 255.0f));
        }
        const uint32_t lw = __ldg(A.line_bits + ((int64_t)b * A.PH + oy) * A.line_words + (ox >> 5)) >> (ox & 31);
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if ((lw >> q) & 1u) img[q * 3] = img[q * 3 + 1] = img[q * 3 + 2] = 255.0f;
    }
    float sc[4][3];
#pragma unroll
    for (int q = 0; q < 4; ++q) sc[q][0] = sc[q][1] = sc[q][2] = 0.0f;
    if (kCs == 3) {
        float rgb[3][3];
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int k = 0; k < 3; ++k) rgb[c][k] = A.sem.rgb[c][k];
        const SegT* sp = static_cast<const SegT*>(A.seg) + pix0 * 3;
        float sv[12];
        if (vec) {                                         // 4 pixels x 3 classes = three 128-bit loads
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const uint4 u = __ldg(reinterpret_cast<const uint4*>(sp) + k);
                const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    sv[k * 4 + e] = seg_word<SegT>(w[e]);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 12; ++i) sv[i] = i < 3 * n ? seg_f32(sp + i) : 0.0f;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int k = 0; k < 3; ++k) sc[q][k] = __fadd_rn(sc[q][k], __fmul_rn(rgb[c][k], sv[q * 3 + c]));
    } else if (kCs != 0) {
        const int Cs = A.sem.num_classes;
        const SegT* sp = static_cast<const SegT*>(A.seg) + pix0 * Cs;
        for (int q = 0; q < n; ++q)
            for (int c = 0; c < Cs; ++c) {
                const float v = seg_f32(sp + q * Cs + c);
#pragma unroll
                for (int k = 0; k < 3; ++k) sc[q][k] = __fadd_rn(sc[q][k], __fmul_rn(A.sem.rgb[c][k], v));
            }
    }
    uint32_t word[3] = {0u, 0u, 0u};
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float v = (float)__float2uint_rz(blend(img[q * 3 + k], cs[q][k], A.inst.alpha));
            if (kCs != 0) v = (float)__float2uint_rz(blend(v, sc[q][k], A.sem.alpha));
            const int i = q * 3 + k;
            word[i >> 2] |= __float2uint_rz(v) << (8 * (i & 3));
        }
    uint8_t* op = A.out + pix0 * 3;
    if (vec) {
#pragma unroll
        for (int k = 0; k < 3; ++k) reinterpret_cast<uint32_t*>(op)[k] = word[k];
    } else {
        for (int i = 0; i < 3 * n; ++i) op[i] = (uint8_t)(word[i >> 2] >> (8 * (i & 3)));
    }
}

// ---- DrawBoxes --------------------------------------------------------------------------------
// clip(float image, 0, 255) -> uint8 (the canvas before the rectangles; identity for uint8 frames)
__global__ void __launch_bounds__(kDrawThreads)
canvas_kernel(const float* __restrict__ images, int64_t n, uint8_t* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * kDrawThreads + threadIdx.x;
    if (i < n) out[i] = (uint8_t)__float2uint_rz(fminf(fmaxf(__ldg(images + i), 0.0f), 255.0f));
}

// DrawBoxes.call (misc.py:486-497) corner arithmetic and the rectangle tf.image.draw_bounding_boxes draws for it
// (draw_bounding_box_op.cc, restated in oracle/draw_oracle.py); false: nothing to draw.
struct BoxRect { int y0, y1, x0, x1; };
__device__ __forceinline__ bool box_rect(const int32_t* row, int H, int W, BoxRect& r) {
    const float cx = (float)max(row[0], 0), cy = (float)max(row[1], 0);       // tf.maximum(det[..., :4], 0)
    const float w = (float)max(row[2], 0), h = (float)max(row[3], 0);
    const float fh = (float)H, fw = (float)W;
    const float hw = __fdiv_rn(w, 2.0f), hh = __fdiv_rn(h, 2.0f);
    const float xmin = __fdiv_rn(__fsub_rn(cx, hw), fw), xmax = __fdiv_rn(__fadd_rn(cx, hw), fw);
    const float ymin = __fdiv_rn(__fsub_rn(cy, hh), fh), ymax = __fdiv_rn(__fadd_rn(cy, hh), fh);
    // box * (size - 1), converted to an integer by C++ truncation (toward zero)
    r.y0 = __float2int_rz(__fmul_rn(ymin, (float)(H - 1))); r.y1 = __float2int_rz(__fmul_rn(ymax, (float)(H - 1)));
    r.x0 = __float2int_rz(__fmul_rn(xmin, (float)(W - 1))); r.x1 = __float2int_rz(__fmul_rn(xmax, (float)(W - 1)));
    if (r.y0 > r.y1 || r.x0 > r.x1) return false;
    if (r.y0 >= H || r.y1 < 0 || r.x0 >= W || r.x1 < 0) return false;
    return true;
}

// One warp per box: the four one-pixel lines, written into the frame.
__global__ void __launch_bounds__(kDrawThreads)
draw_boxes_kernel(const int32_t* __restrict__ det, int B, int m_rows, int m_stride, const int32_t* __restrict__ m_dev,
                  int H, int W, uint8_t* __restrict__ out) {
    int M = m_dev ? *m_dev : m_rows;
    if (M > m_rows) M = m_rows;
    if (m_stride == 0) m_stride = M;
    const int lane = threadIdx.x & 31;
    const int64_t box = (int64_t)blockIdx.x * (kDrawThreads / 32) + (threadIdx.x >> 5);
    if (box >= (int64_t)B * M) return;
    const int b = (int)(box / M), j = (int)(box - (int64_t)b * M);
    BoxRect r;
    if (!box_rect(det + ((int64_t)b * m_stride + j) * 6, H, W, r)) return;
    const int y0 = r.y0, y1 = r.y1, x0 = r.x0, x1 = r.x1;
    const int y0c = max(y0, 0), y1c = min(y1, H - 1), x0c = max(x0, 0), x1c = min(x1, W - 1);
    uint8_t* img = out + (int64_t)b * H * W * 3;
    for (int x = x0c + lane; x <= x1c; x += 32) {
        if (y0 >= 0) { uint8_t* p = img + ((int64_t)y0 * W + x) * 3; p[0] = p[1] = p[2] = 255; }
        if (y1 < H) { uint8_t* p = img + ((int64_t)y1 * W + x) * 3; p[0] = p[1] = p[2] = 255; }
    }
    for (int y = y0c + lane; y <= y1c; y += 32) {
        if (x0 >= 0) { uint8_t* p = img + ((int64_t)y * W + x0) * 3; p[0] = p[1] = p[2] = 255; }
        if (x1 < W) { uint8_t* p = img + ((int64_t)y * W + x1) * 3; p[0] = p[1] = p[2] = 255; }
    }
}

// The same lines as one bit per pixel ([B, H, words], zeroed before the launch) for draw_tiles_kernel: the frame is
// then neither copied nor written before the overlay pass.  Horizontal lines are word-wide ORs.
__global__ void __launch_bounds__(kDrawThreads)
box_lines_kernel(const int32_t* __restrict__ det, const PasteSrc S, int B, int m_rows, int m_stride, int H, int W,
                 int words, uint32_t* __restrict__ bits) {
    int M, thr;
    paste_scalars(S, B, m_rows, M, thr);                  // the rows the tail kept (M on the device)
    if (m_stride == 0) m_stride = M;
    const int lane = threadIdx.x & 31;
    const int64_t box = (int64_t)blockIdx.x * (kDrawThreads / 32) + (threadIdx.x >> 5);
    if (box >= (int64_t)B * M) return;
    const int b = (int)(box / M), j = (int)(box - (int64_t)b * M);
    BoxRect r;
    if (!box_rect(det + ((int64_t)b * m_stride + j) * 6, H, W, r)) return;
    const int y0c = max(r.y0, 0), y1c = min(r.y1, H - 1), x0c = max(r.x0, 0), x1c = min(r.x1, W - 1);
    uint32_t* img = bits + (int64_t)b * H * words;
    for (int wi = (x0c >> 5) + lane; wi <= (x1c >> 5); wi += 32) {
        const int lo = max(x0c - wi * 32, 0), hi = min(x1c - wi * 32, 31);
        const uint32_t m = (hi == 31 ? 0xffffffffu : ((1u << (hi + 1)) - 1u)) & ~((1u << lo) - 1u);
        if (r.y0 >= 0) atomicOr(img + (int64_t)r.y0 * words + wi, m);
        if (r.y1 < H) atomicOr(img + (int64_t)r.y1 * words + wi, m);
    }
    for (int y = y0c + lane; y <= y1c; y += 32) {
        if (r.x0 >= 0) atomicOr(img + (int64_t)y * words + (r.x0 >> 5), 1u << (r.x0 & 31));
        if (r.x1 < W) atomicOr(img + (int64_t)y * words + (r.x1 >> 5), 1u << (r.x1 & 31));
    }
}

int check_colors(const char* who, const mlp_draw_colors* c) {
    MLP_CHECK_ARG(c && c->num_classes >= 1 && c->num_classes <= MLP_MAX_DRAW_CLASSES,
                  "%s: colour table needs 1..%d classes", who, MLP_MAX_DRAW_CLASSES);
    return MLP_OK;
}

}  // namespace

// ================================================================ host side ===
extern "C" int mlp_draw_segmentation(mlp_ctx* ctx, const void* images_dev, int image_dtype, const void* seg_dev,
                                     int seg_dtype, int batch, int frame_h, int frame_w,
                                     const mlp_draw_colors* colors, uint8_t* out_dev, mlp_stream_t stream) {
    MLP_CHECK_ARG(ctx && images_dev && seg_dev && out_dev, "mlp_draw_segmentation: NULL argument");
    MLP_CHECK_ARG(batch >= 1 && frame_h >= 1 && frame_w >= 1, "mlp_draw_segmentation: bad shape");
    MLP_CHECK_ARG(image_dtype == MLP_U8 || image_dtype == MLP_F32, "mlp_draw_segmentation: images must be u8 or f32");
    MLP_CHECK_ARG(seg_dtype == MLP_I32 || seg_dtype == MLP_F32, "mlp_draw_segmentation: seg must be i32 or f32");
    int rc = check_colors("mlp_draw_segmentation", colors);
    if (rc) return rc;
    DeviceGuard g(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope prof(ctx, MLP_ST_DRAW, st);
    const int64_t npix = (int64_t)batch * frame_h * frame_w;
    const int grid = (int)((npix + kDrawThreads - 1) / kDrawThreads);
#define MLP_DRAW_SEG(IT, ST)                                                                             \
    draw_segmentation_kernel<IT, ST><<<grid, kDrawThreads, 0, st>>>(static_cast<const IT*>(images_dev), \
                                                                    static_cast<const ST*>(seg_dev), npix, *colors, out_dev)
    if (image_dtype == MLP_U8 && seg_dtype == MLP_I32) MLP_DRAW_SEG(uint8_t, int32_t);
    else if (image_dtype == MLP_U8) MLP_DRAW_SEG(uint8_t, float);
    else if (seg_dtype == MLP_I32) MLP_DRAW_SEG(float, int32_t);
    else MLP_DRAW_SEG(float, float);
#undef MLP_DRAW_SEG
    MLP_LAUNCH_CHECK(ctx);
    return MLP_OK;
}

extern "C" int mlp_draw_instance(mlp_ctx* ctx, const void* images_dev, int image_dtype, const int32_t* det_i32_dev,
                                 const void* masks_dev, int mask_dtype, int batch, int m_rows, int m_stride,
                                 const int32_t* m_dev, int frame_h, int frame_w, const mlp_draw_colors* colors,
                                 uint8_t* out_dev, mlp_stream_t stream) {
    MLP_CHECK_ARG(ctx && images_dev && det_i32_dev && masks_dev && out_dev, "mlp_draw_instance: NULL argument");
    MLP_CHECK_ARG(batch >= 1 && batch <= 65535 && frame_h >= 1 && frame_w >= 1 && m_rows >= 1 &&
                      m_rows <= kMaxDrawInst && (m_stride >= m_rows || (m_stride == 0 && m_dev)),
                  "mlp_draw_instance: bad shape");
    MLP_CHECK_ARG((int64_t)frame_h * frame_w < (1ll << 30), "mlp_draw_instance: frame too large");
    MLP_CHECK_ARG(image_dtype == MLP_U8 || image_dtype == MLP_F32, "mlp_draw_instance: images must be u8 or f32");
    MLP_CHECK_ARG(mask_dtype == MLP_U8 || mask_dtype == MLP_F32, "mlp_draw_instance: masks must be u8 or f32");
    MLP_CHECK_ARG(mlp_aligned16(masks_dev), "mlp_draw_instance: masks must be 16-byte aligned");
    int rc = check_colors("mlp_draw_instance", colors);
    if (rc) return rc;
    DeviceGuard g(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope prof(ctx, MLP_ST_DRAW, st);
    const int npx = frame_h * frame_w;
    const int ppt = mask_dtype == MLP_U8 ? 16 : 4;          // pixels per thread: 16 bytes of mask
    const dim3 grid((npx + kDrawThreads * ppt - 1) / (kDrawThreads * ppt), batch);
#define MLP_DRAW_INST(IT, MT)                                                                               \
    draw_instance_kernel<IT, MT><<<grid, kDrawThreads, 0, st>>>(static_cast<const IT*>(images_dev), det_i32_dev, \
                                                                static_cast<const MT*>(masks_dev), m_rows,   \
                                                                m_stride, m_dev, npx, *colors, out_dev)
    if (image_dtype == MLP_U8 && mask_dtype == MLP_U8) MLP_DRAW_INST(uint8_t, uint8_t);
    else if (image_dtype == MLP_U8) MLP_DRAW_INST(uint8_t, float);
    else if (mask_dtype == MLP_U8) MLP_DRAW_INST(float, uint8_t);
    else MLP_DRAW_INST(float, float);
#undef MLP_DRAW_INST
    MLP_LAUNCH_CHECK(ctx);
    return MLP_OK;
}

static int draw_tiles_impl(mlp_ctx* ctx, const void* images_dev, int image_dtype, const int32_t* det_i32_dev,
                           const int32_t* masks_i32_dev, const float* roi_masks_dev, int r_rows,
                           const int32_t* r_dev, int num_classes, const int32_t* counts_dev, int batch,
                           int m_rows, int m_stride, const int32_t* m_dev, int mask_h, int mask_w, int frame_h,
                           int frame_w, const mlp_draw_colors* inst_colors, const void* seg_dev, int seg_dtype,
                           const mlp_draw_colors* sem_colors, uint8_t* out_dev, mlp_stream_t stream, bool boxes) {
    MLP_CHECK_ARG(ctx && images_dev && det_i32_dev && out_dev, "mlp_draw_tiles: NULL argument");
    MLP_CHECK_ARG(masks_i32_dev || (roi_masks_dev && counts_dev && num_classes >= 1 && r_rows >= 1),
                  "mlp_draw_tiles: neither int32 tiles nor a prepared fused tail");
    MLP_CHECK_ARG(batch >= 1 && batch <= 65535 && m_rows >= 1 && m_rows <= 65535 && frame_h >= 1 && frame_w >= 1 &&
                      (m_stride >= m_rows || (m_stride == 0 && m_dev)),
                  "mlp_draw_tiles: bad shape");
    MLP_CHECK_ARG(mask_h >= 1 && mask_w >= 1 && mask_w <= 32 && mask_h * mask_w <= kMaxTile,
                  "mlp_draw_tiles: mask tile %dx%d (rows of at most 32 columns)", mask_h, mask_w);
    MLP_CHECK_ARG(image_dtype == MLP_U8 || image_dtype == MLP_F32, "mlp_draw_tiles: images must be u8 or f32");
    MLP_CHECK_ARG((seg_dev != nullptr) == (sem_colors != nullptr), "mlp_draw_tiles: seg_dev and sem_colors go together");
    MLP_CHECK_ARG(!seg_dev || seg_dtype == MLP_I32 || seg_dtype == MLP_F32, "mlp_draw_tiles: seg must be i32 or f32");
    int rc = check_colors("mlp_draw_tiles", inst_colors);
    if (rc) return rc;
    if (sem_colors && (rc = check_colors("mlp_draw_tiles", sem_colors))) return rc;
    DeviceGuard g(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope prof(ctx, MLP_ST_DRAW, st);
    PasteSrc S;
    memset(&S, 0, sizeof(S));
    if (masks_i32_dev) {
        int32_t* thr_dev = ctx->ctr;        // ctr[0]: paste row-filter threshold
        paste_threshold_kernel<<<1, 1024, 0, st>>>(det_i32_dev, batch, m_rows, m_stride, m_dev, thr_dev);
        MLP_LAUNCH_CHECK(ctx);
        S.masks_i32 = masks_i32_dev;
        S.m_dev = m_dev;
        S.thr_dev = thr_dev;
    } else {
        MLP_CHECK_ARG(m_stride == m_rows, "mlp_draw_tiles: the fused tail uses capacity rows (m_stride == m_rows)");
        const FusedTail need = fused_tail_layout(nullptr, batch, m_rows, mask_h, mask_w);
        MLP_CHECK_ARG(ctx->arena[MLP_ARENA_FUSED] && ctx->arena_bytes[MLP_ARENA_FUSED] >= need.bytes,
                      "mlp_draw_tiles: call mlp_trim_paste with the same shapes first");
        const FusedTail ft = fused_tail_layout(ctx->arena[MLP_ARENA_FUSED], batch, m_rows, mask_h, mask_w);
        S.fused = 1;
        S.roi_masks = roi_masks_dev;
        S.tail_src = ft.tail_src;
        S.tail_bits = ft.tail_bits;
        S.r_dev = r_dev;
        S.r_rows = r_rows;
        S.C = num_classes;
        S.planar = ctx->tail_planar;
        S.counts = counts_dev;
        S.confmax = ft.confmax;
        S.scalars = ft.scalars;
    }
    // scratch: geometry [B,m_rows] + bit rows [B,m_rows,mh] + M
    const int64_t n_inst = (int64_t)batch * m_rows;
    // (+ with boxes: one bit per frame pixel for DrawBoxes' rectangles)
    const int line_words = (frame_w + 31) / 32;
    const int64_t line_bytes = boxes ? (int64_t)batch * frame_h * line_words * 4 : 0;
    const int64_t bytes = n_inst * (int64_t)sizeof(DrawGeom) + n_inst * mask_h * 4 + 16 + line_bytes;
    rc = mlp_ensure_scratch(ctx, MLP_ARENA_DRAW, bytes);
    if (rc) return rc;
    DrawGeom* geom = static_cast<DrawGeom*>(ctx->arena[MLP_ARENA_DRAW]);
    uint32_t* bits = reinterpret_cast<uint32_t*>(geom + n_inst);
    int32_t* m_used = reinterpret_cast<int32_t*>(bits + n_inst * mask_h);
    uint32_t* line_bits = boxes ? reinterpret_cast<uint32_t*>(m_used + 4) : nullptr;
    const int wpc = kDrawThreads / 32;
    if (boxes) {
        MLP_CUDA(cudaMemsetAsync(line_bits, 0, (size_t)line_bytes, st));
        box_lines_kernel<<<(int)((n_inst + wpc - 1) / wpc), kDrawThreads, 0, st>>>(det_i32_dev, S, batch, m_rows, m_stride,
                                                                                  frame_h, frame_w, line_words, line_bits);
        MLP_LAUNCH_CHECK(ctx);
    }
    pack_tiles_kernel<<<(int)((n_inst + wpc - 1) / wpc), kDrawThreads, 0, st>>>(
        det_i32_dev, S, batch, m_rows, m_stride, mask_h, mask_w, frame_h, frame_w, inst_colors->num_classes, geom,
        bits, m_used);
    MLP_LAUNCH_CHECK(ctx);
    DrawTilesArgs A;
    memset(&A, 0, sizeof(A));
    A.images = images_dev; A.geom = geom; A.bits = bits; A.m_used = m_used; A.seg = seg_dev; A.B = batch;
    A.m_rows = m_rows; A.mh = mask_h; A.mw = mask_w; A.PH = frame_h; A.PW = frame_w; A.inst = *inst_colors;
    if (sem_colors) A.sem = *sem_colors;
    A.out = out_dev;
    A.line_bits = line_bits; A.line_words = line_words;
    const dim3 grid((frame_w + kBlkW - 1) / kBlkW, (frame_h + kBlkH - 1) / kBlkH, batch);
    const bool seg_f = seg_dev && seg_dtype == MLP_F32;
    const int cs = !seg_dev ? 0 : (sem_colors->num_classes == 3 ? 3 : -1);
#define MLP_DRAW_TILES(IT, ST)                                                                  \
    do {                                                                                        \
        if (cs == 0) draw_tiles_kernel<IT, ST, 0><<<grid, kDrawThreads, 0, st>>>(A);            \
        else if (cs == 3) draw_tiles_kernel<IT, ST, 3><<<grid, kDrawThreads, 0, st>>>(A);       \
        else draw_tiles_kernel<IT, ST, -1><<<grid, kDrawThreads, 0, st>>>(A);                   \
    } while (0)
    if (image_dtype == MLP_U8 && !seg_f) MLP_DRAW_TILES(uint8_t, int32_t);
    else if (image_dtype == MLP_U8) MLP_DRAW_TILES(uint8_t, float);
    else if (!seg_f) MLP_DRAW_TILES(float, int32_t);
    else MLP_DRAW_TILES(float, float);
#undef MLP_DRAW_TILES
    MLP_LAUNCH_CHECK(ctx);
    return MLP_OK;
}

extern "C" int mlp_draw_tiles(mlp_ctx* ctx, const void* images_dev, int image_dtype, const int32_t* det_i32_dev,
                              const int32_t* masks_i32_dev, const float* roi_masks_dev, int r_rows,
                              const int32_t* r_dev, int num_classes, const int32_t* counts_dev, int batch,
                              int m_rows, int m_stride, const int32_t* m_dev, int mask_h, int mask_w, int frame_h,
                              int frame_w, const mlp_draw_colors* inst_colors, const void* seg_dev, int seg_dtype,
                              const mlp_draw_colors* sem_colors, uint8_t* out_dev, mlp_stream_t stream) {
    return draw_tiles_impl(ctx, images_dev, image_dtype, det_i32_dev, masks_i32_dev, roi_masks_dev, r_rows, r_dev,
                           num_classes, counts_dev, batch, m_rows, m_stride, m_dev, mask_h, mask_w, frame_h, frame_w,
                           inst_colors, seg_dev, seg_dtype, sem_colors, out_dev, stream, false);
}

extern "C" int mlp_draw_tiles_boxes(mlp_ctx* ctx, const void* images_dev, int image_dtype, const int32_t* det_i32_dev,
                                    const int32_t* masks_i32_dev, const float* roi_masks_dev, int r_rows,
                                    const int32_t* r_dev, int num_classes, const int32_t* counts_dev, int batch,
                                    int m_rows, int m_stride, const int32_t* m_dev, int mask_h, int mask_w,
                                    int frame_h, int frame_w, const mlp_draw_colors* inst_colors, const void* seg_dev,
                                    int seg_dtype, const mlp_draw_colors* sem_colors, uint8_t* out_dev,
                                    mlp_stream_t stream) {
    return draw_tiles_impl(ctx, images_dev, image_dtype, det_i32_dev, masks_i32_dev, roi_masks_dev, r_rows, r_dev,
                           num_classes, counts_dev, batch, m_rows, m_stride, m_dev, mask_h, mask_w, frame_h, frame_w,
                           inst_colors, seg_dev, seg_dtype, sem_colors, out_dev, stream, true);
}

extern "C" int mlp_draw_boxes(mlp_ctx* ctx, const void* images_dev, int image_dtype, const int32_t* det_i32_dev,
                              int batch, int m_rows, int m_stride, const int32_t* m_dev, int frame_h, int frame_w,
                              uint8_t* out_dev, mlp_stream_t stream) {
    MLP_CHECK_ARG(ctx && images_dev && det_i32_dev && out_dev, "mlp_draw_boxes: NULL argument");
    MLP_CHECK_ARG(batch >= 1 && m_rows >= 1 && frame_h >= 1 && frame_w >= 1 &&
                      (m_stride >= m_rows || (m_stride == 0 && m_dev)),
                  "mlp_draw_boxes: bad shape");
    MLP_CHECK_ARG(image_dtype == MLP_U8 || image_dtype == MLP_F32, "mlp_draw_boxes: images must be u8 or f32");
    DeviceGuard g(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope prof(ctx, MLP_ST_DRAW, st);
    const int64_t n = (int64_t)batch * frame_h * frame_w * 3;
    if (image_dtype == MLP_U8) {
        if (images_dev != out_dev)
            MLP_CUDA(cudaMemcpyAsync(out_dev, images_dev, (size_t)n, cudaMemcpyDeviceToDevice, st));
    } else {
        canvas_kernel<<<(int)((n + kDrawThreads - 1) / kDrawThreads), kDrawThreads, 0, st>>>(
            static_cast<const float*>(images_dev), n, out_dev);
        MLP_LAUNCH_CHECK(ctx);
    }
    const int64_t boxes = (int64_t)batch * m_rows;
    const int wpc = kDrawThreads / 32;
    draw_boxes_kernel<<<(int)((boxes + wpc - 1) / wpc), kDrawThreads, 0, st>>>(det_i32_dev, batch, m_rows, m_stride,
                                                                            m_dev, frame_h, frame_w, out_dev);
    MLP_LAUNCH_CHECK(ctx);
    return MLP_OK;
}
