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

#ifndef MLP_DRAW_MIN_CTAS
#define MLP_DRAW_MIN_CTAS 4
#endif
#ifndef MLP_DRAW_ROWS
#define MLP_DRAW_ROWS 4
#endif
constexpr int kRowsPT = MLP_DRAW_ROWS;    // adjacent frame rows per thread
constexpr int kBlkH = 16 * kRowsPT, kBlkW = 64;     // pixel block of a CTA: 16 x 16 threads, kRowsPT rows x 4 pixels each
constexpr int kMaxCand = 1024;            // instances touching one block kept in shared memory (indices)
constexpr int kGeomCache = 64;            // ... of which the first ones with their geometry (the fast path)

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

// 4 consecutive channel-interleaved pixels (12 values) of a frame row: the raw words of the load (issued early) and
// their float values (taken when the arithmetic starts)
template <typename ImgT> struct PxRaw;
template <> struct PxRaw<uint8_t> {
    uint32_t u[3];
    __device__ __forceinline__ void load(const uint8_t* p, bool vec, int n) {
        if (vec) {
#pragma unroll
            for (int k = 0; k < 3; ++k) u[k] = __ldg(reinterpret_cast<const uint32_t*>(p) + k);
        } else {
            u[0] = u[1] = u[2] = 0u;
#pragma unroll
            for (int i = 0; i < 12; ++i)
                if (i < 3 * n) u[i >> 2] |= (uint32_t)__ldg(p + i) << (8 * (i & 3));
        }
    }
    __device__ __forceinline__ void pixels(float* v, bool) const {
#pragma unroll
        for (int k = 0; k < 3; ++k)
#pragma unroll
            for (int i = 0; i < 4; ++i)                    // byte -> float as (2^23 + byte) - 2^23: no int->float conversion
                v[k * 4 + i] = __fsub_rn(__uint_as_float(__byte_perm(u[k], 0x4B000000u, 0x7650u + i)), 8388608.0f);
    }
};
template <> struct PxRaw<float> {
    float f[12];
    __device__ __forceinline__ void load(const float* p, bool vec, int n) {
        if (vec) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const float4 t = __ldg(reinterpret_cast<const float4*>(p) + k);
                f[k * 4] = t.x; f[k * 4 + 1] = t.y; f[k * 4 + 2] = t.z; f[k * 4 + 3] = t.w;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 12; ++i) f[i] = i < 3 * n ? __ldg(p + i) : 0.0f;
        }
    }
    __device__ __forceinline__ void pixels(float* v, bool) const {
#pragma unroll
        for (int i = 0; i < 12; ++i) v[i] = f[i];
    }
};
template <typename T> __device__ __forceinline__ constexpr bool is_float_seg() { return false; }
template <> __device__ __forceinline__ constexpr bool is_float_seg<float>() { return true; }
__device__ __forceinline__ float clip255(float v) { return fminf(fmaxf(v, 0.0f), 255.0f); }   // clip_by_value(., 0, 255)
__device__ __forceinline__ float seg_f32(const int32_t* p) { return (float)__ldg(p); }
__device__ __forceinline__ float seg_f32(const float* p) { return __ldg(p); }
template <typename T> __device__ __forceinline__ float seg_word(uint32_t w);          // one element of a 128-bit load
template <> __device__ __forceinline__ float seg_word<float>(uint32_t w) { return __uint_as_float(w); }
template <> __device__ __forceinline__ float seg_word<int32_t>(uint32_t w) { return (float)(int32_t)w; }

// A value v in [0, 255] truncated toward zero, as the float 2^23 + trunc(v): an add with round-toward-zero against
// 2^23 (ulp 1 there) drops the fraction.  Its low byte is the uint8 the reference's cast produces; subtracting 2^23
// again gives that integer back as a float, exactly - both without a float<->int conversion.
constexpr float kTwo23 = 8388608.0f;
__device__ __forceinline__ float trunc_magic(float v) { return __fadd_rz(v, kTwo23); }

// kCs: semantic classes known at compile time (0 = no semantic overlay, 3 = the serving graph's three
// classes with their colours in registers and 128-bit map loads, -1 = any number, generic loop).
//
// One CTA per 64x64 pixel block; a thread owns 4 adjacent rows x 4 pixels.  The instances whose clipped box touches
// the block are found once per CTA (ordered compaction) and grouped by class; for each of them a thread computes the
// x half of the resize (source columns as single-bit masks, lx) once for its four pixels and walks its four rows with
// it.  The tiles are {0,1} bit rows, so the x lerp tl + (tr - tl) * lx of a row is one of 0, lx, fadd(1, -lx), 1 -
// picked by the two corner bits, bit-identical to the arithmetic.  What a thread keeps per pixel is only WHICH classes
// passed sum > 0.5 (16 bits); the colour sums are rebuilt in class order at blend time.
template <typename ImgT, typename SegT, int kCs>
__global__ void __launch_bounds__(kDrawThreads, MLP_DRAW_MIN_CTAS)
draw_tiles_kernel(const DrawTilesArgs A) {
    __shared__ unsigned short s_cand[kMaxCand];            // instances touching the block, instance order
    __shared__ signed char s_cls[kMaxCand];
    __shared__ DrawGeom s_geom[kGeomCache];
    __shared__ int s_wcnt[kDrawThreads / 32];
    __shared__ unsigned char s_ord[kGeomCache];
    __shared__ int s_start[MLP_MAX_DRAW_CLASSES], s_cnt[MLP_MAX_DRAW_CLASSES];      // a class's run in s_ord: [start, cnt)
    __shared__ unsigned char s_pcls[MLP_MAX_DRAW_CLASSES];                           // classes present near the block
    __shared__ int s_np;
    __shared__ float4 s_ctab[16], s_stab[8];               // colour * alpha per class subset / per {0,1} map triple
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
    // At most 64 candidates (= the geometry cache): warp 0 groups them by class with ballots, order kept, so that
    // the per-class loops below touch only their own candidates instead of scanning the list once per class; the
    // classes that have a candidate at all (usually one to three) form the list the threads walk.  (Grouping the
    // classes in parallel, one warp each, with per-warp colour tables was measured slower: 144 against 117 us.)
    const bool grouped = ncand > 0 && ncand <= kGeomCache;
    if (warp == 0) {
        if (grouped) {
            int start = 0;
            for (int c = 0; c < C; ++c) {
                if (lane == c) s_start[c] = start;
#pragma unroll
                for (int h = 0; h < kGeomCache; h += 32) {
                    const int i = h + lane;
                    const bool mine = i < ncand && (int)s_cls[i] == c;
                    const unsigned m = __ballot_sync(0xffffffffu, mine);
                    if (mine) s_ord[start + __popc(m & ((1u << lane) - 1u))] = (unsigned char)i;
                    start += __popc(m);
                }
                if (lane == c) s_cnt[c] = start;           // end of the class's run
            }
            __syncwarp();
            const bool present = lane < C && s_cnt[lane] > s_start[lane];
            const unsigned pm = __ballot_sync(0xffffffffu, present);
            if (present) s_pcls[__popc(pm & ((1u << lane) - 1u))] = (unsigned char)lane;
            if (lane == 0) s_np = __popc(pm);
        } else {
            if (lane < C) s_pcls[lane] = (unsigned char)lane;
            if (lane == 0) s_np = ncand > 0 ? C : 0;
        }
        __syncwarp();
        // Colour terms of the two blends, tabulated once per CTA with the operations a pixel would perform itself:
        // lanes 0-15: csum * alpha of DrawInstance for every subset of (at most four) present classes, the colours
        // added in class order; lanes 16-23: the same for DrawSegmentation over {0,1}-valued three-class maps.
        if (lane < 16) {
            float c0 = 0.0f, c1 = 0.0f, c2 = 0.0f;
            const int np = s_np;
            for (int p = 0; p < 4 && p < np; ++p)
                if ((lane >> p) & 1) {
                    const int c = s_pcls[p];
                    c0 = __fadd_rn(c0, A.inst.rgb[c][0]); c1 = __fadd_rn(c1, A.inst.rgb[c][1]); c2 = __fadd_rn(c2, A.inst.rgb[c][2]);
                }
            s_ctab[lane] = make_float4(__fmul_rn(c0, A.inst.alpha), __fmul_rn(c1, A.inst.alpha), __fmul_rn(c2, A.inst.alpha), 0.0f);
        } else if (kCs == 3 && lane < 24) {
            const float s0 = (float)(lane & 1), s1 = (float)((lane >> 1) & 1), s2 = (float)((lane >> 2) & 1);
            float t[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                float v = __fmul_rn(A.sem.rgb[0][k], s0);
                v = __fadd_rn(v, __fmul_rn(A.sem.rgb[1][k], s1));
                v = __fadd_rn(v, __fmul_rn(A.sem.rgb[2][k], s2));
                t[k] = __fmul_rn(v, A.sem.alpha);
            }
            s_stab[lane - 16] = make_float4(t[0], t[1], t[2], 0.0f);
        }
    }
    __syncthreads();
    const int P = s_np;                                    // classes to walk, block-uniform; bit p of a pixel's mask
    const float4* ctab = s_ctab;
    const int oy0 = by0 + (tid >> 4) * kRowsPT, ox = bx0 + (tid & 15) * 4;
    if (oy0 >= A.PH || ox >= A.PW) return;
    const int mh = A.mh, mw = A.mw;
    // classes whose summed masks exceed 0.5, 16 bits per pixel: cm[r][q >> 1] bits 16 * (q & 1) + c
    uint32_t cm[kRowsPT][2];
#pragma unroll
    for (int r = 0; r < kRowsPT; ++r) cm[r][0] = cm[r][1] = 0u;
    if (grouped) {
        for (int p = 0; p < P; ++p) {
            const int c = s_pcls[p];
            const int k0 = s_start[c], k1 = s_cnt[c];
            float acc[kRowsPT][4];                         // reduce_sum of the class's masks, order j
#pragma unroll
            for (int r = 0; r < kRowsPT; ++r)
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[r][q] = 0.0f;
            bool any = false;
            for (int k = k0; k < k1; ++k) {
                const int i = s_ord[k];
                const DrawGeom d = s_geom[i];
                if (oy0 + kRowsPT <= d.ymin || oy0 >= d.ymax || ox + 3 < d.xmin || ox >= d.xmax) continue;
                any = true;
                // x half of the resize, once for the four rows
                uint32_t mlo[4], mhi[4];
                float lx[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int x = ox + q;
                    const bool in = x >= d.xmin && x < d.xmax;
                    const float p = __fmul_rn((float)(x - d.xmin), d.sx);
                    const float fl = floorf(p);
                    const int xlo = max((int)fl, 0), xhi = min((int)ceilf(p), mw - 1);
                    lx[q] = __fsub_rn(p, fl);
                    mlo[q] = in ? 1u << (xlo & 31) : 0u;   // outside the box: no corner bit -> the value added is +0
                    mhi[q] = in ? 1u << (xhi & 31) : 0u;
                }
                const uint32_t* tb = A.bits + ((int64_t)b * A.m_rows + (int)s_cand[i]) * mh;
#pragma unroll
                for (int r = 0; r < kRowsPT; ++r) {
                    const int oy = oy0 + r;
                    if (oy < d.ymin || oy >= d.ymax) continue;
                    const float py = __fmul_rn((float)(oy - d.ymin), d.sy);
                    const float fy = floorf(py);
                    const float ly = __fsub_rn(py, fy);
                    const uint32_t w0 = __ldg(tb + max((int)fy, 0)), w1 = __ldg(tb + min((int)ceilf(py), mh - 1));
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const bool tl = (w0 & mlo[q]) != 0u, tr = (w0 & mhi[q]) != 0u;
                        const bool bl = (w1 & mlo[q]) != 0u, br = (w1 & mhi[q]) != 0u;
                        const float oml = __fsub_rn(1.0f, lx[q]);              // fadd(1, (0 - 1) * lx)
                        const float t = tl ? (tr ? 1.0f : oml) : (tr ? lx[q] : 0.0f);
                        const float bo = bl ? (br ? 1.0f : oml) : (br ? lx[q] : 0.0f);
                        acc[r][q] = __fadd_rn(acc[r][q], __fadd_rn(t, __fmul_rn(__fsub_rn(bo, t), ly)));
                    }
                }
            }
            if (any) {
#pragma unroll
                for (int r = 0; r < kRowsPT; ++r)
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (acc[r][q] > 0.5f) cm[r][q >> 1] |= 1u << (16 * (q & 1) + p);
            }
        }
    } else if (ncand > 0) {
        // list overflow (more than 64 boxes touch the block, or more than 1024): one row at a time, every class walks
        // the candidate list (or all instances of the image) - same order, same result
        auto eval = [&](int oy, int j, const DrawGeom& d, float (&acc)[4]) {
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
        const bool all = ncand > kMaxCand;
        const int i1 = all ? M : ncand;
#pragma unroll 1
        for (int r = 0; r < kRowsPT; ++r) {
            const int oy = oy0 + r;
            if (oy >= A.PH) break;
            uint32_t m0 = 0u, m1 = 0u;
            for (int c = 0; c < C && c < MLP_MAX_DRAW_CLASSES; ++c) {
                float acc[4] = {0.f, 0.f, 0.f, 0.f};
                for (int i = 0; i < i1; ++i) {
                    if (!all && s_cls[i] != c) continue;
                    const int j = all ? i : (int)s_cand[i];
                    const DrawGeom d = (!all && i < kGeomCache) ? s_geom[i] : G[j];
                    if (d.cls != c || oy < d.ymin || oy >= d.ymax || ox + 3 < d.xmin || ox >= d.xmax) continue;
                    eval(oy, j, d, acc);
                }
                if (acc[0] > 0.5f) m0 |= 1u << c;
                if (acc[1] > 0.5f) m0 |= 1u << (16 + c);
                if (acc[2] > 0.5f) m1 |= 1u << c;
                if (acc[3] > 0.5f) m1 |= 1u << (16 + c);
            }
#pragma unroll
            for (int rr = 0; rr < kRowsPT; ++rr)
                if (rr == r) { cm[rr][0] = m0; cm[rr][1] = m1; }
        }
    }
    // ---- blend: DrawInstance's uint8, then (optionally) DrawSegmentation over it (serving.py:38-40)
    const int n = min(4, A.PW - ox);
    const bool vec = n == 4 && (A.PW & 3) == 0;            // 4-pixel groups are 12-byte / 48-byte aligned
    const float ia = A.inst.alpha, sa = A.sem.alpha;
    const int nrows = min(kRowsPT, A.PH - oy0);
#pragma unroll 1                                           // one copy of the row body: it has to stay in the instruction cache
    for (int r = 0; r < nrows; ++r) {
        const int oy = oy0 + r;
        // every load of the row first (frame pixels, rectangle bits, semantic map): one round trip, not three
        const int64_t pix0 = ((int64_t)b * A.PH + oy) * A.PW + ox;
        PxRaw<ImgT> raw;
        raw.load(static_cast<const ImgT*>(A.images) + pix0 * 3, vec, n);
        uint32_t lw = 0u;
        if (A.line_bits) lw = __ldg(A.line_bits + ((int64_t)b * A.PH + oy) * A.line_words + (ox >> 5));
        uint32_t w[12];                                    // kCs == 3: the raw words of the map, 4 pixels x 3 classes
        if (kCs == 3) {
            const SegT* sp = static_cast<const SegT*>(A.seg) + pix0 * 3;
            if (vec) {                                     // three 128-bit loads
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const uint4 u = __ldg(reinterpret_cast<const uint4*>(sp) + k);
                    w[k * 4] = u.x; w[k * 4 + 1] = u.y; w[k * 4 + 2] = u.z; w[k * 4 + 3] = u.w;
                }
            } else {
#pragma unroll
                for (int i = 0; i < 12; ++i) w[i] = i < 3 * n ? __ldg(reinterpret_cast<const uint32_t*>(sp) + i) : 0u;
            }
        }
        // csum * alpha of the classes that passed: from the subset table, or (more than four classes near the
        // block) added up in class order here
        float ca[4][3];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t m = (cm[0][q >> 1] >> (16 * (q & 1))) & 0xffffu;
            if (P <= 4) {
                const float4 t = ctab[m & 15u];
                ca[q][0] = t.x; ca[q][1] = t.y; ca[q][2] = t.z;
            } else {
                float c0 = 0.0f, c1 = 0.0f, c2 = 0.0f;
                for (int p = 0; p < P; ++p)
                    if ((m >> p) & 1u) {
                        const int c = s_pcls[p];
                        c0 = __fadd_rn(c0, A.inst.rgb[c][0]); c1 = __fadd_rn(c1, A.inst.rgb[c][1]); c2 = __fadd_rn(c2, A.inst.rgb[c][2]);
                    }
                ca[q][0] = __fmul_rn(c0, ia); ca[q][1] = __fmul_rn(c1, ia); ca[q][2] = __fmul_rn(c2, ia);
            }
        }
        float img[12];
        raw.pixels(img, vec);
        if (A.line_bits) {
            // DrawBoxes in front of the overlays (serving.py:34): its uint8 canvas (clip + truncation for float frames)
            // with the rectangles' pixels at 255; ox is a multiple of 4, so the four pixels share one word of the bitmap
            if (sizeof(ImgT) != 1) {                       // uint8 frames are their own canvas
#pragma unroll
                for (int i = 0; i < 12; ++i)
                    img[i] = __fsub_rn(trunc_magic(fminf(fmaxf(img[i], 0.0f), 255.0f)), kTwo23);
            }
            lw >>= (ox & 31);
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if ((lw >> q) & 1u) img[q * 3] = img[q * 3 + 1] = img[q * 3 + 2] = 255.0f;
        }
        // (sum_c colours[c] * seg[..., c]) * alpha of DrawSegmentation
        float sc[4][3];
#pragma unroll
        for (int q = 0; q < 4; ++q) sc[q][0] = sc[q][1] = sc[q][2] = 0.0f;
        if (kCs == 3) {
            // an int32 map whose twelve values are all 0 or 1 (a one-hot semantic map): table lookup by the triple
            bool binary = sizeof(SegT) == 4 && !is_float_seg<SegT>();
            if (binary) {
                uint32_t any = 0u;
#pragma unroll
                for (int i = 0; i < 12; ++i) any |= w[i];
                binary = __all_sync(__activemask(), (any & ~1u) == 0u);
            }
            if (binary) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 t = s_stab[w[q * 3] | (w[q * 3 + 1] << 1) | (w[q * 3 + 2] << 2)];
                    sc[q][0] = t.x; sc[q][1] = t.y; sc[q][2] = t.z;
                }
            } else {
                // reduce_sum over the classes, in order; the first term is added to +0, which changes nothing the
                // blend can see (a -0 product survives only into v + (-0) = v)
#pragma unroll
                for (int q = 0; q < 4; ++q)
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        float t = __fmul_rn(A.sem.rgb[0][k], seg_word<SegT>(w[q * 3]));
                        t = __fadd_rn(t, __fmul_rn(A.sem.rgb[1][k], seg_word<SegT>(w[q * 3 + 1])));
                        t = __fadd_rn(t, __fmul_rn(A.sem.rgb[2][k], seg_word<SegT>(w[q * 3 + 2])));
                        sc[q][k] = __fmul_rn(t, sa);
                    }
            }
        } else if (kCs != 0) {
            const int Cs = A.sem.num_classes;
            const SegT* sp = static_cast<const SegT*>(A.seg) + pix0 * Cs;
            for (int q = 0; q < n; ++q) {
                for (int c = 0; c < Cs; ++c) {
                    const float v = seg_f32(sp + q * Cs + c);
#pragma unroll
                    for (int k = 0; k < 3; ++k) sc[q][k] = __fadd_rn(sc[q][k], __fmul_rn(A.sem.rgb[c][k], v));
                }
#pragma unroll
                for (int k = 0; k < 3; ++k) sc[q][k] = __fmul_rn(sc[q][k], sa);
            }
        }
        uint32_t byte[12];
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                float m = trunc_magic(clip255(__fadd_rn(img[q * 3 + k], ca[q][k])));         // 2^23 + uint8
                if (kCs != 0) m = trunc_magic(clip255(__fadd_rn(__fsub_rn(m, kTwo23), sc[q][k])));
                byte[q * 3 + k] = __float_as_uint(m);          // the uint8 is the low byte
            }
        uint8_t* op = A.out + pix0 * 3;
        if (vec) {
#pragma unroll
            for (int k = 0; k < 3; ++k)
                reinterpret_cast<uint32_t*>(op)[k] = __byte_perm(__byte_perm(byte[4 * k], byte[4 * k + 1], 0x0040),
                                                                 __byte_perm(byte[4 * k + 2], byte[4 * k + 3], 0x0040), 0x5410);
        } else {
#pragma unroll
            for (int i = 0; i < 12; ++i)
                if (i < 3 * n) op[i] = (uint8_t)byte[i];
        }
#pragma unroll
        for (int rr = 0; rr + 1 < kRowsPT; ++rr) { cm[rr][0] = cm[rr + 1][0]; cm[rr][1] = cm[rr + 1][1]; }   // next row's masks
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
// then neither copied nor written before the overlay pass.  Horizontal lines are word-wide ORs.  One warp per box.
__device__ __forceinline__ void box_lines_warp(const int32_t* __restrict__ row, int b, int H, int W, int words,
                                               uint32_t* __restrict__ bits, int lane) {
    BoxRect r;
    if (!box_rect(row, H, W, r)) return;
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

// One warp per instance, ahead of draw_tiles_kernel: its clipped-box geometry, the bit rows of its {0,1} tile
// (mask_w <= 32; a tail that already holds bit rows is copied word by word) and - with line_bits - the four
// one-pixel lines of DrawBoxes' rectangle in the one-bit-per-pixel map (every row the tail kept, whatever its class).
__global__ void __launch_bounds__(kDrawThreads)
pack_tiles_kernel(const int32_t* __restrict__ det, const PasteSrc S, int B, int m_rows, int m_stride, int mh, int mw,
                  int PH, int PW, int num_colors, DrawGeom* __restrict__ geom, uint32_t* __restrict__ bits,
                  int32_t* __restrict__ m_used, uint32_t* __restrict__ line_bits, int line_words) {
    int M, thr;
    paste_scalars(S, B, m_rows, M, thr);
    if (m_stride == 0) m_stride = M;
    if (blockIdx.x == 0 && threadIdx.x == 0) *m_used = M;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t inst = (int64_t)blockIdx.x * (kDrawThreads / 32) + warp;
    if (inst >= (int64_t)B * M) return;
    const int b = (int)(inst / M), j = (int)(inst - (int64_t)b * M);
    const int32_t* row = det + ((int64_t)b * m_stride + j) * 6;
    if (line_bits) box_lines_warp(row, b, PH, PW, line_words, line_bits, lane);
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
    if (tref.bits && !tref.mi) {                            // the fused tail's bit rows: a copy
        for (int y = lane; y < mh; y += 32) bits[slot * mh + y] = tref.valid ? __ldg(tref.bits + y) : 0u;
        return;
    }
    for (int y = 0; y < mh; ++y) {
        const int v = lane < mw ? tref.at(y * mw + lane) : 0;
        const unsigned w = __ballot_sync(0xffffffffu, v != 0);
        if (lane == 0) bits[slot * mh + y] = w;
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
    if (boxes) MLP_CUDA(cudaMemsetAsync(line_bits, 0, (size_t)line_bytes, st));
    pack_tiles_kernel<<<(int)((n_inst + wpc - 1) / wpc), kDrawThreads, 0, st>>>(
        det_i32_dev, S, batch, m_rows, m_stride, mask_h, mask_w, frame_h, frame_w, inst_colors->num_classes, geom,
        bits, m_used, line_bits, line_words);
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
