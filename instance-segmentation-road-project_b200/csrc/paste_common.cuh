// paste_common.cuh — pieces of CropAndPadMask shared by the paste kernels (paste.cu) and the
// per-instance reductions evaluated straight from the mask tiles (summary.cu): where an instance's
// 28x28 tile comes from, the clipped box geometry and the two-stage lerp of
// /root/reference/engine/layers/misc.py:373-391 (tf.image.resize, align_corners=True).
#pragma once
#include <limits.h>

#include "common.cuh"

namespace {

constexpr int kMaxTile = 64 * 64;         // mask_h * mask_w <= 4096

// Where the per-instance mask tile and the batch-wide scalars come from.
//   standalone layer : int32 masks [B,stride,mh*mw]; M from m_dev/m_rows; threshold from thr_dev
//   fused tail       : the class channel of the mask head output roi_masks [B,R,mh*mw,C] picked
//                      through the slot->row table of tail_prep_kernel and thresholded at 0.5 on
//                      the fly (TrimInstances + UpSampleOutput never materialise); M and the
//                      confidence threshold are reduced from the per-image counts / conf maxima.
struct PasteSrc {
    const int32_t* masks_i32;
    const int32_t* m_dev;
    const int32_t* thr_dev;
    int fused;
    const float* roi_masks;
    const int32_t* tail_src;      // [B, m_stride] row of roi_masks, -1 = MoldBatch padding
    const uint32_t* tail_bits;    // [B, m_stride, mh] bit rows of the thresholded class channel (mw <= 32), or NULL
    const int32_t* r_dev;
    int r_rows;
    int C;
    const int32_t* counts;        // [B] valid instances per image
    const int32_t* confmax;       // [B] max int confidence of the valid rows (INT_MIN if none)
    int32_t* m_out;               // [1] M written back for the host
    int planar;                   // roi_masks is [B,R,C,mh*mw] (class planes) instead of [B,R,mh*mw,C]
    const int32_t* scalars;       // [2] = (M, row-filter threshold), reduced once by the tail preparation's last CTA
};

struct PasteGeom {
    int xmin, xmax, ymin, ymax;
    float sy, sx;          // resize scales (in-1)/(out-1) or in/out
    bool active;
};

// misc.py:373-386: box = max(box,1) -> float; ceil(c -/+ s/2) -> int -> clip.
__device__ __forceinline__ PasteGeom paste_geometry(const int32_t* row, int thr, int mh, int mw, int PH,
                                                    int PW) {
    PasteGeom g;
    const int conf = row[5];
    const float cx = (float)max(row[0], 1), cy = (float)max(row[1], 1);
    const float w = (float)max(row[2], 1), h = (float)max(row[3], 1);
    const float hw = __fmul_rn(w, 0.5f), hh = __fmul_rn(h, 0.5f);        // w / 2, exactly
    g.xmin = min(max(__float2int_rz(ceilf(__fsub_rn(cx, hw))), 0), PW);
    g.xmax = min(max(__float2int_rz(ceilf(__fadd_rn(cx, hw))), 0), PW);
    g.ymin = min(max(__float2int_rz(ceilf(__fsub_rn(cy, hh))), 0), PH);
    g.ymax = min(max(__float2int_rz(ceilf(__fadd_rn(cy, hh))), 0), PH);
    const int oh = g.ymax - g.ymin, ow = g.xmax - g.xmin;
    g.active = (conf >= thr) && oh > 0 && ow > 0;
    // CalculateResizeScale(in, out, align_corners=true)
    g.sy = (oh > 1) ? __fdiv_rn((float)(mh - 1), (float)(oh - 1)) : __fdiv_rn((float)mh, (float)max(oh, 1));
    g.sx = (ow > 1) ? __fdiv_rn((float)(mw - 1), (float)(ow - 1)) : __fdiv_rn((float)mw, (float)max(ow, 1));
    return g;
}

__device__ __forceinline__ float paste_value(const float* __restrict__ tile, int mh, int mw, int ylo,
                                             int yhi, float ly, int ox_local, float sx) {
    const float p = __fmul_rn((float)ox_local, sx);
    const float fl = floorf(p);
    const int xlo = max((int)fl, 0);
    const int xhi = min((int)ceilf(p), mw - 1);
    const float lx = __fsub_rn(p, fl);
    const float tl = tile[ylo * mw + xlo], tr = tile[ylo * mw + xhi];
    const float bl = tile[yhi * mw + xlo], br = tile[yhi * mw + xhi];
    const float t = __fadd_rn(tl, __fmul_rn(__fsub_rn(tr, tl), lx));
    const float b = __fadd_rn(bl, __fmul_rn(__fsub_rn(br, bl), lx));
    return __fadd_rn(t, __fmul_rn(__fsub_rn(b, t), ly));
}

// Column table of one box: the x half of the resize depends on the column only, so a CTA that is about to
// evaluate many rows of a box computes (xlo, xhi, lx) once per column (same three rounded operations as
// paste_value) and every pixel then costs one 8-byte shared load instead of a multiply, floor, ceil, two
// conversions, a clamp pair and a subtract.  Entry: x = byte offsets xlo*4 | xhi*4 << 16, y = bits of lx.
// Layout: the table covers the frame-aligned segments [sL, sL+bw) of the box; pixel q of segment s lives at
// q * bw + (s - sL), so the lanes of a warp (consecutive segments, same q) read consecutive entries -
// no bank conflicts (the natural [column] order would put them 16 entries = 128 bytes apart: 16-way).
constexpr int kMaxCols = 2048;            // table entries; wider boxes fall back to paste_value

template <int kVec>
__device__ __forceinline__ bool paste_fill_cols(uint2* __restrict__ s_col, const PasteGeom& g, int mw, int tid,
                                                int nthreads) {
    const int sL = g.xmin / kVec;
    const int bw = (g.xmax + kVec - 1) / kVec - sL;
    const int n = bw * kVec;
    if (n > kMaxCols) return false;
    for (int i = tid; i < n; i += nthreads) {
        const int q = i / bw, s = i - q * bw;
        const int c = (sL + s) * kVec + q - g.xmin;          // box column of that pixel (outside the box: unused)
        const float p = __fmul_rn((float)c, g.sx);
        const float fl = floorf(p);
        const int xlo = min(max((int)fl, 0), mw - 1);
        const int xhi = min(max((int)ceilf(p), 0), mw - 1);
        MLP_BOUND(i, kMaxCols);
        s_col[i] = make_uint2((uint32_t)(xlo * 4) | ((uint32_t)(xhi * 4) << 16), __float_as_uint(__fsub_rn(p, fl)));
    }
    return true;
}

// paste_value with the x terms from the column table; row_lo / row_hi point at tile rows ylo / yhi
__device__ __forceinline__ float paste_value_cols(const unsigned char* __restrict__ row_lo,
                                                  const unsigned char* __restrict__ row_hi, float ly, const uint2 e) {
    const uint32_t xlo = e.x & 0xffffu, xhi = e.x >> 16;
    const float lx = __uint_as_float(e.y);
    const float tl = *reinterpret_cast<const float*>(row_lo + xlo), tr = *reinterpret_cast<const float*>(row_lo + xhi);
    const float bl = *reinterpret_cast<const float*>(row_hi + xlo), br = *reinterpret_cast<const float*>(row_hi + xhi);
    const float t = __fadd_rn(tl, __fmul_rn(__fsub_rn(tr, tl), lx));
    const float b = __fadd_rn(bl, __fmul_rn(__fsub_rn(br, bl), lx));
    return __fadd_rn(t, __fmul_rn(__fsub_rn(b, t), ly));
}

// M and the row-filter threshold for this launch (block-uniform, every warp computes it).
__device__ __forceinline__ void paste_scalars(const PasteSrc& S, int B, int m_rows, int& M, int& thr) {
    if (!S.fused) {
        M = S.m_dev ? *S.m_dev : m_rows;
        if (M > m_rows) M = m_rows;
        thr = *S.thr_dev;
        return;
    }
    if (S.scalars) {                      // reduced once per batch (tail_finish): one 8-byte load per thread
        const int2 v = __ldg(reinterpret_cast<const int2*>(S.scalars));
        M = min(v.x, m_rows);
        thr = v.y;
        if ((blockIdx.x | blockIdx.y | blockIdx.z) == 0 && threadIdx.x == 0 && S.m_out) *S.m_out = M;
        return;
    }
    const int lane = threadIdx.x & 31;
    int mx = 0, mn = INT_MAX, cm = INT_MIN;
    for (int i = lane; i < B; i += 32) {
        const int c = S.counts[i];
        mx = max(mx, c); mn = min(mn, c); cm = max(cm, S.confmax[i]);
    }
    for (int o = 16; o > 0; o >>= 1) {
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        cm = max(cm, __shfl_xor_sync(0xffffffffu, cm, o));
    }
    M = max(mx, 1);
    if (M > m_rows) M = m_rows;
    if (mn < M) cm = max(cm, -100);        // MoldBatch padding rows carry conf = int(-1*100)
    thr = (cm > 50) ? 50 : -100;
    if ((blockIdx.x | blockIdx.y | blockIdx.z) == 0 && threadIdx.x == 0 && S.m_out) *S.m_out = M;
}

// Part of the tail preparation, called by ONE CTA per image right after it has published counts[b] / confmax[b]
// (before its tile copies, so the reduction is off the kernel's critical path): the CTA that arrives last reduces
// M = max(1, max counts) and CropAndPadMask's row-filter threshold (misc.py:366-369) ONCE, so that the 25,600
// short CTAs of the paste kernel read two words instead of repeating the reduction.  scalars[2] is the arrival
// counter; the last CTA resets it for the next launch.  Every thread of the CTA calls this.
__device__ __forceinline__ void tail_finish(int32_t* __restrict__ scalars, const int32_t* __restrict__ counts,
                                            const int32_t* __restrict__ confmax, int B, int m_rows, int total_ctas,
                                            int32_t* __restrict__ m_out) {
    __shared__ int s_last;
    __threadfence();                              // this CTA's counts / confmax are visible before it arrives
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(scalars + 2, 1) == total_ctas - 1;
    __syncthreads();
    if (!s_last || threadIdx.x >= 32) return;
    __threadfence();
    const int lane = threadIdx.x;
    int mx = 0, mn = INT_MAX, cm = INT_MIN;
    for (int i = lane; i < B; i += 32) {
        const int c = __ldcg(counts + i);
        mx = max(mx, c); mn = min(mn, c); cm = max(cm, __ldcg(confmax + i));
    }
    for (int o = 16; o > 0; o >>= 1) {
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        cm = max(cm, __shfl_xor_sync(0xffffffffu, cm, o));
    }
    int M = max(mx, 1);
    if (M > m_rows) M = m_rows;
    if (mn < M) cm = max(cm, -100);        // MoldBatch padding rows carry conf = int(-1*100)
    if (lane == 0) {
        scalars[0] = M;
        scalars[1] = (cm > 50) ? 50 : -100;
        scalars[2] = 0;
        if (m_out) *m_out = M;                 // M for the host / the next batch's speculative fill
    }
}

// Tile element i of instance (b, j) as the int the reference's mask tensor would hold.
struct TileRef {
    const int32_t* mi;     // standalone
    const float* mf;       // fused (already offset to the class channel / plane), element stride es
    const uint32_t* bits;  // fused, pre-thresholded bit rows
    int es, mw;
    bool valid;
    __device__ __forceinline__ int at(int i) const {
        if (mi) return __ldg(mi + i);
        if (bits) {
            const int y = i / mw;
            return valid ? (int)((__ldg(bits + y) >> (i - y * mw)) & 1u) : 0;
        }
        return valid ? (int)(__ldg(mf + (int64_t)i * es) > 0.5f) : 0;
    }
    // whole tile -> shared memory as floats, all threads of the CTA (no per-element division on the bit path)
    __device__ __forceinline__ void fill(float* __restrict__ s_tile, int mh, int tid, int nthreads) const {
        if (bits && !mi) {
            const int lane = tid & 31, nw = nthreads >> 5;
            for (int y = tid >> 5; y < mh; y += nw) {
                const uint32_t w = valid ? __ldg(bits + y) : 0u;
                if (lane < mw) s_tile[y * mw + lane] = (float)((w >> lane) & 1u);
            }
            return;
        }
        const int px = mh * mw;
        for (int i = tid; i < px; i += nthreads) s_tile[i] = (float)at(i);
    }
};

__device__ __forceinline__ TileRef tile_ref(const PasteSrc& S, int b, int j, int m_stride, int px, int cls,
                                            int mh, int mw) {
    TileRef t;
    t.mi = nullptr; t.mf = nullptr; t.bits = nullptr; t.es = S.planar ? 1 : S.C; t.mw = mw; t.valid = false;
    if (!S.fused) {
        t.mi = S.masks_i32 + ((int64_t)b * m_stride + j) * px;
        return t;
    }
    const int R = S.r_dev ? *S.r_dev : S.r_rows;
    MLP_BOUND(j, m_stride);
    const int jsrc = S.tail_src[(int64_t)b * m_stride + j];
    MLP_BOUND(jsrc + 1, R + 1);
    t.valid = jsrc >= 0 && cls >= 0 && cls < S.C;
    if (S.tail_bits) t.bits = S.tail_bits + ((int64_t)b * m_stride + j) * mh;
    if (S.planar) t.mf = S.roi_masks + (t.valid ? (((int64_t)b * R + jsrc) * S.C + cls) * px : 0);
    else t.mf = S.roi_masks + (t.valid ? (((int64_t)b * R + jsrc) * px * S.C + cls) : 0);
    return t;
}


// threshold = 50 if max(conf) > 50 else -100  (misc.py:366-369); conf = column 5.
__global__ void __launch_bounds__(1024)
paste_threshold_kernel(const int32_t* __restrict__ det, int B, int m_rows, int m_stride,
                       const int32_t* __restrict__ m_dev, int32_t* __restrict__ thr_out) {
    __shared__ int s_max;
    int M = m_dev ? *m_dev : m_rows;
    if (M > m_rows) M = m_rows;
    if (m_stride == 0) m_stride = M;            // compact [B,M,..] layout, M known on device only
    if (threadIdx.x == 0) s_max = INT_MIN;
    __syncthreads();
    int mx = INT_MIN;
    const int total = B * M;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int b = i / M, j = i - b * M;
        const int v = det[((int64_t)b * m_stride + j) * 6 + 5];
        mx = v > mx ? v : mx;
    }
    for (int o = 16; o > 0; o >>= 1) {
        const int v = __shfl_xor_sync(0xffffffffu, mx, o);
        mx = v > mx ? v : mx;
    }
    if ((threadIdx.x & 31) == 0) atomicMax(&s_max, mx);
    __syncthreads();
    if (threadIdx.x == 0) *thr_out = (s_max > 50) ? 50 : -100;
}

// 16 bytes of one mask row (4 float32 / 16 uint8 pixels) as floats
template <typename T> struct MaskVec;
template <> struct MaskVec<float> {
    static constexpr int kPx = 4;
    uint4 raw;
    __device__ __forceinline__ void load(const float* p) { raw = __ldg(reinterpret_cast<const uint4*>(p)); }
    __device__ __forceinline__ void load_tail(const float* p, int n) {
        uint32_t w[4] = {0u, 0u, 0u, 0u};
        for (int q = 0; q < n; ++q) w[q] = __float_as_uint(__ldg(p + q));
        raw = make_uint4(w[0], w[1], w[2], w[3]);
    }
    __device__ __forceinline__ bool any() const { return ((raw.x | raw.y | raw.z | raw.w) & 0x7fffffffu) != 0u; }
    __device__ __forceinline__ float at(int q) const {
        return __uint_as_float(q == 0 ? raw.x : (q == 1 ? raw.y : (q == 2 ? raw.z : raw.w)));
    }
};
template <> struct MaskVec<uint8_t> {
    static constexpr int kPx = 16;
    uint4 raw;
    __device__ __forceinline__ void load(const uint8_t* p) { raw = __ldg(reinterpret_cast<const uint4*>(p)); }
    __device__ __forceinline__ void load_tail(const uint8_t* p, int n) {
        uint32_t w[4] = {0u, 0u, 0u, 0u};
        for (int q = 0; q < n; ++q) w[q >> 2] |= (uint32_t)__ldg(p + q) << (8 * (q & 3));
        raw = make_uint4(w[0], w[1], w[2], w[3]);
    }
    __device__ __forceinline__ bool any() const { return (raw.x | raw.y | raw.z | raw.w) != 0u; }
    __device__ __forceinline__ float at(int q) const {
        const uint32_t w = (q >> 2) == 0 ? raw.x : ((q >> 2) == 1 ? raw.y : ((q >> 2) == 2 ? raw.z : raw.w));
        return (float)((w >> (8 * (q & 3))) & 0xffu);
    }
};

// Layout of the fused tail's scratch (MLP_ARENA_FUSED), written by tail_prep_kernel in
// mlp_trim_paste: tail_src [B,K] + confmax [B] + bit tiles [B,K,mh] (mask rows of <= 32 columns).
struct FusedTail {
    int32_t* tail_src;
    int32_t* confmax;
    int32_t* scalars;         // [4]: M, row-filter threshold, arrival counter of the preparing CTAs, pad
    uint32_t* tail_bits;      // NULL when bit tiles are not used
    int64_t bytes;
};
inline FusedTail fused_tail_layout(void* base, int batch, int k_rows, int mask_h, int mask_w) {
    // Bit tiles pay off when few instances are pasted (the strided class-channel gather would
    // otherwise be repeated by every band CTA that touches the box); with many instances the
    // gather is better left inside the paste kernel, where its reads overlap the write stream.
    const bool use_bits = mask_w <= 32 && k_rows <= 256;
    FusedTail t;
    const int64_t head = ((int64_t)batch * k_rows + batch + 3) / 4 * 4;          // keeps `scalars` 16-byte aligned
    t.bytes = (head + 4 + (use_bits ? (int64_t)batch * k_rows * mask_h : 0)) * 4;
    t.tail_src = static_cast<int32_t*>(base);
    t.confmax = t.tail_src + (int64_t)batch * k_rows;
    t.scalars = t.tail_src + head;
    t.tail_bits = use_bits ? reinterpret_cast<uint32_t*>(t.scalars + 4) : nullptr;
    return t;
}

}  // namespace
