// paste.cu — CropAndPadMask (a13) with the consumers' > 0.5 threshold (a14) fused in, and the
// fused tail of the path (TrimInstances a11 + UpSampleOutput a12 + CropAndPadMask a13/a14).
//
// Reference: /root/reference/engine/layers/misc.py:358-401 (paste), :457 and :611-615
// (binary mask = pasted > 0.5), :169-188 (UpSampleOutput), engine/layers/instance.py:258-277
// (TrimInstances).  For every instance the reference resizes its 28x28 int mask to the clipped
// box with tf.image.resize(align_corners=True) (legacy bilinear, restated in oracle/tf_ops.py)
// and zero-pads to the frame.  Output [B,M,PH,PW] is the largest tensor of the whole path
// (16.8 GB as uint8 at B=32, M=1000, 512x1024), so the paste kernel is a pure streaming-store
// kernel: every thread-store is 128 bits (16 uint8 / 4 float pixels), only pixels inside the
// clipped box evaluate the two-stage lerp from the mask tile held in shared memory, everything
// else is written as zero without reading memory.
#include <limits.h>
#include <stdlib.h>

#include "paste_common.cuh"

namespace {

constexpr int kPasteThreads = 256;
constexpr int kTileRegs = 4;              // tile elements prefetched per thread (covers 32x32)

// dynamic shared memory of the vector paste kernels: the mask tile as floats, then the column table
__host__ __device__ inline int paste_tile_bytes(int px) { return (px * 4 + 15) & ~15; }
inline size_t paste_smem_bytes(int mask_h, int mask_w, int frame_w) {
    (void)frame_w;
    return (size_t)paste_tile_bytes(mask_h * mask_w) + (size_t)kMaxCols * sizeof(uint4);   // 16-byte entries on the bit path
}

// One 16-byte output segment (kVec pixels starting at x0 = seg * kVec) of frame row oy of an instance: the
// reference's two-stage lerp from the mask tile in shared memory, > 0.5 fused for the binary modes.
// kCols: x terms from the box's column table (paste_fill_cols; col points at entry (q = 0, this segment),
// pitch = segments per box row); otherwise computed per pixel (boxes wider than the table).
template <int kMode, bool kCols>
__device__ __forceinline__ uint4 paste_segment(const float* __restrict__ s_tile, const uint2* __restrict__ col,
                                               int pitch, const PasteGeom& g, int mh, int mw, int oy, int x0) {
    constexpr int kVec = kMode == MLP_PASTE_F32 ? 4 : (kMode == MLP_PASTE_U8 ? 16 : 128);
    const float p = __fmul_rn((float)(oy - g.ymin), g.sy);
    const float fl = floorf(p);
    const int ylo = max((int)fl, 0);
    const int yhi = min((int)ceilf(p), mh - 1);
    const float ly = __fsub_rn(p, fl);
    const unsigned char* row_lo = reinterpret_cast<const unsigned char*>(s_tile + ylo * mw);
    const unsigned char* row_hi = reinterpret_cast<const unsigned char*>(s_tile + yhi * mw);
    const int q0 = max(g.xmin - x0, 0), q1 = min(g.xmax - x0, kVec);       // pixels of the segment inside the box
    auto value = [&](int q) -> float {
        if (kCols) return paste_value_cols(row_lo, row_hi, ly, col[q * pitch]);
        return paste_value(s_tile, mh, mw, ylo, yhi, ly, x0 + q - g.xmin, g.sx);
    };
    uint4 v;
    if (kMode == MLP_PASTE_U8) {
        uint32_t w[4] = {0u, 0u, 0u, 0u};
        // one loop for interior and edge segments alike: a warp holds both kinds, and two code paths would
        // be executed one after the other by every warp (measured: 1.65x the pixel evaluations)
#pragma unroll
        for (int q = 0; q < 16; ++q)
            if (q >= q0 && q < q1 && value(q) > 0.5f) w[q >> 2] |= 1u << ((q & 3) * 8);
        v = make_uint4(w[0], w[1], w[2], w[3]);
    } else if (kMode == MLP_PASTE_BITS) {
        // bit k of byte i = pixel 8*i + k  (numpy packbits, bitorder='little')
        uint32_t w[4] = {0u, 0u, 0u, 0u};
        for (int q = q0; q < q1; ++q)
            if (value(q) > 0.5f) w[q >> 5] |= 1u << (q & 31);
        v = make_uint4(w[0], w[1], w[2], w[3]);
    } else {
        float f[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (q >= q0 && q < q1) f[q] = value(q);
        v = make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]),
                       __float_as_uint(f[3]));
    }
    return v;
}

// ---- the same for tiles that arrive as bit rows (the fused tail: {0,1} by construction, mw <= 32) ----------------
// Column table entry: single-bit masks of the two source columns, lx and fadd(1, -lx); columns of the table's
// segments that lie outside the box carry empty masks, so their value is +0 without a range test.  The x lerp
// tl + (tr - tl) * lx of a tile row is then one of 0, lx, 1 - lx, 1 - picked by the two corner bits with the same
// float32 result as the arithmetic - and a pixel costs one 16-byte shared load instead of five loads.
template <int kVec>
__device__ __forceinline__ void paste_fill_cols_bits(uint4* __restrict__ s_col, const PasteGeom& g, int mw, int tid,
                                                     int nthreads) {
    const int sL = g.xmin / kVec;
    const int bw = (g.xmax + kVec - 1) / kVec - sL;
    const int n = bw * kVec;                                 // <= kMaxCols (the caller checked)
    const int wpx = g.xmax - g.xmin;
    for (int i = tid; i < n; i += nthreads) {
        const int q = i / bw, sg = i - q * bw;
        const int c = (sL + sg) * kVec + q - g.xmin;         // box column of that pixel
        const float p = __fmul_rn((float)c, g.sx);
        const float fl = floorf(p);
        const int xlo = min(max((int)fl, 0), mw - 1);
        const int xhi = min(max((int)ceilf(p), 0), mw - 1);
        const float lx = __fsub_rn(p, fl);
        const bool in = c >= 0 && c < wpx;
        MLP_BOUND(i, kMaxCols);
        s_col[i] = make_uint4(in ? 1u << xlo : 0u, in ? 1u << xhi : 0u, __float_as_uint(lx),
                              __float_as_uint(__fsub_rn(1.0f, lx)));
    }
}

__device__ __forceinline__ float paste_value_bits(uint32_t w0, uint32_t w1, float ly, const uint4 e) {
    const float lx = __uint_as_float(e.z), oml = __uint_as_float(e.w);
    const bool tl = (w0 & e.x) != 0u, tr = (w0 & e.y) != 0u, bl = (w1 & e.x) != 0u, br = (w1 & e.y) != 0u;
    const float t = tl ? (tr ? 1.0f : oml) : (tr ? lx : 0.0f);
    const float b = bl ? (br ? 1.0f : oml) : (br ? lx : 0.0f);
    return __fadd_rn(t, __fmul_rn(__fsub_rn(b, t), ly));
}

template <int kMode>
__device__ __forceinline__ uint4 paste_segment_bits(const uint32_t* __restrict__ s_rows, const uint4* __restrict__ col,
                                                    int pitch, const PasteGeom& g, int mh, int oy, int x0) {
    constexpr int kVec = kMode == MLP_PASTE_F32 ? 4 : (kMode == MLP_PASTE_U8 ? 16 : 128);
    const float p = __fmul_rn((float)(oy - g.ymin), g.sy);
    const float fl = floorf(p);
    const uint32_t w0 = s_rows[max((int)fl, 0)], w1 = s_rows[min((int)ceilf(p), mh - 1)];
    const float ly = __fsub_rn(p, fl);
    uint4 v;
    if (kMode == MLP_PASTE_U8) {
        uint32_t w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
        for (int q = 0; q < 16; ++q)
            if (paste_value_bits(w0, w1, ly, col[q * pitch]) > 0.5f) w[q >> 2] |= 1u << ((q & 3) * 8);
        v = make_uint4(w[0], w[1], w[2], w[3]);
    } else if (kMode == MLP_PASTE_BITS) {
        uint32_t w[4] = {0u, 0u, 0u, 0u};
        const int q0 = max(g.xmin - x0, 0), q1 = min(g.xmax - x0, kVec);   // pixels of the segment inside the box
        for (int q = q0; q < q1; ++q)
            if (paste_value_bits(w0, w1, ly, col[q * pitch]) > 0.5f) w[q >> 5] |= 1u << (q & 31);
        v = make_uint4(w[0], w[1], w[2], w[3]);
    } else {
        float f[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) f[q] = paste_value_bits(w0, w1, ly, col[q * pitch]);
        v = make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3]));
    }
    return v;
}

template <int kMode>
__device__ __forceinline__ void paste_box_rows_bits(const uint32_t* __restrict__ s_rows, const uint4* __restrict__ s_col,
                                                    const PasteGeom& g, int mh, int y_first, int nrows, int sL, int bw,
                                                    uint4* __restrict__ rows_out, int spr, int tid) {
    constexpr int kVec = kMode == MLP_PASTE_F32 ? 4 : (kMode == MLP_PASTE_U8 ? 16 : 128);
    const int nb = nrows * bw;
    // (sharing a segment between 2 or 4 threads for boxes with few segments was measured: no change)
    // (row, segment) of item i = tid + k * 256, advanced without a division per item
    int r = tid / bw, sg = tid - r * bw;
    const int dr = kPasteThreads / bw, ds = kPasteThreads - dr * bw;
    for (int i = tid; i < nb; i += kPasteThreads) {
        const uint4 v = paste_segment_bits<kMode>(s_rows, s_col + sg, bw, g, mh, y_first + r, (sL + sg) * kVec);
        stg_stream_u4(rows_out + (int64_t)r * spr + sL + sg, v);
        r += dr; sg += ds;
        if (sg >= bw) { sg -= bw; ++r; }
    }
}

// All segments [0, nb) of `nrows` box rows starting at frame row y_first, flattened over the CTA.
template <int kMode>
__device__ __forceinline__ void paste_box_rows(const float* __restrict__ s_tile, const uint2* __restrict__ s_col,
                                               bool cols, const PasteGeom& g, int mh, int mw, int y_first, int nrows,
                                               int sL, int bw, uint4* __restrict__ rows_out, int spr, int tid) {
    constexpr int kVec = kMode == MLP_PASTE_F32 ? 4 : (kMode == MLP_PASTE_U8 ? 16 : 128);
    const int nb = nrows * bw;
    if (cols) {
        for (int i = tid; i < nb; i += kPasteThreads) {
            const int r = i / bw, s = i - r * bw;
            const uint4 v = paste_segment<kMode, true>(s_tile, s_col + s, bw, g, mh, mw, y_first + r, (sL + s) * kVec);
            stg_stream_u4(rows_out + (int64_t)r * spr + sL + s, v);
        }
    } else {
        for (int i = tid; i < nb; i += kPasteThreads) {
            const int r = i / bw, s = i - r * bw;
            const uint4 v = paste_segment<kMode, false>(s_tile, nullptr, 0, g, mh, mw, y_first + r, (sL + s) * kVec);
            stg_stream_u4(rows_out + (int64_t)r * spr + sL + s, v);
        }
    }
}

// One CTA per (instance, band) item, NOT a persistent grid: on B200 a write-only stream of many
// short-lived CTAs, each owning one contiguous 64 KB band, reaches ~7.4 TB/s while persistent
// CTAs top out near 6.3 TB/s (tools/write_bw.cu, profiles/write_bw_r01.txt).  The grid is sized
// for the capacity m_rows; CTAs past the device-side M exit at once.  kU8: 16 pixels per
// thread-store, else 4 float pixels; PW must be a multiple of that (host checks; the scalar
// kernel handles the rest).  Three phases so that no warp mixes cheap and expensive lanes:
//   A  rows of the band above/below the clipped box: contiguous, flat unrolled zero stores;
//   B1 rows crossing the box: the 16-byte segments left and right of it (zeros);
//   B2 the segments intersecting the box, flattened over ALL threads of the CTA: each evaluates
//      the reference's two-stage lerp per pixel from the mask tile in shared memory.
#ifndef MLP_PASTE_MIN_CTAS
#define MLP_PASTE_MIN_CTAS 4
#endif
template <int kMode>       // MLP_PASTE_F32 / MLP_PASTE_U8 / MLP_PASTE_BITS
__global__ void __launch_bounds__(kPasteThreads, MLP_PASTE_MIN_CTAS)
paste_kernel(const int32_t* __restrict__ det, const PasteSrc S, int B, int m_rows, int m_stride, int mh,
             int mw, int PH, int PW, int band_rows, int bands, int grid3, int use_cols, void* __restrict__ out) {
    constexpr int kVec = kMode == MLP_PASTE_F32 ? 4 : (kMode == MLP_PASTE_U8 ? 16 : 128);   // pixels per 16-byte store
    extern __shared__ __align__(16) unsigned char s_dyn[];           // [tile floats][column table], paste_smem_bytes
    float* s_tile = reinterpret_cast<float*>(s_dyn);
    uint2* s_col_buf = reinterpret_cast<uint2*>(s_dyn + paste_tile_bytes(mh * mw));
    // Items are numbered over the CAPACITY grid (image, slot < m_rows, band), so that a CTA knows its instance - and
    // can fetch its detection row - without waiting for the device-side M; slots >= M leave once M has arrived.
    // grid3: the launch is (band, slot, image) - one item per CTA, its coordinates straight from blockIdx (no
    // integer division on the way to the first store); else a 1-D grid-stride loop over the same numbering.
    const int64_t items = grid3 ? 1 : (int64_t)B * m_rows * bands;
    const int spr = PW / kVec;                     // 16-byte segments per frame row
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int px = mh * mw;
    const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
    int M = -1, thr = 0;

    for (int64_t item = grid3 ? 0 : blockIdx.x; item < items; item += gridDim.x) {
        int band, b, j;
        if (grid3) {
            band = blockIdx.x; j = blockIdx.y; b = blockIdx.z;
        } else {
            int slot;
            if (items <= 0x7fffffffLL) {            // the usual case: 32-bit division
                slot = (int)((uint32_t)item / (uint32_t)bands);
                band = (int)((uint32_t)item - (uint32_t)slot * (uint32_t)bands);
            } else {
                slot = (int)(item / bands);
                band = (int)(item - (int64_t)slot * bands);
            }
            b = slot / m_rows; j = slot - b * m_rows;
        }
        int row[6];
        if (m_stride != 0) {                        // capacity layout: the row does not depend on M - load it first
            const int2* p = reinterpret_cast<const int2*>(det + ((int64_t)b * m_stride + j) * 6);
            const int2 a0 = __ldg(p), a1 = __ldg(p + 1), a2 = __ldg(p + 2);
            row[0] = a0.x; row[1] = a0.y; row[2] = a1.x; row[3] = a1.y; row[4] = a2.x; row[5] = a2.y;
        }
        if (M < 0) paste_scalars(S, B, m_rows, M, thr);
        if (j >= M) continue;                       // block-uniform
        const int stride = m_stride ? m_stride : M; // compact [B,M,..] input layout
        if (m_stride == 0) {
            const int2* p = reinterpret_cast<const int2*>(det + ((int64_t)b * stride + j) * 6);
            const int2 a0 = __ldg(p), a1 = __ldg(p + 1), a2 = __ldg(p + 2);
            row[0] = a0.x; row[1] = a0.y; row[2] = a1.x; row[3] = a1.y; row[4] = a2.x; row[5] = a2.y;
        }
        const int inst = b * M + j;                 // position in the [B,M,PH,PW] output
        const PasteGeom g = paste_geometry(row, thr, mh, mw, PH, PW);
        const int y0 = band * band_rows;
        const int y1 = min(y0 + band_rows, PH);
        // rows of this band that cross the clipped box: [ya, yb)
        int ya = y1, yb = y1;
        if (g.active) { ya = min(max(g.ymin, y0), y1); yb = min(max(g.ymax, ya), y1); }
        const bool touches = yb > ya;              // block-uniform
        // the mask tile is only needed by phase B2: issue its loads now, park them in shared
        // memory after the zero rows have been streamed out
        int tile_regs[kTileRegs];
        uint32_t row_word = 0u;                     // bit path: tile row `tid`
        bool bitp = false;                          // block-uniform
        TileRef tref;
        tref.mi = nullptr; tref.mf = nullptr; tref.bits = nullptr; tref.es = 1; tref.mw = mw; tref.valid = false;
        if (touches) {
            tref = tile_ref(S, b, j, stride, px, row[4], mh, mw);
            // bit rows + a column table that fits: the bit path (one 4-byte load per tile row instead of the tile)
            bitp = use_cols && tref.bits && !tref.mi && mh <= kPasteThreads &&
                   ((g.xmax + kVec - 1) / kVec - g.xmin / kVec) * kVec <= kMaxCols;
            if (bitp) {
                if (tid < mh && tref.valid) row_word = __ldg(tref.bits + tid);
            } else {
#pragma unroll
                for (int q = 0; q < kTileRegs; ++q) {
                    const int i = tid + q * kPasteThreads;
                    tile_regs[q] = (i < px) ? tref.at(i) : 0;
                }
            }
        }
        uint4* band_base = reinterpret_cast<uint4*>(out) + ((int64_t)inst * PH + y0) * spr;
        // ---- phase A: zero rows [y0,ya) and [yb,y1)
        {
            const int n_top = (ya - y0) * spr;
            int i = tid;
            for (; i + 3 * kPasteThreads < n_top; i += 4 * kPasteThreads) {
                stg_stream_u4(band_base + i, zero4);
                stg_stream_u4(band_base + i + kPasteThreads, zero4);
                stg_stream_u4(band_base + i + 2 * kPasteThreads, zero4);
                stg_stream_u4(band_base + i + 3 * kPasteThreads, zero4);
            }
            for (; i < n_top; i += kPasteThreads) stg_stream_u4(band_base + i, zero4);
            uint4* bot = band_base + (int64_t)(yb - y0) * spr;
            const int n_bot = (y1 - yb) * spr;
            i = tid;
            for (; i + 3 * kPasteThreads < n_bot; i += 4 * kPasteThreads) {
                stg_stream_u4(bot + i, zero4);
                stg_stream_u4(bot + i + kPasteThreads, zero4);
                stg_stream_u4(bot + i + 2 * kPasteThreads, zero4);
                stg_stream_u4(bot + i + 3 * kPasteThreads, zero4);
            }
            for (; i < n_bot; i += kPasteThreads) stg_stream_u4(bot + i, zero4);
        }
        if (!touches) continue;
        __syncthreads();                           // previous item's phase B2 done with s_tile
        bool cols = false;
        if (bitp) {
            if (tid < mh) reinterpret_cast<uint32_t*>(s_tile)[tid] = row_word;
            paste_fill_cols_bits<kVec>(reinterpret_cast<uint4*>(s_col_buf), g, mw, tid, kPasteThreads);
        } else {
#pragma unroll
            for (int q = 0; q < kTileRegs; ++q) {
                const int i = tid + q * kPasteThreads;
                if (i < px) s_tile[i] = (float)tile_regs[q];
            }
            for (int i = tid + kTileRegs * kPasteThreads; i < px; i += kPasteThreads)
                s_tile[i] = (float)tref.at(i);
            cols = use_cols && paste_fill_cols<kVec>(s_col_buf, g, mw, tid, kPasteThreads);
        }
        const int sL = g.xmin / kVec;                       // first segment touching the box
        const int sR = (g.xmax + kVec - 1) / kVec;          // one past the last
        const int bw = sR - sL;
        uint4* box_rows = band_base + (int64_t)(ya - y0) * spr;
        // ---- phase B1: zero strips left/right of the box, one warp per row
        {
            const int nz = spr - bw;
            for (int r = warp; r < yb - ya; r += kPasteThreads / 32) {
                uint4* rp = box_rows + (int64_t)r * spr;
                for (int k = lane; k < nz; k += 32) stg_stream_u4(rp + (k < sL ? k : k + bw), zero4);
            }
        }
        __syncthreads();                           // s_tile / column table ready
        // ---- phase B2: segments intersecting the box, flattened over the CTA
        if (bitp)
            paste_box_rows_bits<kMode>(reinterpret_cast<const uint32_t*>(s_tile), reinterpret_cast<const uint4*>(s_col_buf),
                                       g, mh, ya, yb - ya, sL, bw, box_rows, spr, tid);
        else
            paste_box_rows<kMode>(s_tile, s_col_buf, cols, g, mh, mw, ya, yb - ya, sL, bw, box_rows, spr, tid);
    }
}

// Generic fall-back for frame widths that are not a multiple of the vector width.
template <int kMode>
__global__ void __launch_bounds__(kPasteThreads)
paste_scalar_kernel(const int32_t* __restrict__ det, const PasteSrc S, int B, int m_rows, int m_stride,
                    int mh, int mw, int PH, int PW, void* __restrict__ out) {
    __shared__ float s_tile[kMaxTile];
    int M, thr;
    paste_scalars(S, B, m_rows, M, thr);
    if (m_stride == 0) m_stride = M;
    const int px = mh * mw;
    const int64_t items = (int64_t)B * M;
    for (int64_t inst = blockIdx.x; inst < items; inst += gridDim.x) {
        const int b = (int)(inst / M), j = (int)(inst - (int64_t)b * M);
        const int32_t* row = det + ((int64_t)b * m_stride + j) * 6;
        const PasteGeom g = paste_geometry(row, thr, mh, mw, PH, PW);
        __syncthreads();
        if (g.active) {
            const TileRef tref = tile_ref(S, b, j, m_stride, px, row[4], mh, mw);
            for (int i = threadIdx.x; i < px; i += kPasteThreads) s_tile[i] = (float)tref.at(i);
        }
        __syncthreads();
        const int64_t npx = (int64_t)PH * PW;
        auto value_at = [&](int oy, int ox) -> float {
            if (!(g.active && oy >= g.ymin && oy < g.ymax && ox >= g.xmin && ox < g.xmax)) return 0.0f;
            const float p = __fmul_rn((float)(oy - g.ymin), g.sy);
            const float fl = floorf(p);
            const int ylo = max((int)fl, 0);
            const int yhi = min((int)ceilf(p), mh - 1);
            return paste_value(s_tile, mh, mw, ylo, yhi, __fsub_rn(p, fl), ox - g.xmin, g.sx);
        };
        if (kMode == MLP_PASTE_BITS) {                       // PW % 8 == 0 (host checks)
            const int bpr = PW >> 3;
            const int64_t nbytes = (int64_t)PH * bpr;
            for (int64_t i = threadIdx.x; i < nbytes; i += kPasteThreads) {
                const int oy = (int)(i / bpr), bx = (int)(i - (int64_t)oy * bpr);
                unsigned byte = 0;
#pragma unroll
                for (int k = 0; k < 8; ++k) byte |= (unsigned)(value_at(oy, bx * 8 + k) > 0.5f) << k;
                static_cast<unsigned char*>(out)[inst * nbytes + i] = (unsigned char)byte;
            }
        } else {
            for (int64_t i = threadIdx.x; i < npx; i += kPasteThreads) {
                const int oy = (int)(i / PW), ox = (int)(i - (int64_t)oy * PW);
                const float val = value_at(oy, ox);
                if (kMode == MLP_PASTE_U8) static_cast<unsigned char*>(out)[inst * npx + i] = val > 0.5f;
                else static_cast<float*>(out)[inst * npx + i] = val;
            }
        }
    }
}

// ---- work items of the boxes-only paste ---------------------------------------------------------
// Behind a background fill (mlp_paste_prefill) only the 16-byte segments that intersect a box are left
// to write: 18 M of 1,678 M pixels at cfg-2, very unevenly spread (a 45x45 box beside a frame-sized one).
// The tail preparation therefore cuts every box into chunks of about kBoxSegs segments - (instance,
// first row, rows) - and appends them to a list (one warp-aggregated atomicAdd per 32 instances); the
// paste then walks the list with a grid-stride loop, every CTA doing equal work.
constexpr int kBoxSegs = 2048;             // segments per work item (8 per thread)

struct BoxItems {
    uint2* items;            // x = b*K + slot, y = (first frame row << 16) | rows
    int32_t* n_items;        // [1], zeroed before the tail preparation
    int cap;
    int PH, PW, vec;         // frame size, pixels per 16-byte segment; vec == 0: no list wanted
};

// rows of one chunk for a box of `bw` segments per row
__device__ __forceinline__ int box_chunk_rows(int bw) { return max(1, kBoxSegs / max(bw, 1)); }

// Called by all 32 lanes; lanes with `has` own a valid instance (int32 row o[6]).  Geometry as in
// paste_geometry but without the confidence filter (unknown until every image is prepared): filtered
// rows cost the paste one skipped item.
__device__ __forceinline__ void box_items_append(const BoxItems& Q, bool has, const int32_t* o, int inst) {
    int chunks = 0, ymin = 0, ymax = 0, rows = 1;
    if (has) {
        const PasteGeom g = paste_geometry(o, INT_MIN, 1, 1, Q.PH, Q.PW);
        if (g.active) {
            const int bw = (g.xmax + Q.vec - 1) / Q.vec - g.xmin / Q.vec;
            rows = box_chunk_rows(bw);
            ymin = g.ymin; ymax = g.ymax;
            chunks = (ymax - ymin + rows - 1) / rows;
        }
    }
    const int lane = threadIdx.x & 31;
    int incl = chunks;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    if (total == 0) return;
    int base = 0;
    if (lane == 31) base = atomicAdd(Q.n_items, total);
    base = __shfl_sync(0xffffffffu, base, 31) + incl - chunks;
    for (int k = 0; k < chunks; ++k) {
        const int y0 = ymin + k * rows;
        if (base + k < Q.cap)
            Q.items[base + k] = make_uint2((uint32_t)inst, ((uint32_t)y0 << 16) | (uint32_t)min(rows, ymax - y0));
    }
}

// ---- fused tail, step 1 -------------------------------------------------------------------
// One CTA per image: ranks the rows of roi_boxes [B,R,6] whose class != -1 in order j
// (TrimInstances), converts them to the int32 rows of UpSampleOutput, and records slot -> row so
// that the paste kernel can pick every instance's own class channel straight from the mask head
// output.  Capacity layout (K rows per image), so nothing here depends on M.
constexpr int kPrepThreads = 256;
constexpr int kPrepParts = 8;             // minimum CTAs per image (each repeats the cheap scan, owns 1/parts of the slots)

// ---- background fill --------------------------------------------------------------------------
// Zeroes the valid prefix [B,M,PH,row] of the paste output: 97 % of CropAndPadMask's bytes do not
// depend on the masks at all, only on M.  mlp_paste_prefill launches this right after the NMS
// kernels on a parallel branch, so the stream of zeros overlaps RoIAlign, the mask head and the
// tail preparation instead of waiting for them; the paste kernel then only writes the boxes
// (paste_boxes_kernel).  One short-lived CTA per 64 KB, the launch shape that streams stores fastest on B200
// (profiles/write_bw_r01.txt); CTAs past the device-side M exit at once.
constexpr int kFillThreads = 256;
constexpr int kFillVecs = 4096;           // uint4 per CTA = 64 KB

// Fills the instance rows [m_from, m_to) of every image's slab, i.e. the flat range
// [B * m_from * inst_vecs, B * m_to * inst_vecs): m_from = *m_from_dev (0 when NULL); m_to = max(1, max counts)
// when counts is given, else *m_to_dev; both clamped to the capacity m_rows.
__global__ void __launch_bounds__(kFillThreads)
paste_fill_kernel(const int32_t* __restrict__ m_from_dev, const int32_t* __restrict__ counts,
                  const int32_t* __restrict__ m_to_dev, int B, int m_rows, int64_t inst_vecs,
                  uint4* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    int m_to;
    if (counts) {
        int mx = 0;
        for (int i = lane; i < B; i += 32) mx = max(mx, counts[i]);
        for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        m_to = max(mx, 1);
    } else {
        m_to = *m_to_dev;
    }
    m_to = min(max(m_to, 0), m_rows);
    const int m_from = m_from_dev ? min(max(*m_from_dev, 0), m_rows) : 0;
    if (m_from >= m_to) return;
    const int64_t first = (int64_t)B * m_from * inst_vecs;
    const int64_t total = (int64_t)B * m_to * inst_vecs;
    const int64_t start = first + (int64_t)blockIdx.x * kFillVecs;
    if (start >= total) return;
    const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
    uint4* p = out + start;
    if (total - start >= kFillVecs) {
#pragma unroll
        for (int q = 0; q < kFillVecs / kFillThreads; ++q) stg_stream_u4(p + q * kFillThreads + threadIdx.x, zero4);
    } else {
        const int n = (int)(total - start);
        for (int i = threadIdx.x; i < n; i += kFillThreads) stg_stream_u4(p + i, zero4);
    }
}

// ---- fused tail, step 1, TMA form ---------------------------------------------------------------
// Same contract as tail_prep_kernel; the bit tiles are built differently.  The mask head's output of
// one RoI is ONE contiguous block ([mh*mw, C] interleaved: 15.7 KB at 28x28x5, or the 3 KB class
// plane with the planar layout), so a warp fetches it with one cp.async.bulk into its own staging
// slot (mbarrier completion, no registers held by loads in flight, full lines from DRAM) and then
// thresholds its class channel out of shared memory with one ballot per mask row.
constexpr int kPrepTmaWarps = 4;

// kByRow: a valid row belongs to the CTA `j % parts` (its source row) instead of `slot % parts`, and s_slot[j]
// receives the slot of every row (kNoSlot: not an instance, or beyond the capacity K) - the bulk-copy kernel starts
// its copies by source row before the ranks are known.
constexpr int kSlotMapCap = 4096;          // source rows per image the slot map holds
constexpr unsigned short kNoSlot = 0xffffu;

template <int kWarps, bool kByRow>
__device__ __forceinline__ int tail_scan(const float* __restrict__ rows, int R, int K, float rh, float rw,
                                         int part, int parts, int32_t* __restrict__ drows,
                                         int32_t* __restrict__ src, int* s_cnt, int* s_base, int* s_cm,
                                         const BoxItems& Q, int b, unsigned short* s_slot = nullptr) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int seg = (R + kWarps - 1) / kWarps;
    const int j0 = min(warp * seg, R), j1 = min(j0 + seg, R);
    int cnt = 0;
    for (int jb = j0; jb < j1; jb += 32) {
        const int j = jb + lane;
        const bool hit = (j < j1) && (rows[(int64_t)j * 6 + 4] != -1.0f);
        cnt += __popc(__ballot_sync(0xffffffffu, hit));
    }
    if (lane == 0) s_cnt[warp] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int w = 0; w < kWarps; ++w) { s_base[w] = acc; acc += s_cnt[w]; }
        s_base[kWarps] = acc;
    }
    __syncthreads();
    const int total = min(s_base[kWarps], K);
    int base = s_base[warp];
    int cm = INT_MIN;
    for (int jb = j0; jb < j1; jb += 32) {
        const int j = jb + lane;
        const bool hit = (j < j1) && (rows[(int64_t)j * 6 + 4] != -1.0f);
        const unsigned mask = __ballot_sync(0xffffffffu, hit);
        int32_t o[6] = {0, 0, 0, 0, 0, 0};
        int slot = -1;
        bool mine = false;
        if (hit) {
            slot = base + __popc(mask & ((1u << lane) - 1u));
            if (slot < K) {
                float r[6];
#pragma unroll
                for (int q = 0; q < 6; ++q) r[q] = rows[(int64_t)j * 6 + q];
                upsample_row(r, rh, rw, o);
                cm = max(cm, o[5]);
                mine = kByRow ? (j % parts == part) : (slot % parts == part);
                if (mine) {
#pragma unroll
                    for (int q = 0; q < 6; ++q) drows[slot * 6 + q] = o[q];
                    src[slot] = j;
                }
            }
        }
        if (kByRow && j < j1) s_slot[j] = (hit && slot < K) ? (unsigned short)slot : kNoSlot;
        if (Q.vec && mask) box_items_append(Q, mine, o, b * K + slot);       // warp-uniform branch
        base += __popc(mask);
    }
    for (int o = 16; o > 0; o >>= 1) cm = max(cm, __shfl_xor_sync(0xffffffffu, cm, o));
    if (lane == 0) s_cm[warp] = cm;
    // MoldBatch padding rows (-1) after UpSampleOutput
    {
        const float m1[6] = {-1.f, -1.f, -1.f, -1.f, -1.f, -1.f};
        int32_t o[6];
        upsample_row(m1, rh, rw, o);
        for (int s = total + part + parts * (int)threadIdx.x; s < K; s += parts * kWarps * 32) {
#pragma unroll
            for (int q = 0; q < 6; ++q) drows[s * 6 + q] = o[q];
            src[s] = -1;
        }
    }
    __syncthreads();                      // src[] / drows[] of this CTA's slots are written
    return total;
}

__global__ void __launch_bounds__(kPrepTmaWarps * 32)
tail_prep_tma_kernel(const float* __restrict__ roi_boxes, const float* __restrict__ roi_masks, int r_rows,
                     const int32_t* __restrict__ r_dev, int K, int mh, int mw, int C, int planar, float rh,
                     float rw, int32_t* __restrict__ det_i32, int32_t* __restrict__ tail_src,
                     uint32_t* __restrict__ tail_bits, int32_t* __restrict__ counts,
                     int32_t* __restrict__ confmax, uint32_t slot_bytes, const BoxItems Q,
                     int32_t* __restrict__ scalars, int32_t* __restrict__ m_out) {
    constexpr int kWarps = kPrepTmaWarps;
    extern __shared__ __align__(128) unsigned char s_stage[];     // [kWarps][slot_bytes]
    __shared__ __align__(8) uint64_t s_bar[kWarps];
    __shared__ int s_cnt[kWarps], s_base[kWarps + 1], s_cm[kWarps];
    __shared__ unsigned short s_slot[kSlotMapCap];                // source row -> slot (R <= kSlotMapCap, host checks)
    const int b = blockIdx.x, part = blockIdx.y, parts = gridDim.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) mbar_init(&s_bar[warp], 1);
    int R = r_dev ? *r_dev : r_rows;
    if (R > r_rows) R = r_rows;
    const float* rows = roi_boxes + (int64_t)b * R * 6;
    int32_t* drows = det_i32 + (int64_t)b * K * 6;
    int32_t* src = tail_src + (int64_t)b * K;
    const int px = mh * mw;
    unsigned char* stage = s_stage + (size_t)warp * slot_bytes;
    uint64_t* bar = &s_bar[warp];
    uint32_t phase = 0;
    const int es = planar ? 1 : C;
    // The copy of a RoI's block depends on its SOURCE row only, its destination on the rank of that row: the warp
    // starts the copy of its first row now and learns the ranks (tail_scan) while the bytes are in flight.
    // Returns the class (>= 0) if a copy was started, -1 for a class outside the head's channels (all-zero tile, no
    // copy), -2 for a row that is no instance.  Called by the whole warp.
    auto start_copy = [&](int j) -> int {
        const float cf = __ldg(rows + (int64_t)j * 6 + 4);
        if (cf == -1.0f) return -2;
        const int cls = __float2int_rz(cf);
        if (cls < 0 || cls >= C) return -1;
        const float* g = planar ? roi_masks + (((int64_t)b * R + j) * C + cls) * px
                                : roi_masks + ((int64_t)b * R + j) * px * C;
        if (lane == 0) {
            mbar_expect_tx(bar, slot_bytes);
            // the warp's generic-proxy reads of the previous tile (ordered by the __syncwarp before the call)
            // come before the async proxy overwrites the slot
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            tma_load_1d(stage, g, slot_bytes, bar);
        }
        __syncwarp();
        return cls;
    };
    const int jstep = parts * kWarps;
    int j = part + parts * warp;                      // the rows of this warp: j, j + jstep, ...
    __syncwarp();                                     // mbarrier initialised before its first use
    int st = j < R ? start_copy(j) : -2;
    const int total = tail_scan<kWarps, true>(rows, R, K, rh, rw, part, parts, drows, src, s_cnt, s_base, s_cm, Q, b,
                                              s_slot);
    if (part == 0) {                                  // one CTA per image publishes its count / confidence maximum and
        if (threadIdx.x == 0) {                       // arrives; the last image reduces M and the row-filter threshold
            int m = INT_MIN;                          // while every CTA's tile copies are still in flight
            for (int w = 0; w < kWarps; ++w) m = max(m, s_cm[w]);
            counts[b] = total;
            confmax[b] = m;
        }
        tail_finish(scalars, counts, confmax, gridDim.x, K, gridDim.x, m_out);
    }
    while (j < R) {
        const unsigned slot = s_slot[j];
        if (st >= 0) {                                // the copy of row j has landed (a row beyond K discards it)
            mbar_wait(bar, phase);
            phase ^= 1u;
        }
        if (slot != kNoSlot) {
            MLP_BOUND((int)slot, K);
            uint32_t* out = tail_bits + ((int64_t)b * K + slot) * mh;
            if (st < 0) {                             // tf.gather_nd would raise: all-zero tile
                for (int y = lane; y < mh; y += 32) out[y] = 0u;
            } else {
                const float* t = reinterpret_cast<const float*>(stage) + (planar ? 0 : st);
                for (int y0 = 0; y0 < mh; y0 += 32) {
                    uint32_t mine = 0u;
                    const int yn = min(32, mh - y0);
                    for (int u = 0; u < yn; ++u) {
                        const float v = (lane < mw) ? t[((y0 + u) * mw + lane) * es] : 0.0f;
                        const uint32_t w = __ballot_sync(0xffffffffu, v > 0.5f);
                        if (lane == u) mine = w;
                    }
                    if (lane < yn) out[y0 + lane] = mine;
                }
            }
        }
        __syncwarp();
        j += jstep;
        st = j < R ? start_copy(j) : -2;
    }
}

// Register-gather form of the tail preparation (the fallback of tail_prep_tma_kernel: RoI blocks that are
// not a multiple of 16 bytes or too large to stage, and the tile-less case K > 256).
__global__ void __launch_bounds__(kPrepThreads)
tail_prep_kernel(const float* __restrict__ roi_boxes, const float* __restrict__ roi_masks, int r_rows,
                 const int32_t* __restrict__ r_dev, int K, int mh, int mw, int C, int planar, float rh, float rw,
                 int32_t* __restrict__ det_i32, int32_t* __restrict__ tail_src,
                 uint32_t* __restrict__ tail_bits, int32_t* __restrict__ counts,
                 int32_t* __restrict__ confmax, const BoxItems Q, int32_t* __restrict__ scalars,
                 int32_t* __restrict__ m_out) {
    constexpr int kWarps = kPrepThreads / 32;
    __shared__ int s_cnt[kWarps], s_base[kWarps + 1], s_cm[kWarps];
    const int b = blockIdx.x, part = blockIdx.y, parts = gridDim.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int R = r_dev ? *r_dev : r_rows;
    if (R > r_rows) R = r_rows;
    int32_t* drows = det_i32 + (int64_t)b * K * 6;
    int32_t* src = tail_src + (int64_t)b * K;
    const int total = tail_scan<kWarps, false>(roi_boxes + (int64_t)b * R * 6, R, K, rh, rw, part, parts, drows, src,
                                        s_cnt, s_base, s_cm, Q, b);
    if (part == 0) {
        if (threadIdx.x == 0) {
            int m = INT_MIN;
            for (int w = 0; w < kWarps; ++w) m = max(m, s_cm[w]);
            counts[b] = total;
            confmax[b] = m;
        }
        tail_finish(scalars, counts, confmax, gridDim.x, K, gridDim.x, m_out);
    }
    // bit tiles of this CTA's slots: one warp per slot, lane = mask column, ballot per mask row
    const int px = mh * mw;
    for (int s = part + parts * warp; tail_bits != nullptr && s < total; s += parts * kWarps) {
        const int j = src[s];
        const int cls = drows[s * 6 + 4];
        uint32_t* out = tail_bits + ((int64_t)b * K + s) * mh;
        const bool ok = cls >= 0 && cls < C;
        const int es = planar ? 1 : C;
        const float* m = planar ? roi_masks + (((int64_t)b * R + j) * C + (ok ? cls : 0)) * px
                                : roi_masks + (((int64_t)b * R + j) * px) * C + (ok ? cls : 0);
        // all mask rows of the slot in flight at once (one DRAM round trip per 32 rows)
        for (int y0 = 0; y0 < mh; y0 += 32) {
            float v[32];
#pragma unroll
            for (int u = 0; u < 32; ++u) {
                const int y = y0 + u;
                v[u] = (ok && y < mh && lane < mw) ? __ldg(m + (int64_t)(y * mw + lane) * es) : 0.0f;
            }
#pragma unroll
            for (int u = 0; u < 32; ++u) {
                const unsigned w = __ballot_sync(0xffffffffu, v[u] > 0.5f);
                if (lane == 0 && y0 + u < mh) out[y0 + u] = w;
            }
        }
    }
}

// ---- boxes-only paste: walks the work items of the tail preparation behind a background fill --------
template <int kMode>
__global__ void __launch_bounds__(kPasteThreads)
paste_boxes_kernel(const int32_t* __restrict__ det, const PasteSrc S, int B, int K, int mh, int mw, int PH, int PW,
                   const uint2* __restrict__ items, const int32_t* __restrict__ n_items, int item_cap,
                   void* __restrict__ out) {
    constexpr int kVec = kMode == MLP_PASTE_F32 ? 4 : (kMode == MLP_PASTE_U8 ? 16 : 128);
    extern __shared__ __align__(16) unsigned char s_dyn[];           // [tile floats][column table], paste_smem_bytes
    float* s_tile = reinterpret_cast<float*>(s_dyn);
    uint2* s_col_buf = reinterpret_cast<uint2*>(s_dyn + paste_tile_bytes(mh * mw));
    int M, thr;
    paste_scalars(S, B, K, M, thr);
    const int n = min(*n_items, item_cap);
    const int spr = PW / kVec;
    const int tid = threadIdx.x;
    const int px = mh * mw;
    for (int it = blockIdx.x; it < n; it += gridDim.x) {
        const uint2 q = __ldg(items + it);
        const int b = (int)q.x / K, j = (int)q.x - b * K;
        const int y0 = (int)(q.y >> 16), nrows = (int)(q.y & 0xffffu);
        MLP_BOUND(b, B);
        MLP_BOUND(y0 + nrows - 1, PH);
        int row[6];
        {
            const int2* p = reinterpret_cast<const int2*>(det + ((int64_t)b * K + j) * 6);
            const int2 a0 = __ldg(p), a1 = __ldg(p + 1), a2 = __ldg(p + 2);
            row[0] = a0.x; row[1] = a0.y; row[2] = a1.x; row[3] = a1.y; row[4] = a2.x; row[5] = a2.y;
        }
        const PasteGeom g = paste_geometry(row, thr, mh, mw, PH, PW);
        if (!g.active || j >= M) continue;                  // filtered by the batch-wide confidence rule (block-uniform)
        const TileRef tref = tile_ref(S, b, j, K, px, row[4], mh, mw);
        const int sL = g.xmin / kVec;
        const int bw = (g.xmax + kVec - 1) / kVec - sL;
        const bool bitp = tref.bits && !tref.mi && mh <= kPasteThreads && bw * kVec <= kMaxCols;   // block-uniform
        __syncthreads();                                    // previous item done with s_tile / s_col
        bool cols = false;
        if (bitp) {
            if (tid < mh) reinterpret_cast<uint32_t*>(s_tile)[tid] = tref.valid ? __ldg(tref.bits + tid) : 0u;
            paste_fill_cols_bits<kVec>(reinterpret_cast<uint4*>(s_col_buf), g, mw, tid, kPasteThreads);
        } else {
            tref.fill(s_tile, mh, tid, kPasteThreads);
            cols = paste_fill_cols<kVec>(s_col_buf, g, mw, tid, kPasteThreads);
        }
        __syncthreads();
        uint4* rows_out = reinterpret_cast<uint4*>(out) + (((int64_t)b * M + j) * PH + y0) * spr;
        if (bitp)
            paste_box_rows_bits<kMode>(reinterpret_cast<const uint32_t*>(s_tile), reinterpret_cast<const uint4*>(s_col_buf),
                                       g, mh, y0, nrows, sL, bw, rows_out, spr, tid);
        else
            paste_box_rows<kMode>(s_tile, s_col_buf, cols, g, mh, mw, y0, nrows, sL, bw, rows_out, spr, tid);
    }
}

int paste_launch(mlp_ctx* ctx, const int32_t* det_i32_dev, const PasteSrc& S, int batch, int m_rows,
                 int m_stride, int mask_h, int mask_w, int frame_h, int frame_w, int out_mode,
                 void* out_dev, cudaStream_t st, const BoxItems* boxes = nullptr) {
    ProfScope prof(ctx, MLP_ST_PASTE, st);
    const int vec = out_mode == MLP_PASTE_F32 ? 4 : (out_mode == MLP_PASTE_U8 ? 16 : 128);
    int band_kb = 64;                                  // tuning knobs (tools/bench_paste.py)
    int ctas_per_sm = 0;                               // 0: one CTA per item (default)
    if (const char* e = getenv("MLP_PASTE_CTAS_PER_SM")) ctas_per_sm = atoi(e);
    if (const char* e = getenv("MLP_PASTE_BAND_KB")) band_kb = atoi(e) > 0 ? atoi(e) : 64;
    if (boxes) {
        // behind a background fill: only the box segments, from the work-item list of the tail preparation
        MLP_CHECK_ARG(frame_w % vec == 0, "paste: a prefilled output needs a frame width that is a multiple of %d", vec);
        const int grid = ctx->sm_count * 8;
        const size_t smem = paste_smem_bytes(mask_h, mask_w, frame_w);
#define MLP_PASTE_LAUNCH(MODE)                                                                                 \
    paste_boxes_kernel<MODE><<<grid, kPasteThreads, smem, st>>>(det_i32_dev, S, batch, m_rows, mask_h, mask_w,    \
                                                            frame_h, frame_w, boxes->items, boxes->n_items,    \
                                                            boxes->cap, out_dev)
        if (out_mode == MLP_PASTE_U8) MLP_PASTE_LAUNCH(MLP_PASTE_U8);
        else if (out_mode == MLP_PASTE_BITS) MLP_PASTE_LAUNCH(MLP_PASTE_BITS);
        else MLP_PASTE_LAUNCH(MLP_PASTE_F32);
#undef MLP_PASTE_LAUNCH
    } else if (frame_w % vec == 0) {
        const int row_bytes = frame_w / vec * 16;
        int band_rows = (band_kb * 1024) / row_bytes;
        if (band_rows < 1) band_rows = 1;
        if (band_rows > frame_h) band_rows = frame_h;
        const int bands = (frame_h + band_rows - 1) / band_rows;
        const int64_t items = (int64_t)batch * m_rows * bands;
        // one CTA per item, numbered (band, slot, image) by a 3-D grid; a 1-D grid-stride loop for the tuning knob and
        // for shapes beyond the grid limits
        const bool grid3 = ctas_per_sm <= 0 && m_rows <= 65535 && batch <= 65535;
        dim3 grid(bands, m_rows, batch);
        if (!grid3) grid = dim3((unsigned)(ctas_per_sm > 0 ? ctx->sm_count * ctas_per_sm : (items < (1ll << 30) ? items : (1ll << 30))));
        size_t smem = paste_smem_bytes(mask_h, mask_w, frame_w);
        int use_cols = 1;
        if (const char* e = getenv("MLP_PASTE_COLS")) use_cols = atoi(e);
        if (const char* e = getenv("MLP_PASTE_SMEM_KB")) {           // tuning knob: fewer resident CTAs per SM
            const size_t want = (size_t)atoi(e) * 1024;
            if (want > smem && want <= 200 * 1024) {
                smem = want;
                if (out_mode == MLP_PASTE_U8)
                    MLP_CUDA(cudaFuncSetAttribute(paste_kernel<MLP_PASTE_U8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                else if (out_mode == MLP_PASTE_BITS)
                    MLP_CUDA(cudaFuncSetAttribute(paste_kernel<MLP_PASTE_BITS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                else
                    MLP_CUDA(cudaFuncSetAttribute(paste_kernel<MLP_PASTE_F32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            }
        }
#define MLP_PASTE_LAUNCH(MODE)                                                                         \
    paste_kernel<MODE><<<grid, kPasteThreads, smem, st>>>(det_i32_dev, S, batch, m_rows, m_stride, mask_h, \
                                                      mask_w, frame_h, frame_w, band_rows, bands, grid3 ? 1 : 0, use_cols, out_dev)
        if (out_mode == MLP_PASTE_U8) MLP_PASTE_LAUNCH(MLP_PASTE_U8);
        else if (out_mode == MLP_PASTE_BITS) MLP_PASTE_LAUNCH(MLP_PASTE_BITS);
        else MLP_PASTE_LAUNCH(MLP_PASTE_F32);
#undef MLP_PASTE_LAUNCH
    } else {
        const int64_t items = (int64_t)batch * m_rows;
        const int grid = (int)(items < (1ll << 30) ? items : (1ll << 30));
#define MLP_PASTE_LAUNCH(MODE)                                                                       \
    paste_scalar_kernel<MODE><<<grid, kPasteThreads, 0, st>>>(det_i32_dev, S, batch, m_rows, m_stride, \
                                                             mask_h, mask_w, frame_h, frame_w, out_dev)
        if (out_mode == MLP_PASTE_U8) MLP_PASTE_LAUNCH(MLP_PASTE_U8);
        else if (out_mode == MLP_PASTE_BITS) MLP_PASTE_LAUNCH(MLP_PASTE_BITS);
        else MLP_PASTE_LAUNCH(MLP_PASTE_F32);
#undef MLP_PASTE_LAUNCH
    }
    MLP_LAUNCH_CHECK(ctx);
    return MLP_OK;
}

int check_paste_args(const char* who, int batch, int mask_h, int mask_w, int frame_h, int frame_w,
                     int out_mode, const void* out_dev) {
    MLP_CHECK_ARG(batch >= 1, "%s: batch=%d", who, batch);
    MLP_CHECK_ARG(mask_h >= 1 && mask_w >= 1 && mask_h * mask_w <= kMaxTile,
                  "%s: mask tile %dx%d too large (max %d elements)", who, mask_h, mask_w, kMaxTile);
    MLP_CHECK_ARG(frame_h >= 1 && frame_w >= 1, "%s: frame %dx%d", who, frame_h, frame_w);
    MLP_CHECK_ARG(out_mode == MLP_PASTE_F32 || out_mode == MLP_PASTE_U8 || out_mode == MLP_PASTE_BITS,
                  "%s: unknown out_mode %d", who, out_mode);
    MLP_CHECK_ARG(out_mode != MLP_PASTE_BITS || frame_w % 8 == 0,
                  "%s: bit-packed output needs a frame width that is a multiple of 8 (got %d)", who, frame_w);
    MLP_CHECK_ARG(mlp_aligned16(out_dev), "%s: out_dev must be 16-byte aligned", who);
    return MLP_OK;
}

}  // namespace

extern "C" int mlp_crop_and_pad_mask(mlp_ctx* ctx, const int32_t* det_i32_dev,
                                     const int32_t* masks_i32_dev, int batch, int m_rows, int m_stride,
                                     const int32_t* m_dev, int mask_h, int mask_w, int frame_h,
                                     int frame_w, int out_mode, void* out_dev, mlp_stream_t stream) {
    MLP_CHECK_ARG(ctx && det_i32_dev && masks_i32_dev && out_dev, "mlp_crop_and_pad_mask: NULL argument");
    MLP_CHECK_ARG(m_rows >= 1 && (m_stride >= m_rows || (m_stride == 0 && m_dev)),
                  "mlp_crop_and_pad_mask: bad shape M=%d stride=%d", m_rows, m_stride);
    int rc = check_paste_args("mlp_crop_and_pad_mask", batch, mask_h, mask_w, frame_h, frame_w, out_mode,
                              out_dev);
    if (rc) return rc;
    DeviceGuard g(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    int32_t* thr_dev = ctx->ctr;        // ctr[0]: paste row-filter threshold
    {
        ProfScope prof(ctx, MLP_ST_PASTE_THR, st);
        paste_threshold_kernel<<<1, 1024, 0, st>>>(det_i32_dev, batch, m_rows, m_stride, m_dev, thr_dev);
        MLP_LAUNCH_CHECK(ctx);
    }
    PasteSrc S;
    memset(&S, 0, sizeof(S));
    S.masks_i32 = masks_i32_dev;
    S.m_dev = m_dev;
    S.thr_dev = thr_dev;
    return paste_launch(ctx, det_i32_dev, S, batch, m_rows, m_stride, mask_h, mask_w, frame_h, frame_w,
                        out_mode, out_dev, st);
}

// Fused second half of the path (a11-a14): two kernels.
extern "C" int mlp_trim_paste(mlp_ctx* ctx, const float* roi_boxes_dev, const float* roi_masks_dev,
                              int batch, int r_rows, const int32_t* r_dev, int mask_h, int mask_w,
                              int num_classes, float ratio_h, float ratio_w, int k_rows, int frame_h,
                              int frame_w, int out_mode, int32_t* det_i32_dev, int32_t* counts_dev,
                              int32_t* m_dev, void* out_dev, mlp_stream_t stream) {
    const int flags = out_mode & ~0xff;
    out_mode &= 0xff;
    const bool prefilled = (flags & MLP_PASTE_PREFILLED) != 0;
    const int planar = (flags & MLP_MASKS_PLANAR) ? 1 : 0;
    MLP_CHECK_ARG((flags & ~(MLP_PASTE_PREFILLED | MLP_MASKS_PLANAR)) == 0, "mlp_trim_paste: unknown flags 0x%x", flags);
    MLP_CHECK_ARG(ctx && roi_boxes_dev && roi_masks_dev && det_i32_dev && counts_dev && m_dev &&
                      (out_dev || out_mode == MLP_PASTE_NONE),
                  "mlp_trim_paste: NULL argument");
    MLP_CHECK_ARG(r_rows >= 1 && k_rows >= 1 && num_classes >= 1, "mlp_trim_paste: bad shape R=%d K=%d C=%d",
                  r_rows, k_rows, num_classes);
    MLP_CHECK_ARG(mlp_aligned16(roi_masks_dev), "mlp_trim_paste: roi_masks_dev must be 16-byte aligned");
    int rc = check_paste_args("mlp_trim_paste", batch, mask_h, mask_w, frame_h, frame_w,
                              out_mode == MLP_PASTE_NONE ? MLP_PASTE_U8 : out_mode, out_dev);
    if (rc) return rc;
    DeviceGuard g(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    FusedTail ft = fused_tail_layout(nullptr, batch, k_rows, mask_h, mask_w);
    rc = mlp_ensure_scratch(ctx, MLP_ARENA_FUSED, ft.bytes);
    if (rc) return rc;
    ft = fused_tail_layout(ctx->arena[MLP_ARENA_FUSED], batch, k_rows, mask_h, mask_w);
    // arrival counter of tail_finish: zero it (stream-ordered, 16 bytes) whenever the layout in the arena changes;
    // afterwards the last CTA of every launch leaves it at zero
    if (ctx->tail_layout_key != ((int64_t)batch << 40 ^ (int64_t)k_rows << 16 ^ (int64_t)mask_h << 8 ^ mask_w) ||
        ctx->tail_layout_base != ctx->arena[MLP_ARENA_FUSED]) {
        MLP_CUDA(cudaMemsetAsync(ft.scalars, 0, 16, st));
        ctx->tail_layout_key = (int64_t)batch << 40 ^ (int64_t)k_rows << 16 ^ (int64_t)mask_h << 8 ^ mask_w;
        ctx->tail_layout_base = ctx->arena[MLP_ARENA_FUSED];
    }
    int32_t* tail_src = ft.tail_src;
    int32_t* confmax = ft.confmax;
    uint32_t* tail_bits = ft.tail_bits;
    ctx->tail_planar = planar;
    BoxItems Q;
    memset(&Q, 0, sizeof(Q));
    if (prefilled && out_mode != MLP_PASTE_NONE) {
        const int vec = out_mode == MLP_PASTE_F32 ? 4 : (out_mode == MLP_PASTE_U8 ? 16 : 128);
        MLP_CHECK_ARG(frame_w % vec == 0, "mlp_trim_paste: a prefilled output needs a frame width that is a multiple of %d", vec);
        MLP_CHECK_ARG(frame_h < 65536, "mlp_trim_paste: a prefilled output needs a frame height below 65536");
        // worst case: every box spans the frame; a chunk holds at least max(1, kBoxSegs / segments per row) rows
        const int spr = frame_w / vec;
        const int rows_min = kBoxSegs / spr > 1 ? kBoxSegs / spr : 1;
        const int64_t cap = (int64_t)batch * k_rows * ((frame_h + rows_min - 1) / rows_min);
        MLP_CHECK_ARG(cap < (1ll << 28), "mlp_trim_paste: work-item list too large");
        rc = mlp_ensure_scratch(ctx, MLP_ARENA_PASTE, 256 + cap * (int64_t)sizeof(uint2));
        if (rc) return rc;
        Q.n_items = static_cast<int32_t*>(ctx->arena[MLP_ARENA_PASTE]);
        Q.items = reinterpret_cast<uint2*>(static_cast<char*>(ctx->arena[MLP_ARENA_PASTE]) + 256);
        Q.cap = (int)cap;
        Q.PH = frame_h; Q.PW = frame_w; Q.vec = vec;
        MLP_CUDA(cudaMemsetAsync(Q.n_items, 0, 4, st));
    }
    {
        ProfScope prof(ctx, MLP_ST_TAIL_FUSED, st);
        // One RoI's mask-head block is contiguous, so when bit tiles are built a warp stages it with one
        // bulk copy (tail_prep_tma_kernel); the register-gather kernel remains for blocks that are not a
        // multiple of 16 bytes or too large to stage, and for the tile-less case (many instances).
        const int px = mask_h * mask_w;
        const int64_t slot_bytes = (int64_t)px * (planar ? 1 : num_classes) * 4;
        bool tma = tail_bits != nullptr && slot_bytes % 16 == 0 && slot_bytes * kPrepTmaWarps <= 190 * 1024 &&
                   r_rows <= kSlotMapCap;
        if (const char* e = getenv("MLP_TAIL_TMA")) tma = tma && atoi(e) != 0;              // A/B knob
        if (tma) {
            // about one tile per warp (measured 1 / 2 / 3 with the copies started by source row: 19.1 / 20.9 / 21.8 us)
            int per_warp = 1;                                                             // tuning knob
            if (const char* e = getenv("MLP_TAIL_SLOTS_PER_WARP")) per_warp = atoi(e) > 0 ? atoi(e) : 1;
            int parts = (k_rows + per_warp * kPrepTmaWarps - 1) / (per_warp * kPrepTmaWarps);
            parts = parts < 1 ? 1 : (parts > 64 ? 64 : parts);
            const size_t smem = (size_t)slot_bytes * kPrepTmaWarps;
            MLP_CUDA(cudaFuncSetAttribute(tail_prep_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)smem));
            tail_prep_tma_kernel<<<dim3(batch, parts), kPrepTmaWarps * 32, smem, st>>>(
                roi_boxes_dev, roi_masks_dev, r_rows, r_dev, k_rows, mask_h, mask_w, num_classes, planar, ratio_h,
                ratio_w, det_i32_dev, tail_src, tail_bits, counts_dev, confmax, (uint32_t)slot_bytes, Q, ft.scalars, m_dev);
        } else {
            // CTAs per image grow with the capacity so that each warp gathers at most ~2 tiles
            int parts = k_rows / 16;
            parts = parts < kPrepParts ? kPrepParts : (parts > 64 ? 64 : parts);
            tail_prep_kernel<<<dim3(batch, parts), kPrepThreads, 0, st>>>(
                roi_boxes_dev, roi_masks_dev, r_rows, r_dev, k_rows, mask_h, mask_w, num_classes, planar, ratio_h,
                ratio_w, det_i32_dev, tail_src, tail_bits, counts_dev, confmax, Q, ft.scalars, m_dev);
        }
        MLP_LAUNCH_CHECK(ctx);
    }
    if (out_mode == MLP_PASTE_NONE) return MLP_OK;       // mlp_tile_summary consumes the prepared tail
    PasteSrc S;
    memset(&S, 0, sizeof(S));
    S.fused = 1;
    S.roi_masks = roi_masks_dev;
    S.tail_src = tail_src;
    S.tail_bits = tail_bits;
    S.r_dev = r_dev;
    S.r_rows = r_rows;
    S.C = num_classes;
    S.planar = planar;
    S.counts = counts_dev;
    S.confmax = confmax;
    S.scalars = ft.scalars;
    S.m_out = m_dev;
    return paste_launch(ctx, det_i32_dev, S, batch, k_rows, k_rows, mask_h, mask_w, frame_h, frame_w,
                        out_mode, out_dev, st, prefilled ? &Q : nullptr);
}

// Background of CropAndPadMask, ahead of time: zero instance rows [m_from, m_to) of out_dev's flat [B,M,PH,PW]
// prefix.  m_to = max(1, max_b counts_dev[b]) (misc.py:235-236) when counts_dev is given, else *m_to_dev;
// m_from = *m_from_dev, 0 when NULL.  Two calls make the overlapped fill of PostProcessPipeline: a SPECULATIVE one
// before the NMS kernels with m_to_dev = the M of the previous batch (any prefix of the output is background
// or box, so zeroing too much of it is only wasted work, never wrong), and a completing one after them with
// m_from_dev = that same scalar and counts_dev = this batch's counts.  Join the stream, then call
// mlp_trim_paste with MLP_PASTE_PREFILLED.  TrimInstances can only drop rows, so the M of the tail never
// exceeds the one reduced from the detection counts.
extern "C" int mlp_paste_prefill(mlp_ctx* ctx, const int32_t* m_from_dev, const int32_t* counts_dev,
                                 const int32_t* m_to_dev, int batch, int k_rows, int frame_h, int frame_w,
                                 int out_mode, void* out_dev, mlp_stream_t stream) {
    MLP_CHECK_ARG(ctx && out_dev && (counts_dev || m_to_dev), "mlp_paste_prefill: NULL argument");
    MLP_CHECK_ARG(batch >= 1 && k_rows >= 1 && frame_h >= 1 && frame_w >= 1, "mlp_paste_prefill: bad shape");
    MLP_CHECK_ARG(out_mode == MLP_PASTE_F32 || out_mode == MLP_PASTE_U8 || out_mode == MLP_PASTE_BITS,
                  "mlp_paste_prefill: unknown out_mode %d", out_mode);
    const int vec = out_mode == MLP_PASTE_F32 ? 4 : (out_mode == MLP_PASTE_U8 ? 16 : 128);
    MLP_CHECK_ARG(frame_w % vec == 0, "mlp_paste_prefill: frame width %d is not a multiple of %d", frame_w, vec);
    MLP_CHECK_ARG(mlp_aligned16(out_dev), "mlp_paste_prefill: out_dev must be 16-byte aligned");
    DeviceGuard g(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope prof(ctx, MLP_ST_PASTE_FILL, st);
    const int64_t inst_vecs = (int64_t)frame_h * (frame_w / vec);
    const int64_t ctas = ((int64_t)batch * k_rows * inst_vecs + kFillVecs - 1) / kFillVecs;
    MLP_CHECK_ARG(ctas < (1ll << 31), "mlp_paste_prefill: output too large");
    paste_fill_kernel<<<(unsigned)ctas, kFillThreads, 0, st>>>(m_from_dev, counts_dev, m_to_dev, batch, k_rows,
                                                             inst_vecs, static_cast<uint4*>(out_dev));
    MLP_LAUNCH_CHECK(ctx);
    return MLP_OK;
}
