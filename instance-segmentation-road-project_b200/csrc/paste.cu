// paste.cu — CropAndPadMask (a13) with the consumers' > 0.5 threshold (a14) fused in.
//
// Reference: /root/reference/engine/layers/misc.py:358-401 (paste), :457 and :611-615
// (binary mask = pasted > 0.5).  For every instance the reference resizes its 28x28
// int mask to the clipped box with tf.image.resize(align_corners=True) (legacy bilinear,
// restated in oracle/tf_ops.py) and zero-pads to the frame.  Output [B,M,PH,PW] is the
// largest tensor of the whole path (16.8 GB as uint8 at B=32, M=1000, 512x1024), so this
// kernel is a pure streaming-store kernel: every thread owns 16 output bytes (one 128-bit
// st.global per 16 uint8 pixels / 4 float pixels), only pixels inside the clipped box
// evaluate the two-stage lerp from the mask tile held in shared memory, everything else
// is written as zero without touching memory for reads.
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int kPasteThreads = 256;
constexpr int kMaxTile = 64 * 64;         // mask_h * mask_w <= 4096

// threshold = 50 if max(conf) > 50 else -100  (misc.py:366-369); conf = column 5.
__global__ void __launch_bounds__(1024)
paste_threshold_kernel(const int32_t* __restrict__ det, int B, int m_rows, int m_stride,
                       const int32_t* __restrict__ m_dev, int32_t* __restrict__ thr_out) {
    __shared__ int s_max;
    int M = m_dev ? *m_dev : m_rows;
    if (M > m_rows) M = m_rows;
    if (m_stride == 0) m_stride = M;            // compact [B,M,..] layout, M known on device only
    if (threadIdx.x == 0) s_max = INT32_MIN;
    __syncthreads();
    int mx = INT32_MIN;
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

struct PasteGeom {
    int xmin, xmax, ymin, ymax;
    float sy, sx;          // resize scales (in-1)/(out-1) or in/out
    bool active;
};

// misc.py:373-386: box = max(box,1) -> float; ceil(c -/+ s/2) -> int -> clip.
__device__ __forceinline__ PasteGeom paste_geometry(const int32_t* row, int thr, int mh,
                                                    int mw, int PH, int PW) {
    PasteGeom g;
    const int conf = row[5];
    const float cx = (float)max(row[0], 1), cy = (float)max(row[1], 1);
    const float w = (float)max(row[2], 1), h = (float)max(row[3], 1);
    const float hw = __fdiv_rn(w, 2.0f), hh = __fdiv_rn(h, 2.0f);
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

// One work item = (image b, slot j, band of `band_rows` frame rows).  Every thread-store is
// 128 bits: 16 uint8 pixels (kU8) or 4 float pixels.  PW must be a multiple of that vector
// width (host checks; the scalar kernel below handles everything else).  The band is written
// in three phases so that no warp mixes cheap and expensive lanes:
//   A  rows of the band above/below the clipped box: one contiguous byte range, flat
//      unrolled zero stores, no index arithmetic;
//   B1 rows crossing the box: the 16-byte segments left and right of it (zeros);
//   B2 the segments that intersect the box, flattened over ALL threads of the CTA: each
//      evaluates the reference's two-stage lerp per pixel from the mask tile in smem.
template <bool kU8>
__global__ void __launch_bounds__(kPasteThreads)
paste_kernel(const int32_t* __restrict__ det, const int32_t* __restrict__ masks, int B, int m_rows,
             int m_stride, const int32_t* __restrict__ m_dev, const int32_t* __restrict__ thr_dev,
             int mh, int mw, int PH, int PW, int band_rows, void* __restrict__ out) {
    constexpr int kVec = kU8 ? 16 : 4;
    constexpr int kPx = kU8 ? 1 : 4;               // bytes per pixel
    __shared__ float s_tile[kMaxTile];
    int M = m_dev ? *m_dev : m_rows;
    if (M > m_rows) M = m_rows;
    if (m_stride == 0) m_stride = M;
    const int thr = *thr_dev;
    const int bands = (PH + band_rows - 1) / band_rows;
    const int64_t items = (int64_t)B * M * bands;
    const int spr = PW / kVec;                     // 16-byte segments per frame row
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
    // One CTA per (instance, band) item, NOT a persistent grid: on B200 a write-only stream
    // of many short-lived CTAs, each owning one contiguous 32 KB band, reaches ~7.4 TB/s while
    // persistent CTAs top out near 6.3 TB/s (tools/write_bw.cu, profiles/write_bw_r01.txt).
    // The grid is sized for the capacity m_rows; CTAs past the device-side M exit at once.
    constexpr int kTileRegs = 4;                   // covers tiles up to 32x32; larger ones load late
    for (int64_t item = blockIdx.x; item < items; item += gridDim.x) {
        const int inst = (int)(item / bands);
        const int band = (int)(item - (int64_t)inst * bands);
        const int b = inst / M, j = inst - b * M;
        int row[6];
        {
            const int2* p = reinterpret_cast<const int2*>(det + ((int64_t)b * m_stride + j) * 6);
            const int2 a0 = __ldg(p), a1 = __ldg(p + 1), a2 = __ldg(p + 2);
            row[0] = a0.x; row[1] = a0.y; row[2] = a1.x; row[3] = a1.y; row[4] = a2.x; row[5] = a2.y;
        }
        const PasteGeom g = paste_geometry(row, thr, mh, mw, PH, PW);
        const int y0 = band * band_rows;
        const int y1 = min(y0 + band_rows, PH);
        // rows of this band that cross the clipped box: [ya, yb)
        int ya = y1, yb = y1;
        if (g.active) { ya = min(max(g.ymin, y0), y1); yb = min(max(g.ymax, ya), y1); }
        const bool touches = yb > ya;              // block-uniform
        // the mask tile is only needed by phase B2: issue its loads now, park them in
        // shared memory after the zero rows have been streamed out
        int tile_regs[kTileRegs];
        if (touches) {
            const int32_t* m = masks + ((int64_t)b * m_stride + j) * mh * mw;
#pragma unroll
            for (int q = 0; q < kTileRegs; ++q) {
                const int i = tid + q * kPasteThreads;
                tile_regs[q] = (i < mh * mw) ? __ldg(m + i) : 0;
            }
        }
        uint4* band_base = reinterpret_cast<uint4*>(static_cast<unsigned char*>(out) +
                                                    ((int64_t)inst * PH + y0) * PW * kPx);
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
#pragma unroll
        for (int q = 0; q < kTileRegs; ++q) {
            const int i = tid + q * kPasteThreads;
            if (i < mh * mw) s_tile[i] = (float)tile_regs[q];
        }
        if (mh * mw > kTileRegs * kPasteThreads) {
            const int32_t* m = masks + ((int64_t)b * m_stride + j) * mh * mw;
            for (int i = tid + kTileRegs * kPasteThreads; i < mh * mw; i += kPasteThreads)
                s_tile[i] = (float)__ldg(m + i);
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
        __syncthreads();                           // s_tile ready
        // ---- phase B2: segments intersecting the box, flattened over the CTA
        const int nb = (yb - ya) * bw;
        for (int i = tid; i < nb; i += kPasteThreads) {
            const int r = i / bw;
            const int seg = sL + (i - r * bw);
            const int oy = ya + r;
            const int x0 = seg * kVec;
            const float p = __fmul_rn((float)(oy - g.ymin), g.sy);
            const float fl = floorf(p);
            const int ylo = max((int)fl, 0);
            const int yhi = min((int)ceilf(p), mh - 1);
            const float ly = __fsub_rn(p, fl);
            uint4 v;
            if (kU8) {
                uint32_t w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    const int ox = x0 + q;
                    if (ox >= g.xmin && ox < g.xmax) {
                        const float val = paste_value(s_tile, mh, mw, ylo, yhi, ly, ox - g.xmin, g.sx);
                        if (val > 0.5f) w[q >> 2] |= 1u << ((q & 3) * 8);
                    }
                }
                v = make_uint4(w[0], w[1], w[2], w[3]);
            } else {
                float f[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int ox = x0 + q;
                    if (ox >= g.xmin && ox < g.xmax)
                        f[q] = paste_value(s_tile, mh, mw, ylo, yhi, ly, ox - g.xmin, g.sx);
                }
                v = make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]),
                               __float_as_uint(f[3]));
            }
            stg_stream_u4(box_rows + (int64_t)r * spr + seg, v);
        }
    }
}

// Generic fall-back for frame widths that are not a multiple of the vector width.
template <bool kU8>
__global__ void __launch_bounds__(kPasteThreads)
paste_scalar_kernel(const int32_t* __restrict__ det, const int32_t* __restrict__ masks, int B,
                    int m_rows, int m_stride, const int32_t* __restrict__ m_dev,
                    const int32_t* __restrict__ thr_dev, int mh, int mw, int PH, int PW,
                    void* __restrict__ out) {
    __shared__ float s_tile[kMaxTile];
    int M = m_dev ? *m_dev : m_rows;
    if (M > m_rows) M = m_rows;
    if (m_stride == 0) m_stride = M;
    const int thr = *thr_dev;
    const int64_t items = (int64_t)B * M;
    for (int64_t inst = blockIdx.x; inst < items; inst += gridDim.x) {
        const int b = (int)(inst / M), j = (int)(inst - (int64_t)b * M);
        const int32_t* row = det + ((int64_t)b * m_stride + j) * 6;
        const PasteGeom g = paste_geometry(row, thr, mh, mw, PH, PW);
        __syncthreads();
        if (g.active) {
            const int32_t* m = masks + ((int64_t)b * m_stride + j) * mh * mw;
            for (int i = threadIdx.x; i < mh * mw; i += kPasteThreads) s_tile[i] = (float)m[i];
        }
        __syncthreads();
        const int64_t npx = (int64_t)PH * PW;
        for (int64_t i = threadIdx.x; i < npx; i += kPasteThreads) {
            const int oy = (int)(i / PW), ox = (int)(i - (int64_t)oy * PW);
            float val = 0.0f;
            if (g.active && oy >= g.ymin && oy < g.ymax && ox >= g.xmin && ox < g.xmax) {
                const float p = __fmul_rn((float)(oy - g.ymin), g.sy);
                const float fl = floorf(p);
                const int ylo = max((int)fl, 0);
                const int yhi = min((int)ceilf(p), mh - 1);
                val = paste_value(s_tile, mh, mw, ylo, yhi, __fsub_rn(p, fl), ox - g.xmin, g.sx);
            }
            if (kU8) static_cast<unsigned char*>(out)[inst * npx + i] = val > 0.5f;
            else static_cast<float*>(out)[inst * npx + i] = val;
        }
    }
}

}  // namespace

extern "C" int mlp_crop_and_pad_mask(mlp_ctx* ctx, const int32_t* det_i32_dev,
                                     const int32_t* masks_i32_dev, int batch, int m_rows, int m_stride,
                                     const int32_t* m_dev, int mask_h, int mask_w, int frame_h,
                                     int frame_w, int out_mode, void* out_dev, mlp_stream_t stream) {
    MLP_CHECK_ARG(ctx && det_i32_dev && masks_i32_dev && out_dev, "mlp_crop_and_pad_mask: NULL argument");
    MLP_CHECK_ARG(batch >= 1 && m_rows >= 1 && (m_stride >= m_rows || (m_stride == 0 && m_dev)),
                  "mlp_crop_and_pad_mask: bad shape B=%d M=%d stride=%d", batch, m_rows, m_stride);
    MLP_CHECK_ARG(mask_h >= 1 && mask_w >= 1 && mask_h * mask_w <= kMaxTile,
                  "mlp_crop_and_pad_mask: mask tile %dx%d too large (max %d elements)", mask_h, mask_w,
                  kMaxTile);
    MLP_CHECK_ARG(frame_h >= 1 && frame_w >= 1, "mlp_crop_and_pad_mask: frame %dx%d", frame_h, frame_w);
    MLP_CHECK_ARG(out_mode == MLP_PASTE_F32 || out_mode == MLP_PASTE_U8,
                  "mlp_crop_and_pad_mask: unknown out_mode %d", out_mode);
    MLP_CHECK_ARG(mlp_aligned16(out_dev), "mlp_crop_and_pad_mask: out_dev must be 16-byte aligned");
    DeviceGuard g(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    int32_t* thr_dev = ctx->ctr;        // ctr[0]: paste row-filter threshold
    {
        ProfScope prof(ctx, MLP_ST_PASTE_THR, st);
        paste_threshold_kernel<<<1, 1024, 0, st>>>(det_i32_dev, batch, m_rows, m_stride, m_dev, thr_dev);
        MLP_LAUNCH_CHECK(ctx);
    }
    ProfScope prof(ctx, MLP_ST_PASTE, st);
    const bool u8 = out_mode == MLP_PASTE_U8;
    const int vec = u8 ? 16 : 4;
    // persistent grid: exactly the resident CTA count, each CTA streams one contiguous slab
    int occ = 0;
    if (u8) MLP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, paste_kernel<true>, kPasteThreads, 0));
    else MLP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, paste_kernel<false>, kPasteThreads, 0));
    if (occ < 1) occ = 1;
    int ctas_per_sm = 0;                          // 0: one CTA per item (default)
    int band_kb = 64;
    if (const char* e = getenv("MLP_PASTE_CTAS_PER_SM")) ctas_per_sm = atoi(e);     // tuning knobs
    if (const char* e = getenv("MLP_PASTE_BAND_KB")) band_kb = atoi(e) > 0 ? atoi(e) : 64;
    int grid = ctx->sm_count * (ctas_per_sm > 0 ? ctas_per_sm : occ);
    if (frame_w % vec == 0) {
        // bands of ~64 KB of output keep >> grid items in flight even for one small batch
        int band_rows = (band_kb * 1024) / (frame_w * (u8 ? 1 : 4));
        if (band_rows < 1) band_rows = 1;
        if (band_rows > frame_h) band_rows = frame_h;
        if (ctas_per_sm <= 0) {
            const int64_t items = (int64_t)batch * m_rows * ((frame_h + band_rows - 1) / band_rows);
            grid = (int)(items < (1ll << 30) ? items : (1ll << 30));
        }
        if (u8)
            paste_kernel<true><<<grid, kPasteThreads, 0, st>>>(det_i32_dev, masks_i32_dev, batch, m_rows,
                                                              m_stride, m_dev, thr_dev, mask_h, mask_w,
                                                              frame_h, frame_w, band_rows, out_dev);
        else
            paste_kernel<false><<<grid, kPasteThreads, 0, st>>>(det_i32_dev, masks_i32_dev, batch, m_rows,
                                                               m_stride, m_dev, thr_dev, mask_h, mask_w,
                                                               frame_h, frame_w, band_rows, out_dev);
    } else {
        if (u8)
            paste_scalar_kernel<true><<<grid, kPasteThreads, 0, st>>>(
                det_i32_dev, masks_i32_dev, batch, m_rows, m_stride, m_dev, thr_dev, mask_h, mask_w,
                frame_h, frame_w, out_dev);
        else
            paste_scalar_kernel<false><<<grid, kPasteThreads, 0, st>>>(
                det_i32_dev, masks_i32_dev, batch, m_rows, m_stride, m_dev, thr_dev, mask_h, mask_w,
                frame_h, frame_w, out_dev);
    }
    MLP_LAUNCH_CHECK(ctx);
    return MLP_OK;
}
