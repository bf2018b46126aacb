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
__device__ __forceinline__ PasteGeom paste_geometry(const int32_t* __restrict__ row, int thr, int mh,
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

// One work item = (image b, slot j, band of `band_rows` frame rows).  kU8: 16 pixels per
// thread-store; else 4 float pixels per thread-store.  PW must be a multiple of the
// vector width (host checks; the scalar kernel below handles everything else).
template <bool kU8>
__global__ void __launch_bounds__(kPasteThreads)
paste_kernel(const int32_t* __restrict__ det, const int32_t* __restrict__ masks, int B, int m_rows,
             int m_stride, const int32_t* __restrict__ m_dev, const int32_t* __restrict__ thr_dev,
             int mh, int mw, int PH, int PW, int band_rows, void* __restrict__ out) {
    constexpr int kVec = kU8 ? 16 : 4;
    __shared__ float s_tile[kMaxTile];
    int M = m_dev ? *m_dev : m_rows;
    if (M > m_rows) M = m_rows;
    if (m_stride == 0) m_stride = M;
    const int thr = *thr_dev;
    const int bands = (PH + band_rows - 1) / band_rows;
    const int64_t items = (int64_t)B * M * bands;
    const int segs_per_row = PW / kVec;
    int tile_owner = -1;                       // instance whose mask is in s_tile

    for (int64_t item = blockIdx.x; item < items; item += gridDim.x) {
        const int inst = (int)(item / bands);
        const int band = (int)(item - (int64_t)inst * bands);
        const int b = inst / M, j = inst - b * M;
        const int32_t* row = det + ((int64_t)b * m_stride + j) * 6;
        const PasteGeom g = paste_geometry(row, thr, mh, mw, PH, PW);
        const int y0 = band * band_rows;
        const int y1 = min(y0 + band_rows, PH);
        const bool touches = g.active && y0 < g.ymax && y1 > g.ymin;
        if (touches && tile_owner != inst) {
            __syncthreads();
            const int32_t* m = masks + ((int64_t)b * m_stride + j) * mh * mw;
            for (int i = threadIdx.x; i < mh * mw; i += kPasteThreads) s_tile[i] = (float)m[i];
            tile_owner = inst;
            __syncthreads();
        }
        unsigned char* obase = static_cast<unsigned char*>(out) +
                               ((int64_t)inst * PH + y0) * PW * (kU8 ? 1 : 4);
        const int nseg = (y1 - y0) * segs_per_row;
        for (int s = threadIdx.x; s < nseg; s += kPasteThreads) {
            const int ry = s / segs_per_row;
            const int x0 = (s - ry * segs_per_row) * kVec;
            const int oy = y0 + ry;
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (touches && oy >= g.ymin && oy < g.ymax && x0 < g.xmax && x0 + kVec > g.xmin) {
                const float p = __fmul_rn((float)(oy - g.ymin), g.sy);
                const float fl = floorf(p);
                const int ylo = max((int)fl, 0);
                const int yhi = min((int)ceilf(p), mh - 1);
                const float ly = __fsub_rn(p, fl);
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
            }
            stg_stream_u4(reinterpret_cast<uint4*>(obase) + s, v);
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
    // persistent grid, 8 CTAs of 256 threads per SM (write-only: occupancy hides store latency)
    const int grid = ctx->sm_count * 8;
    if (frame_w % vec == 0) {
        // bands of ~64 KB of output keep >> grid items in flight even for one small batch
        int band_rows = (64 * 1024) / (frame_w * (u8 ? 1 : 4));
        if (band_rows < 1) band_rows = 1;
        if (band_rows > frame_h) band_rows = frame_h;
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
