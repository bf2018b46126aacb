// resize.cu — the bilinear resizes either side of the path (resize part):
// DownSampleInput (/root/reference/engine/layers/misc.py:133-161) and the semantic half of
// UpSampleOutput (:190-195: resize to the frame, > 0.5, int32).  Both are
// tf.compat.v1.image.resize_bilinear(align_corners=True) - the legacy ResizeBilinear kernel
// (resize_bilinear_op.cc + image_resizer_state.h, half_pixel_centers = false), restated in
// oracle/tf_ops.py / oracle/semantic_oracle.py: scale = (in-1)/(out-1) (in/out when out == 1),
// source coordinate = index * scale, lower = floor, upper = min(ceil, in-1), two-stage lerp
// top/bottom in x then in y, every float32 operation rounded separately (-fmad=false).
//
// One thread per output pixel, all channels (NHWC: the channels of a pixel are contiguous, the four
// source pixels are read as contiguous channel runs).  Up-sampling is write-bound (int32/f32 out),
// down-sampling read-bound; there is no reuse worth staging - neighbouring threads hit the same
// source lines in L1.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int kResizeThreads = 256;

__device__ __forceinline__ float src_f32(const float* p) { return __ldg(p); }
__device__ __forceinline__ float src_f32(const uint8_t* p) { return (float)__ldg(p); }
__device__ __forceinline__ float src_f32(const int32_t* p) { return (float)__ldg(p); }

template <typename InT, bool kThreshold>
__global__ void __launch_bounds__(kResizeThreads)
resize_bilinear_kernel(const InT* __restrict__ in, int B, int ih, int iw, int S, int oh, int ow, float sy, float sx,
                       void* __restrict__ out) {
    const int64_t npix = (int64_t)B * oh * ow;
    for (int64_t p = (int64_t)blockIdx.x * kResizeThreads + threadIdx.x; p < npix;
         p += (int64_t)gridDim.x * kResizeThreads) {
        const int x = (int)(p % ow);
        const int64_t t = p / ow;
        const int y = (int)(t % oh), b = (int)(t / oh);
        const float py = __fmul_rn((float)y, sy), px = __fmul_rn((float)x, sx);
        const float fy = floorf(py), fx = floorf(px);
        const int ylo = max((int)fy, 0), yhi = min((int)ceilf(py), ih - 1);
        const int xlo = max((int)fx, 0), xhi = min((int)ceilf(px), iw - 1);
        const float ly = __fsub_rn(py, fy), lx = __fsub_rn(px, fx);
        const InT* img = in + (int64_t)b * ih * iw * S;
        const InT* tl = img + ((int64_t)ylo * iw + xlo) * S;
        const InT* tr = img + ((int64_t)ylo * iw + xhi) * S;
        const InT* bl = img + ((int64_t)yhi * iw + xlo) * S;
        const InT* br = img + ((int64_t)yhi * iw + xhi) * S;
        for (int c = 0; c < S; ++c) {
            const float a = src_f32(tl + c), bq = src_f32(tr + c), cq = src_f32(bl + c), d = src_f32(br + c);
            const float top = __fadd_rn(a, __fmul_rn(__fsub_rn(bq, a), lx));
            const float bot = __fadd_rn(cq, __fmul_rn(__fsub_rn(d, cq), lx));
            const float v = __fadd_rn(top, __fmul_rn(__fsub_rn(bot, top), ly));
            if (kThreshold) static_cast<int32_t*>(out)[p * S + c] = v > 0.5f ? 1 : 0;
            else static_cast<float*>(out)[p * S + c] = v;
        }
    }
}

// Four consecutive output pixels of a row per thread (out_w % 4 == 0, S <= 4): the y terms are
// shared, the 4*S results leave as S 128-bit streaming stores.
template <typename InT, bool kThreshold, int S>
__global__ void __launch_bounds__(kResizeThreads)
resize_bilinear4_kernel(const InT* __restrict__ in, int B, int ih, int iw, int oh, int ow, float sy, float sx,
                        void* __restrict__ out) {
    const int owq = ow >> 2;
    const int64_t nq = (int64_t)B * oh * owq;
    for (int64_t i = (int64_t)blockIdx.x * kResizeThreads + threadIdx.x; i < nq;
         i += (int64_t)gridDim.x * kResizeThreads) {
        const int xq = (int)(i % owq);
        const int64_t t = i / owq;
        const int y = (int)(t % oh), b = (int)(t / oh);
        const float py = __fmul_rn((float)y, sy);
        const float fy = floorf(py);
        const int ylo = max((int)fy, 0), yhi = min((int)ceilf(py), ih - 1);
        const float ly = __fsub_rn(py, fy);
        const InT* r0 = in + ((int64_t)b * ih + ylo) * iw * S;
        const InT* r1 = in + ((int64_t)b * ih + yhi) * iw * S;
        float v[4 * S];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float px = __fmul_rn((float)(xq * 4 + q), sx);
            const float fx = floorf(px);
            const int xlo = max((int)fx, 0), xhi = min((int)ceilf(px), iw - 1);
            const float lx = __fsub_rn(px, fx);
#pragma unroll
            for (int c = 0; c < S; ++c) {
                const float a = src_f32(r0 + xlo * S + c), bq = src_f32(r0 + xhi * S + c);
                const float cq = src_f32(r1 + xlo * S + c), d = src_f32(r1 + xhi * S + c);
                const float top = __fadd_rn(a, __fmul_rn(__fsub_rn(bq, a), lx));
                const float bot = __fadd_rn(cq, __fmul_rn(__fsub_rn(d, cq), lx));
                v[q * S + c] = __fadd_rn(top, __fmul_rn(__fsub_rn(bot, top), ly));
            }
        }
        uint4* o = reinterpret_cast<uint4*>(out) + i * S;       // 4*S values = S x 16 bytes
#pragma unroll
        for (int k = 0; k < S; ++k) {
            uint32_t w[4];
#pragma unroll
            for (int e = 0; e < 4; ++e)
                w[e] = kThreshold ? (v[k * 4 + e] > 0.5f ? 1u : 0u) : __float_as_uint(v[k * 4 + e]);
            stg_stream_u4(o + k, make_uint4(w[0], w[1], w[2], w[3]));
        }
    }
}

template <typename InT, bool kThreshold>
void launch_resize4(int S, int grid, cudaStream_t st, const void* in, int B, int ih, int iw, int oh, int ow, float sy,
                    float sx, void* out) {
    const InT* p = static_cast<const InT*>(in);
    if (S == 1) resize_bilinear4_kernel<InT, kThreshold, 1><<<grid, kResizeThreads, 0, st>>>(p, B, ih, iw, oh, ow, sy, sx, out);
    else if (S == 2) resize_bilinear4_kernel<InT, kThreshold, 2><<<grid, kResizeThreads, 0, st>>>(p, B, ih, iw, oh, ow, sy, sx, out);
    else if (S == 3) resize_bilinear4_kernel<InT, kThreshold, 3><<<grid, kResizeThreads, 0, st>>>(p, B, ih, iw, oh, ow, sy, sx, out);
    else resize_bilinear4_kernel<InT, kThreshold, 4><<<grid, kResizeThreads, 0, st>>>(p, B, ih, iw, oh, ow, sy, sx, out);
}

// ---- SemanticSmoothing ----------------------------------------------------------------------
// Grey erosion then dilation with a flat k x k structuring element, padding SAME
// (/root/reference/engine/layers/semantic.py:270-284; tf.nn.erosion2d is -dilation2d(-v, reverse(k)),
// so with the all-zero kernel both read the window p-(k-1)/2 .. p-(k-1)/2+k-1 and skip positions
// outside the map; restated in oracle/semantic_oracle.py).  min/max are exact, so each 2-D window
// is evaluated separably: four streaming passes (min along x, min along y, max along x, max along y),
// one thread per element, k L1-served taps each; the last pass applies the weight.
// Element i = (outer * len + p) * step + inner; a thread owns FOUR consecutive positions p of one
// (outer, inner) line, so neighbouring outputs share their taps (k + 3 loads instead of 4 k).
template <bool kMax>
__global__ void __launch_bounds__(kResizeThreads)
window_pass_kernel(const float* __restrict__ in, int64_t lines_outer, int len, int64_t step, int k, int pt,
                   float weight, float* __restrict__ out) {
    const int groups = (len + 3) >> 2;
    const int64_t total = lines_outer * groups * step;
    const float neutral = kMax ? -INFINITY : INFINITY;
    for (int64_t t = (int64_t)blockIdx.x * kResizeThreads + threadIdx.x; t < total;
         t += (int64_t)gridDim.x * kResizeThreads) {
        const int64_t inner = t % step;
        const int64_t og = t / step;
        const int p0 = (int)(og % groups) * 4;
        const int64_t outer = og / groups;
        const float* line = in + outer * len * step + inner;
        float acc[4] = {neutral, neutral, neutral, neutral};
        const int lo = max(p0 - pt, 0), hi = min(p0 + 3 + (k - 1 - pt), len - 1);
        for (int s = lo; s <= hi; ++s) {                    // source position; feeds outputs s-(k-1-pt) .. s+pt
            const float v = __ldg(line + (int64_t)s * step);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int d = s - (p0 + q);
                if (d >= -pt && d <= k - 1 - pt) acc[q] = kMax ? fmaxf(acc[q], v) : fminf(acc[q], v);
            }
        }
        float* o = out + outer * len * step + inner;
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (p0 + q < len) o[(int64_t)(p0 + q) * step] = __fmul_rn(acc[q], weight);
    }
}

// The same opening in ONE pass over the map for window sizes up to kFusedMaxK: a CTA owns a 32 x 64 tile of one
// channel, stages it with a halo of two windows (2 pt before, 2 pb after) in shared memory and runs the four separable
// passes there - every element is read from and written to HBM once (the halo comes out of L2) instead of four times.
// Positions outside the map are +inf for the erosion and -inf for the dilation (the reference skips them).  Every
// pass gives a thread four consecutive outputs of a line (k + 3 taps for 4 outputs); tasks are numbered so that the
// lanes of a warp sit on neighbouring lines (row pitch odd): no bank conflicts in either direction.
constexpr int kFusedMaxK = 16;
constexpr int kFusedTh = 32, kFusedTw = 64;
constexpr int kFusedRh = kFusedTh + 2 * (kFusedMaxK - 1), kFusedRw = kFusedTw + 2 * (kFusedMaxK - 1);
constexpr int kFusedPitch = kFusedRw | 1;

// Four consecutive outputs p0 .. p0+3 of a line from the K + 3 taps src[base + j * kStep], j = 0 .. K+2 (output q
// covers taps q .. q+K-1).  The taps all four share (3 .. K-1) are reduced once: K + 6 min/max for four outputs.  K
// and the step are compile-time: the taps are loads at immediate offsets, no loop.  Tap indices are clamped to jmax
// (only the last three can pass it, and only for outputs nobody stores).
template <bool kMax, int K, int kStep>
__device__ __forceinline__ void window4(const float* __restrict__ src, int jmax, float (&o)[4]) {
    auto op = [](float a, float b) { return kMax ? fmaxf(a, b) : fminf(a, b); };
    auto tap = [&](int j) { return src[min(j, jmax) * kStep]; };
    if (K >= 4) {
        float c = src[3 * kStep];
#pragma unroll
        for (int j = 4; j < K; ++j) c = op(c, src[j * kStep]);
        const float t0 = src[0], t1 = src[kStep], t2 = src[2 * kStep];
        const float u0 = tap(K), u1 = tap(K + 1), u2 = tap(K + 2);
        const float a = op(t1, t2), b = op(u0, u1);
        o[0] = op(c, op(t0, a));
        o[1] = op(c, op(a, u0));
        o[2] = op(c, op(t2, b));
        o[3] = op(c, op(b, u2));
    } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float v = tap(q);
#pragma unroll
            for (int j = 1; j < K; ++j) v = op(v, tap(q + j));
            o[q] = v;
        }
    }
}

// dst[r][c] = min / max of src over the window [p - pt, p - pt + K - 1] along x (kAlongX) or y, for r in [r0, r1),
// c in [c0, c1); lim: valid extent of src along the window direction
template <bool kMax, bool kAlongX, int K>
__device__ __forceinline__ void fused_window_pass(const float* __restrict__ src, float* __restrict__ dst, int r0, int r1,
                                                  int c0, int c1, int lim) {
    constexpr int pt = (K - 1) / 2;
    const int nr = r1 - r0, nc = c1 - c0;
    const int lines = kAlongX ? nr : nc, len = kAlongX ? nc : nr;
    const int groups = (len + 3) >> 2;
    // task t = group * lines + line, walked with a stride of the CTA size: one division per pass, then additions
    const int dl = kResizeThreads % lines, dg = kResizeThreads / lines;
    int line = (int)threadIdx.x % lines, grp = (int)threadIdx.x / lines;
    for (; grp < groups; line += dl, grp += dg) {
        if (line >= lines) { line -= lines; ++grp; if (grp >= groups) break; }
        const int p0 = (kAlongX ? c0 : r0) + grp * 4;
        const int fixed = (kAlongX ? r0 : c0) + line;
        const int first = p0 - pt;                                             // tap 0
        float acc[4];
        window4<kMax, K, kAlongX ? 1 : kFusedPitch>(
            src + (kAlongX ? fixed * kFusedPitch + first : first * kFusedPitch + fixed), lim - 1 - first, acc);
        const int pend = kAlongX ? c1 : r1;
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (p0 + q < pend) {
                if (kAlongX) dst[fixed * kFusedPitch + p0 + q] = acc[q];
                else dst[(p0 + q) * kFusedPitch + fixed] = acc[q];
            }
    }
}

template <int K>
__global__ void __launch_bounds__(kResizeThreads)
smoothing_fused_kernel(const float* __restrict__ in, int H, int W, int S, float weight, float* __restrict__ out) {
    __shared__ float sa[kFusedRh * kFusedPitch];
    __shared__ float sb[kFusedRh * kFusedPitch];
    constexpr int pt = (K - 1) / 2, pb = K - 1 - pt;
    constexpr int rh = kFusedTh + 2 * (K - 1), rw = kFusedTw + 2 * (K - 1);   // region actually used
    const int b = blockIdx.z / S, c = blockIdx.z - b * S;
    const int ty0 = blockIdx.y * kFusedTh, tx0 = blockIdx.x * kFusedTw;
    const int oy = ty0 - 2 * pt, ox = tx0 - 2 * pt;                            // map coordinates of region (0, 0)
    const float* img = in + (int64_t)b * H * W * S + c;
    // stage the region (a warp per row, lanes along x); outside the map: +inf (skipped by the erosion)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool inside = oy >= 0 && oy + rh <= H && ox >= 0 && ox + rw <= W;    // CTA-uniform: no position to skip
    for (int ry = warp; ry < rh; ry += kResizeThreads / 32) {
        const int y = oy + ry;
        float* srow = sa + ry * kFusedPitch;
        if (inside) {
            const float* rowp = img + ((int64_t)y * W + ox) * S;
            for (int rx = lane; rx < rw; rx += 32) srow[rx] = __ldg(rowp + rx * S);
        } else {
            const bool yin = y >= 0 && y < H;
            for (int rx = lane; rx < rw; rx += 32) {
                const int x = ox + rx;
                srow[rx] = (yin && x >= 0 && x < W) ? __ldg(img + ((int64_t)y * W + x) * S) : INFINITY;
            }
        }
    }
    __syncthreads();
    // erosion along x: every region row, columns [pt, rw - pb)
    fused_window_pass<false, true, K>(sa, sb, 0, rh, pt, rw - pb, rw);
    __syncthreads();
    // erosion along y: rows [pt, rh - pb)
    fused_window_pass<false, false, K>(sb, sa, pt, rh - pb, pt, rw - pb, rh);
    __syncthreads();
    // eroded values outside the map do not exist: -inf (skipped by the dilation); interior tiles have none
    if (!inside) {
        for (int ry = pt + warp; ry < rh - pb; ry += kResizeThreads / 32) {
            const int y = oy + ry;
            for (int rx = pt + lane; rx < rw - pb; rx += 32) {
                const int x = ox + rx;
                if (y < 0 || y >= H || x < 0 || x >= W) sa[ry * kFusedPitch + rx] = -INFINITY;
            }
        }
        __syncthreads();
    }
    // dilation along x: rows [pt, rh - pb), columns of the tile [2 pt, 2 pt + Tw)
    fused_window_pass<true, true, K>(sa, sb, pt, rh - pb, 2 * pt, 2 * pt + kFusedTw, rw);
    __syncthreads();
    // dilation along y straight to the map, times the weight
    {
        constexpr int groups = kFusedTh / 4;
        float* o = out + (int64_t)b * H * W * S + c;
        for (int t = threadIdx.x; t < kFusedTw * groups; t += kResizeThreads) {
            const int col = t % kFusedTw, g = t / kFusedTw;
            const int x = tx0 + col;
            const int r0 = 2 * pt + g * 4;                                     // region row of the first output
            float acc[4];
            window4<true, K, kFusedPitch>(sb + (r0 - pt) * kFusedPitch + 2 * pt + col, rh - 1 - (r0 - pt), acc);
            if (x < W) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int y = ty0 + g * 4 + q;
                    if (y < H) o[((int64_t)y * W + x) * S] = __fmul_rn(acc[q], weight);
                }
            }
        }
    }
}

template <int K>
void launch_smoothing_fused(dim3 grid, cudaStream_t st, const float* in, int H, int W, int S, float weight, float* out) {
    smoothing_fused_kernel<K><<<grid, kResizeThreads, 0, st>>>(in, H, W, S, weight, out);
}

__global__ void __launch_bounds__(kResizeThreads)
scale_kernel(const float* __restrict__ in, int64_t n, float weight, float* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * kResizeThreads + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * kResizeThreads)
        out[i] = __fmul_rn(__ldg(in + i), weight);
}

}  // namespace

extern "C" int mlp_semantic_smoothing(mlp_ctx* ctx, const float* in_dev, int batch, int height, int width,
                                      int channels, int kernel_size, float weight, float* out_dev,
                                      mlp_stream_t stream) {
    MLP_CHECK_ARG(ctx && in_dev && out_dev, "mlp_semantic_smoothing: NULL argument");
    MLP_CHECK_ARG(batch >= 1 && height >= 1 && width >= 1 && channels >= 1, "mlp_semantic_smoothing: bad shape");
    MLP_CHECK_ARG(kernel_size <= 1024, "mlp_semantic_smoothing: kernel_size %d too large", kernel_size);
    DeviceGuard g(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope prof(ctx, MLP_ST_RESIZE, st);
    const int64_t n = (int64_t)batch * height * width * channels;
    int64_t blocks = (n + kResizeThreads - 1) / kResizeThreads;
    const int64_t cap = (int64_t)ctx->sm_count * 32;
    const int grid = (int)(blocks < cap ? blocks : cap);
    if (kernel_size <= 0) {                                  // semantic.py:283-284
        scale_kernel<<<grid, kResizeThreads, 0, st>>>(in_dev, n, weight, out_dev);
        MLP_LAUNCH_CHECK(ctx);
        return MLP_OK;
    }
    if (kernel_size <= kFusedMaxK && (int64_t)batch * channels <= 65535 && !getenv("MLP_SMOOTH_PASSES")) {
        const dim3 fgrid((width + kFusedTw - 1) / kFusedTw, (height + kFusedTh - 1) / kFusedTh, batch * channels);
#define MLP_SMOOTH_K(K) case K: launch_smoothing_fused<K>(fgrid, st, in_dev, height, width, channels, weight, out_dev); break;
        switch (kernel_size) {
            MLP_SMOOTH_K(1) MLP_SMOOTH_K(2) MLP_SMOOTH_K(3) MLP_SMOOTH_K(4) MLP_SMOOTH_K(5) MLP_SMOOTH_K(6)
            MLP_SMOOTH_K(7) MLP_SMOOTH_K(8) MLP_SMOOTH_K(9) MLP_SMOOTH_K(10) MLP_SMOOTH_K(11) MLP_SMOOTH_K(12)
            MLP_SMOOTH_K(13) MLP_SMOOTH_K(14) MLP_SMOOTH_K(15) MLP_SMOOTH_K(16)
        }
#undef MLP_SMOOTH_K
        MLP_LAUNCH_CHECK(ctx);
        return MLP_OK;
    }
    int rc = mlp_ensure_scratch(ctx, MLP_ARENA_SMOOTH, n * 4);
    if (rc) return rc;
    float* tmp = static_cast<float*>(ctx->arena[MLP_ARENA_SMOOTH]);
    const int k = kernel_size, pt = (kernel_size - 1) / 2;
    const int64_t sx = channels, sy = (int64_t)width * channels;
    const int64_t outer_x = (int64_t)batch * height, outer_y = batch;      // lines along x / along y
    auto grid_for = [&](int64_t outer, int len, int64_t step) {
        const int64_t work = outer * ((len + 3) / 4) * step;
        const int64_t bl = (work + kResizeThreads - 1) / kResizeThreads;
        return (int)(bl < cap ? (bl < 1 ? 1 : bl) : cap);
    };
    window_pass_kernel<false><<<grid_for(outer_x, width, sx), kResizeThreads, 0, st>>>(in_dev, outer_x, width, sx, k, pt, 1.0f, tmp);
    MLP_LAUNCH_CHECK(ctx);
    window_pass_kernel<false><<<grid_for(outer_y, height, sy), kResizeThreads, 0, st>>>(tmp, outer_y, height, sy, k, pt, 1.0f, out_dev);
    MLP_LAUNCH_CHECK(ctx);
    window_pass_kernel<true><<<grid_for(outer_x, width, sx), kResizeThreads, 0, st>>>(out_dev, outer_x, width, sx, k, pt, 1.0f, tmp);
    MLP_LAUNCH_CHECK(ctx);
    window_pass_kernel<true><<<grid_for(outer_y, height, sy), kResizeThreads, 0, st>>>(tmp, outer_y, height, sy, k, pt, weight, out_dev);
    MLP_LAUNCH_CHECK(ctx);
    return MLP_OK;
}

extern "C" int mlp_resize_bilinear(mlp_ctx* ctx, const void* in_dev, int in_dtype, int batch, int in_h, int in_w,
                                   int channels, int out_h, int out_w, int flags, void* out_dev,
                                   mlp_stream_t stream) {
    const int threshold = flags & MLP_RESIZE_THRESHOLD;
    const bool align = !(flags & MLP_RESIZE_NO_ALIGN_CORNERS);
    MLP_CHECK_ARG(ctx && in_dev && out_dev, "mlp_resize_bilinear: NULL argument");
    MLP_CHECK_ARG(batch >= 1 && in_h >= 1 && in_w >= 1 && channels >= 1, "mlp_resize_bilinear: bad input shape");
    MLP_CHECK_ARG(out_h >= 1 && out_w >= 1, "mlp_resize_bilinear: output dimensions must be positive (%dx%d)", out_h,
                  out_w);
    MLP_CHECK_ARG(in_dtype == MLP_F32 || in_dtype == MLP_U8 || in_dtype == MLP_I32,
                  "mlp_resize_bilinear: input must be f32, u8 or i32");
    DeviceGuard g(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope prof(ctx, MLP_ST_RESIZE, st);
    // CalculateResizeScale(in, out, align_corners)
    const float sy = (align && out_h > 1) ? (float)(in_h - 1) / (float)(out_h - 1) : (float)in_h / (float)out_h;
    const float sx = (align && out_w > 1) ? (float)(in_w - 1) / (float)(out_w - 1) : (float)in_w / (float)out_w;
    const int64_t npix = (int64_t)batch * out_h * out_w;
    const int64_t cap = (int64_t)ctx->sm_count * 32;
    if (channels <= 4 && (out_w & 3) == 0 && mlp_aligned16(out_dev)) {
        const int64_t blocks4 = (npix / 4 + kResizeThreads - 1) / kResizeThreads;
        const int grid4 = (int)(blocks4 < cap ? (blocks4 < 1 ? 1 : blocks4) : cap);
#define MLP_RESIZE4(T)                                                                                              \
    do {                                                                                                            \
        if (threshold) launch_resize4<T, true>(channels, grid4, st, in_dev, batch, in_h, in_w, out_h, out_w, sy, sx, out_dev); \
        else launch_resize4<T, false>(channels, grid4, st, in_dev, batch, in_h, in_w, out_h, out_w, sy, sx, out_dev);          \
    } while (0)
        if (in_dtype == MLP_F32) MLP_RESIZE4(float);
        else if (in_dtype == MLP_U8) MLP_RESIZE4(uint8_t);
        else MLP_RESIZE4(int32_t);
#undef MLP_RESIZE4
        MLP_LAUNCH_CHECK(ctx);
        return MLP_OK;
    }
    int64_t blocks = (npix + kResizeThreads - 1) / kResizeThreads;
    const int grid = (int)(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
#define MLP_RESIZE(T, THR)                                                                                   \
    resize_bilinear_kernel<T, THR><<<grid, kResizeThreads, 0, st>>>(static_cast<const T*>(in_dev), batch, in_h, \
                                                                    in_w, channels, out_h, out_w, sy, sx, out_dev)
    if (in_dtype == MLP_F32) { if (threshold) MLP_RESIZE(float, true); else MLP_RESIZE(float, false); }
    else if (in_dtype == MLP_U8) { if (threshold) MLP_RESIZE(uint8_t, true); else MLP_RESIZE(uint8_t, false); }
    else { if (threshold) MLP_RESIZE(int32_t, true); else MLP_RESIZE(int32_t, false); }
#undef MLP_RESIZE
    MLP_LAUNCH_CHECK(ctx);
    return MLP_OK;
}
