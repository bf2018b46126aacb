// roi.cu — PyramidRoiAlign (a10) and TrimInstances (a11).
//
// PyramidRoiAlign (/root/reference/engine/layers/instance.py:109-139): every kept box
// is cropped from the FPN level MaskDistribute assigned it to with TF crop_and_resize
// arithmetic (bilinear, extrapolation 0; restated in oracle/tf_ops.py), then MoldBatch
// regroups crops per level as [B,Mf,ch,cw,Cf] padded with -1.  Here:
//   plan : one CTA per image ranks its boxes per level (ballot prefix scan, order j)
//          -> slot -> source-row table, per-level counts and Mf = max(1, max_b count).
//   run  : persistent CTAs, one RoI (level, image, slot) at a time; each warp owns one
//          output pixel, each lane a float4 of channels: the 4 bilinear corners are four
//          fully coalesced 16*32-byte NHWC reads (L1-allocating: neighbouring output
//          pixels of an up-sampled RoI share corners) and the result one streaming
//          128-bit store per lane.  Padded slots are filled with -1 the same way.
// The stage is write-dominated (B*M*196*Cf*4 bytes out vs <= the FPN maps in).
#include "common.cuh"

namespace {

constexpr int kRoiThreads = 256;
constexpr int kMaxCrop = 64;          // crop_h, crop_w <= 64

struct RoiLevels {
    const float* fmap[MLP_MAX_LEVELS];
    float* crops[MLP_MAX_LEVELS];
    int fh[MLP_MAX_LEVELS];
    int fw[MLP_MAX_LEVELS];
};

// ---- plan ---------------------------------------------------------------------
// roi_src [L][B][m_rows] : source row j of (level, image, slot); counts [L][B]; level_m [L].
__global__ void __launch_bounds__(32 * MLP_MAX_LEVELS)
roi_plan_kernel(const float* __restrict__ dist, int B, int m_rows, int m_stride,
                const int32_t* __restrict__ m_dev, int L, int32_t* __restrict__ roi_src,
                int32_t* __restrict__ counts, int32_t* __restrict__ level_m) {
    const int b = blockIdx.x;
    const int f = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (f >= L) return;
    int M = m_dev ? *m_dev : m_rows;
    if (M > m_rows) M = m_rows;
    const float* rows = dist + (int64_t)b * m_stride * 7;
    int32_t* src = roi_src + ((int64_t)f * B + b) * m_rows;
    int base = 0;
    for (int j0 = 0; j0 < M; j0 += 32) {
        const int j = j0 + lane;
        const bool hit = (j < M) && (rows[(int64_t)j * 7] == (float)f);
        const unsigned mask = __ballot_sync(0xffffffffu, hit);
        if (hit) src[base + __popc(mask & ((1u << lane) - 1u))] = j;
        base += __popc(mask);
    }
    if (lane == 0) {
        counts[f * B + b] = base;
        atomicMax(level_m + f, base > 1 ? base : 1);
    }
}

// ---- run ----------------------------------------------------------------------
constexpr int kPixUnroll = 4;         // output pixels in flight per warp (16 corner loads)
constexpr int kMaxPix = 1024;         // crop_h * crop_w <= 1024 (per-RoI pixel table in smem)

__global__ void __launch_bounds__(kRoiThreads)
roi_align_kernel(const RoiLevels lv, int L, int Cf, const float* __restrict__ dist, int B,
                 int m_rows, int m_stride, float image_h, float image_w, int ch, int cw,
                 const int32_t* __restrict__ roi_src, const int32_t* __restrict__ counts,
                 int32_t* __restrict__ level_m, float* __restrict__ roi_boxes) {
    // per-RoI tables, built once per RoI by ch+cw (coordinates) and ch*cw (pixels) threads so
    // that the streaming loop below carries no divisions and no 64-bit address arithmetic
    __shared__ float s_iny[kMaxCrop], s_inx[kMaxCrop];
    __shared__ uint4 s_poff[kMaxPix];         // BYTE offsets of TL,TR,BL,BR inside this image's map
    __shared__ float2 s_pw[kMaxPix];          // (lx, ly); lx < 0 marks "outside -> 0"
    __shared__ int s_mf[MLP_MAX_LEVELS], s_off[MLP_MAX_LEVELS + 1];
    if (threadIdx.x == 0) {
        int off = 0;
        for (int f = 0; f < L; ++f) {
            const int m = level_m[f];
            s_mf[f] = m; s_off[f] = off;
            off += m;
        }
        s_off[L] = off;
        if (blockIdx.x == 0) level_m[L] = off;               // R = sum of Mf, for TrimInstances
    }
    __syncthreads();
    const int R = s_off[L];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nwarps = kRoiThreads / 32;
    const int npix = ch * cw;
    const int C4 = Cf >> 2;
    const bool vec = (Cf & 3) == 0;
    // One CTA per (level, image, slot) over the CAPACITY m_rows (not a persistent grid: the
    // stage is write-dominated and short-lived CTAs stream best, see paste.cu); slots past the
    // device-side Mf exit at once.
    const int64_t items = (int64_t)L * B * m_rows;

    for (int64_t item = blockIdx.x; item < items; item += gridDim.x) {
        const int slot = (int)(item % m_rows);
        const int fb = (int)(item / m_rows);
        const int f = fb / B, b = fb - f * B;
        const int mf = s_mf[f];
        if (slot >= mf) continue;
        const int cnt = counts[f * B + b];
        float* out = lv.crops[f] + ((int64_t)b * mf + slot) * npix * Cf;
        float* rb = roi_boxes + ((int64_t)b * R + s_off[f] + slot) * 6;

        if (slot >= cnt) {                                   // MoldBatch padding
            if (threadIdx.x < 6) rb[threadIdx.x] = -1.0f;
            const int64_t n = (int64_t)npix * Cf;
            if (vec) {
                float4* o4 = reinterpret_cast<float4*>(out);
                const float4 m1 = make_float4(-1.f, -1.f, -1.f, -1.f);
                for (int64_t i = threadIdx.x; i < (n >> 2); i += kRoiThreads) stg_stream_f4(o4 + i, m1);
            } else {
                for (int64_t i = threadIdx.x; i < n; i += kRoiThreads) out[i] = -1.0f;
            }
            continue;
        }
        const int j = roi_src[((int64_t)f * B + b) * m_rows + slot];
        const float* row = dist + ((int64_t)b * m_stride + j) * 7;
        const int Hf = lv.fh[f], Wf = lv.fw[f];
        __syncthreads();                                     // previous RoI done with the tables
        if (threadIdx.x < 6) rb[threadIdx.x] = row[1 + threadIdx.x];
        if (threadIdx.x < ch + cw) {
            // NormalizeBoxes(shape=image) then crop_and_resize source coordinates
            const bool is_y = threadIdx.x < ch;
            const int idx = is_y ? threadIdx.x : threadIdx.x - ch;
            const float c = is_y ? row[2] : row[1];          // cy : cx
            const float s = is_y ? row[4] : row[3];          // h  : w
            const float dim = is_y ? image_h : image_w;
            const float half = __fdiv_rn(s, 2.0f);
            const float lo = __fdiv_rn(__fsub_rn(c, half), dim);      // y1 : x1
            const float hi = __fdiv_rn(__fadd_rn(c, half), dim);      // y2 : x2
            const int nout = is_y ? ch : cw;
            const float fm1 = (float)((is_y ? Hf : Wf) - 1);
            float in;
            if (nout > 1) {
                const float scale = __fdiv_rn(__fmul_rn(__fsub_rn(hi, lo), fm1), (float)(nout - 1));
                in = __fadd_rn(__fmul_rn(lo, fm1), __fmul_rn((float)idx, scale));
            } else {
                in = __fmul_rn(__fmul_rn(0.5f, __fadd_rn(lo, hi)), fm1);
            }
            if (is_y) s_iny[idx] = in; else s_inx[idx] = in;
        }
        __syncthreads();
        {
            const float hm1 = (float)(Hf - 1), wm1 = (float)(Wf - 1);
            for (int p = threadIdx.x; p < npix; p += kRoiThreads) {
                const int y = p / cw, x = p - y * cw;
                const float in_y = s_iny[y], in_x = s_inx[x];
                uint4 o = make_uint4(0u, 0u, 0u, 0u);
                float2 w = make_float2(-1.0f, 0.0f);
                // TF: in < 0 || in > size-1 -> extrapolation value; NaN counts as outside
                if (in_y >= 0.0f && in_y <= hm1 && in_x >= 0.0f && in_x <= wm1) {
                    const float fy = floorf(in_y), fx = floorf(in_x);
                    const int top = (int)fy, bot = (int)ceilf(in_y);
                    const int left = (int)fx, right = (int)ceilf(in_x);
                    w = make_float2(__fsub_rn(in_x, fx), __fsub_rn(in_y, fy));
                    o = make_uint4((unsigned)((top * Wf + left) * Cf) * 4u, (unsigned)((top * Wf + right) * Cf) * 4u,
                                   (unsigned)((bot * Wf + left) * Cf) * 4u, (unsigned)((bot * Wf + right) * Cf) * 4u);
                }
                s_poff[p] = o;
                s_pw[p] = w;
            }
        }
        __syncthreads();
        const float* img = lv.fmap[f] + (int64_t)b * Hf * Wf * Cf;
        // each warp owns kPixUnroll consecutive output pixels per iteration
        for (int p0 = warp * kPixUnroll; p0 < npix; p0 += nwarps * kPixUnroll) {
            if (vec) {
                for (int c4 = lane; c4 < C4; c4 += 32) {
                    float4 tl[kPixUnroll], tr[kPixUnroll], bl[kPixUnroll], br[kPixUnroll];
                    float2 w[kPixUnroll];
                    // 64-bit lane base + 32-bit unsigned byte offsets: two adds per address
                    const char* base = reinterpret_cast<const char*>(img) + c4 * 16;
#pragma unroll
                    for (int u = 0; u < kPixUnroll; ++u) {            // 16 loads in flight
                        const int p = min(p0 + u, npix - 1);
                        const uint4 o = s_poff[p];
                        w[u] = s_pw[p];
                        tl[u] = __ldg(reinterpret_cast<const float4*>(base + o.x));
                        tr[u] = __ldg(reinterpret_cast<const float4*>(base + o.y));
                        bl[u] = __ldg(reinterpret_cast<const float4*>(base + o.z));
                        br[u] = __ldg(reinterpret_cast<const float4*>(base + o.w));
                    }
                    float4* obase = reinterpret_cast<float4*>(out + (int64_t)p0 * Cf) + c4;
#pragma unroll
                    for (int u = 0; u < kPixUnroll; ++u) {
                        const float lx = w[u].x, ly = w[u].y;
                        const bool inside = lx >= 0.0f;               // else extrapolation_value = 0
                        // bilinear: top/bottom lerp in x, then lerp in y (crop_and_resize order),
                        // adds packed two channels per instruction
                        const uint64_t t0 = lerp2_rn(pack2(tl[u].x, tl[u].y), pack2(tr[u].x, tr[u].y), lx);
                        const uint64_t t1 = lerp2_rn(pack2(tl[u].z, tl[u].w), pack2(tr[u].z, tr[u].w), lx);
                        const uint64_t b0 = lerp2_rn(pack2(bl[u].x, bl[u].y), pack2(br[u].x, br[u].y), lx);
                        const uint64_t b1 = lerp2_rn(pack2(bl[u].z, bl[u].w), pack2(br[u].z, br[u].w), lx);
                        float4 r;
                        unpack2(lerp2_rn(t0, b0, ly), r.x, r.y);
                        unpack2(lerp2_rn(t1, b1, ly), r.z, r.w);
                        if (!inside) r = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (p0 + u < npix) stg_stream_f4(obase + (size_t)u * C4, r);
                    }
                }
            } else {
#pragma unroll
                for (int u = 0; u < kPixUnroll; ++u) {
                    const int p = p0 + u;
                    if (p >= npix) continue;
                    const uint4 o = s_poff[p];
                    const float2 w = s_pw[p];
                    float* op = out + p * Cf;
                    for (int c = lane; c < Cf; c += 32) {
                        float r = 0.0f;
                        if (w.x >= 0.0f) {
                            const float tl = img[(o.x >> 2) + c], tr = img[(o.y >> 2) + c], bl = img[(o.z >> 2) + c], br = img[(o.w >> 2) + c];
                            const float t_ = __fadd_rn(tl, __fmul_rn(__fsub_rn(tr, tl), w.x));
                            const float b_ = __fadd_rn(bl, __fmul_rn(__fsub_rn(br, bl), w.x));
                            r = __fadd_rn(t_, __fmul_rn(__fsub_rn(b_, t_), w.y));
                        }
                        op[c] = r;
                    }
                }
            }
        }
    }
}

// ---- TrimInstances ------------------------------------------------------------
// plan: counts[b] = #rows with class != -1 among the first R rows; M = max(1, max_b).
__global__ void __launch_bounds__(256)
trim_plan_kernel(const float* __restrict__ roi_boxes, int r_rows, const int32_t* __restrict__ r_dev,
                 int32_t* __restrict__ counts, int32_t* __restrict__ m_dev) {
    __shared__ int s_cnt;
    const int b = blockIdx.x;
    const int R = r_dev ? *r_dev : r_rows;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    const float* rows = roi_boxes + (int64_t)b * R * 6;
    int c = 0;
    for (int j = threadIdx.x; j < R; j += blockDim.x) c += rows[(int64_t)j * 6 + 4] != -1.0f;
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s_cnt, c);
    __syncthreads();
    if (threadIdx.x == 0) {
        counts[b] = s_cnt;
        atomicMax(m_dev, s_cnt > 1 ? s_cnt : 1);
    }
}

// index: one warp per image ranks valid rows in order j -> trim_src[b][slot] = j, and
// writes the box rows (valid or -1 padded) of out_boxes [B,M,6].
__global__ void __launch_bounds__(32)
trim_index_kernel(const float* __restrict__ roi_boxes, int r_rows, const int32_t* __restrict__ r_dev,
                  const int32_t* __restrict__ m_dev, int32_t* __restrict__ trim_src, int src_stride,
                  float* __restrict__ out_boxes) {
    const int b = blockIdx.x, lane = threadIdx.x;
    const int R = r_dev ? *r_dev : r_rows;
    const int M = *m_dev;
    const float* rows = roi_boxes + (int64_t)b * R * 6;
    int32_t* src = trim_src + (int64_t)b * src_stride;
    float* ob = out_boxes + (int64_t)b * M * 6;
    int base = 0;
    for (int j0 = 0; j0 < R; j0 += 32) {
        const int j = j0 + lane;
        const bool hit = (j < R) && (rows[(int64_t)j * 6 + 4] != -1.0f);
        const unsigned mask = __ballot_sync(0xffffffffu, hit);
        if (hit) {
            const int slot = base + __popc(mask & ((1u << lane) - 1u));
            src[slot] = j;
#pragma unroll
            for (int q = 0; q < 6; ++q) ob[(int64_t)slot * 6 + q] = rows[(int64_t)j * 6 + q];
        }
        base += __popc(mask);
    }
    for (int i = base * 6 + lane; i < M * 6; i += 32) ob[i] = -1.0f;
    for (int s = base + lane; s < M; s += 32) src[s] = -1;
}

// gather: one CTA per (image, slot): out_masks[b,slot,:,:] = roi_masks[b,j,:,:,class].
__global__ void __launch_bounds__(256)
trim_gather_kernel(const float* __restrict__ roi_boxes, const float* __restrict__ roi_masks, int B,
                   int r_rows, const int32_t* __restrict__ r_dev, int mask_px, int C,
                   const int32_t* __restrict__ m_dev, const int32_t* __restrict__ trim_src,
                   int src_stride, float* __restrict__ out_masks) {
    const int R = r_dev ? *r_dev : r_rows;
    const int M = *m_dev;
    const int64_t items = (int64_t)B * M;
    for (int64_t item = blockIdx.x; item < items; item += gridDim.x) {
        const int b = (int)(item / M), slot = (int)(item - (int64_t)b * M);
        const int j = trim_src[(int64_t)b * src_stride + slot];
        float* o = out_masks + item * mask_px;
        if (j < 0) {
            for (int p = threadIdx.x; p < mask_px; p += blockDim.x) o[p] = -1.0f;
            continue;
        }
        const int cls = (int)roi_boxes[((int64_t)b * R + j) * 6 + 4];
        const float* m = roi_masks + ((int64_t)b * R + j) * mask_px * C;
        if (cls < 0 || cls >= C) {       // tf.gather_nd would raise; define as -1 fill
            for (int p = threadIdx.x; p < mask_px; p += blockDim.x) o[p] = -1.0f;
            continue;
        }
        for (int p = threadIdx.x; p < mask_px; p += blockDim.x) o[p] = __ldg(m + (int64_t)p * C + cls);
    }
}

}  // namespace

// ================================================================ host side ===
extern "C" int mlp_roi_align_plan(mlp_ctx* ctx, const float* dist_dev, int batch, int m_rows,
                                  int m_stride, const int32_t* m_dev, int num_levels,
                                  int32_t* level_counts_dev, int32_t* level_m_dev,
                                  mlp_stream_t stream) {
    MLP_CHECK_ARG(ctx && dist_dev && level_counts_dev && level_m_dev, "mlp_roi_align_plan: NULL argument");
    MLP_CHECK_ARG(batch >= 1 && m_rows >= 1 && m_stride >= m_rows,
                  "mlp_roi_align_plan: bad shape B=%d m_rows=%d m_stride=%d", batch, m_rows, m_stride);
    MLP_CHECK_ARG(num_levels >= 1 && num_levels <= MLP_MAX_LEVELS,
                  "mlp_roi_align_plan: num_levels=%d out of range", num_levels);
    DeviceGuard g(ctx->device);
    int rc = mlp_ensure_scratch(ctx, MLP_ARENA_ROI, (int64_t)num_levels * batch * m_rows * 4);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope prof(ctx, MLP_ST_ROI_PLAN, st);
    MLP_CUDA(cudaMemsetAsync(level_m_dev, 0, (size_t)(num_levels + 1) * 4, st));
    roi_plan_kernel<<<batch, 32 * MLP_MAX_LEVELS, 0, st>>>(
        dist_dev, batch, m_rows, m_stride, m_dev, num_levels,
        static_cast<int32_t*>(ctx->arena[MLP_ARENA_ROI]), level_counts_dev, level_m_dev);
    MLP_LAUNCH_CHECK(ctx);
    return MLP_OK;
}

extern "C" int mlp_roi_align_run(mlp_ctx* ctx, const float* const* fmaps_dev, const int32_t* fh,
                                 const int32_t* fw, int num_levels, int channels,
                                 const float* dist_dev, int batch, int m_rows, int m_stride,
                                 const int32_t* m_dev, float image_h, float image_w, int crop_h,
                                 int crop_w, const int32_t* level_counts_dev,
                                 int32_t* level_m_dev, float* const* crops_dev,
                                 float* roi_boxes_dev, mlp_stream_t stream) {
    (void)m_dev;
    MLP_CHECK_ARG(ctx && fmaps_dev && fh && fw && dist_dev && level_counts_dev && level_m_dev &&
                      crops_dev && roi_boxes_dev,
                  "mlp_roi_align_run: NULL argument");
    MLP_CHECK_ARG(num_levels >= 1 && num_levels <= MLP_MAX_LEVELS,
                  "mlp_roi_align_run: num_levels=%d out of range", num_levels);
    MLP_CHECK_ARG(channels >= 1 && batch >= 1 && m_rows >= 1 && m_stride >= m_rows,
                  "mlp_roi_align_run: bad shape");
    MLP_CHECK_ARG(crop_h >= 1 && crop_w >= 1 && crop_h <= kMaxCrop && crop_w <= kMaxCrop &&
                      crop_h + crop_w <= kRoiThreads && crop_h * crop_w <= kMaxPix,
                  "mlp_roi_align_run: crop size %dx%d out of range [1,%d]", crop_h, crop_w, kMaxCrop);
    MLP_CHECK_ARG(ctx->arena[MLP_ARENA_ROI] &&
                      ctx->arena_bytes[MLP_ARENA_ROI] >= (int64_t)num_levels * batch * m_rows * 4,
                  "mlp_roi_align_run: call mlp_roi_align_plan with the same shapes first");
    RoiLevels lv;
    memset(&lv, 0, sizeof(lv));
    for (int f = 0; f < num_levels; ++f) {
        MLP_CHECK_ARG(fmaps_dev[f] && crops_dev[f], "mlp_roi_align_run: NULL level %d pointer", f);
        MLP_CHECK_ARG(mlp_aligned16(fmaps_dev[f]) && mlp_aligned16(crops_dev[f]),
                      "mlp_roi_align_run: level %d pointers must be 16-byte aligned", f);
        MLP_CHECK_ARG(fh[f] >= 1 && fw[f] >= 1, "mlp_roi_align_run: level %d map is %dx%d", f, fh[f], fw[f]);
        MLP_CHECK_ARG((int64_t)fh[f] * fw[f] * channels < (1ll << 29) &&
                          (int64_t)crop_h * crop_w * channels < (1ll << 31),
                      "mlp_roi_align_run: level %d map too large for 32-bit offsets", f);
        lv.fmap[f] = fmaps_dev[f];
        lv.crops[f] = crops_dev[f];
        lv.fh[f] = fh[f];
        lv.fw[f] = fw[f];
    }
    DeviceGuard g(ctx->device);
    ProfScope prof(ctx, MLP_ST_ROI_ALIGN, (cudaStream_t)stream);
    const int64_t items = (int64_t)num_levels * batch * m_rows;
    const int grid = (int)(items < (1ll << 30) ? items : (1ll << 30));
    roi_align_kernel<<<grid, kRoiThreads, 0, (cudaStream_t)stream>>>(
        lv, num_levels, channels, dist_dev, batch, m_rows, m_stride, image_h, image_w, crop_h, crop_w,
        static_cast<const int32_t*>(ctx->arena[MLP_ARENA_ROI]), level_counts_dev,
        level_m_dev, roi_boxes_dev);
    MLP_LAUNCH_CHECK(ctx);
    return MLP_OK;
}

extern "C" int mlp_trim_plan(mlp_ctx* ctx, const float* roi_boxes_dev, int batch, int r_rows,
                             const int32_t* r_dev, int32_t* counts_dev, int32_t* m_dev,
                             mlp_stream_t stream) {
    MLP_CHECK_ARG(ctx && roi_boxes_dev && counts_dev && m_dev, "mlp_trim_plan: NULL argument");
    MLP_CHECK_ARG(batch >= 1 && r_rows >= 1, "mlp_trim_plan: bad shape B=%d R=%d", batch, r_rows);
    DeviceGuard g(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope prof(ctx, MLP_ST_TRIM, st);
    MLP_CUDA(cudaMemsetAsync(m_dev, 0, 4, st));
    trim_plan_kernel<<<batch, 256, 0, st>>>(roi_boxes_dev, r_rows, r_dev, counts_dev, m_dev);
    MLP_LAUNCH_CHECK(ctx);
    return MLP_OK;
}

extern "C" int mlp_trim_run(mlp_ctx* ctx, const float* roi_boxes_dev, const float* roi_masks_dev,
                            int batch, int r_rows, const int32_t* r_dev, int mask_h, int mask_w,
                            int num_classes, const int32_t* m_dev, float* out_boxes_dev,
                            float* out_masks_dev, mlp_stream_t stream) {
    MLP_CHECK_ARG(ctx && roi_boxes_dev && roi_masks_dev && m_dev && out_boxes_dev && out_masks_dev,
                  "mlp_trim_run: NULL argument");
    MLP_CHECK_ARG(batch >= 1 && r_rows >= 1 && mask_h >= 1 && mask_w >= 1 && num_classes >= 1,
                  "mlp_trim_run: bad shape");
    DeviceGuard g(ctx->device);
    int rc = mlp_ensure_scratch(ctx, MLP_ARENA_TRIM, (int64_t)batch * r_rows * 4);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    int32_t* trim_src = static_cast<int32_t*>(ctx->arena[MLP_ARENA_TRIM]);
    ProfScope prof(ctx, MLP_ST_TRIM, st);
    trim_index_kernel<<<batch, 32, 0, st>>>(roi_boxes_dev, r_rows, r_dev, m_dev, trim_src, r_rows,
                                           out_boxes_dev);
    MLP_LAUNCH_CHECK(ctx);
    trim_gather_kernel<<<ctx->sm_count * 8, 256, 0, st>>>(roi_boxes_dev, roi_masks_dev, batch, r_rows,
                                                         r_dev, mask_h * mask_w, num_classes, m_dev,
                                                         trim_src, r_rows, out_masks_dev);
    MLP_LAUNCH_CHECK(ctx);
    return MLP_OK;
}
