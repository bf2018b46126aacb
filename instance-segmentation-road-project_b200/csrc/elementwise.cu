// elementwise.cu — the streaming per-row stages: PriorLayer (a2), RestoreBoxes (a3),
// NormalizeBoxes (a4), MaskDistribute (a9), UpSampleOutput (a12).
// All are HBM-bound: one 128-bit load and one 128-bit store per anchor row, grids
// sized as a multiple of the SM count, grid-stride loops.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

inline int grid_for(const mlp_ctx* ctx, int64_t items, int per_sm = 8) {
    int64_t blocks = (items + kThreads - 1) / kThreads;
    int64_t cap = (int64_t)ctx->sm_count * per_sm;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

// ---- a2 ----------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
prior_layer_kernel(const __grid_constant__ PriorDev P, int batch, int4* __restrict__ out) {
    const int64_t total = (int64_t)batch * P.total;
    for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * kThreads) {
        int n = (int)(i % P.total);
        int4 a = prior_anchor(P, n);
        stg_stream_u4(reinterpret_cast<uint4*>(out + i), make_uint4(a.x, a.y, a.z, a.w));
    }
}

// ---- a3 ----------------------------------------------------------------------
template <bool kPriorF32>
__global__ void __launch_bounds__(kThreads)
restore_boxes_kernel(const float4* __restrict__ loc, const void* __restrict__ prior, int64_t rows,
                     float4* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < rows;
         i += (int64_t)gridDim.x * kThreads) {
        float4 l = ldg_stream_f4(loc + i);
        float4 o;
        if (kPriorF32) {
            float4 p = ldg_stream_f4(reinterpret_cast<const float4*>(prior) + i);
            o.x = __fadd_rn(__fmul_rn(l.x, p.z), p.x);
            o.y = __fadd_rn(__fmul_rn(l.y, p.w), p.y);
            o.z = __fmul_rn(exp_cr(l.z), p.z);
            o.w = __fmul_rn(exp_cr(l.w), p.w);
        } else {
            float4 pf = ldg_stream_f4(reinterpret_cast<const float4*>(prior) + i);
            int4 p = make_int4(__float_as_int(pf.x), __float_as_int(pf.y), __float_as_int(pf.z),
                               __float_as_int(pf.w));
            o = restore_box(l, p);
        }
        stg_stream_f4(out + i, o);
    }
}

__global__ void __launch_bounds__(kThreads)
restore_from_prior_kernel(const __grid_constant__ PriorDev P, const float4* __restrict__ loc,
                          int batch, float4* __restrict__ out) {
    const int64_t total = (int64_t)batch * P.total;
    for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * kThreads) {
        int n = (int)(i % P.total);
        float4 l = ldg_stream_f4(loc + i);
        stg_stream_f4(out + i, restore_box(l, prior_anchor(P, n)));
    }
}

// ---- a4 ----------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
normalize_boxes_kernel(const float* __restrict__ boxes, int64_t rows, int row_stride, float ih,
                       float iw, float4* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < rows;
         i += (int64_t)gridDim.x * kThreads) {
        const float* r = boxes + i * row_stride;
        float cx, cy, w, h;
        if (row_stride == 4) {
            float4 v = ldg_stream_f4(reinterpret_cast<const float4*>(r));
            cx = v.x; cy = v.y; w = v.z; h = v.w;
        } else {
            cx = r[0]; cy = r[1]; w = r[2]; h = r[3];
        }
        float hw = __fdiv_rn(w, 2.0f), hh = __fdiv_rn(h, 2.0f);
        float4 o;                                    // (y1, x1, y2, x2)
        o.x = __fdiv_rn(__fsub_rn(cy, hh), ih);
        o.y = __fdiv_rn(__fsub_rn(cx, hw), iw);
        o.z = __fdiv_rn(__fadd_rn(cy, hh), ih);
        o.w = __fdiv_rn(__fadd_rn(cx, hw), iw);
        stg_stream_f4(out + i, o);
    }
}

// ---- a9 ----------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
mask_distribute_kernel(const float* __restrict__ det, int64_t rows, float base_eps, float max_k,
                       float* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < rows;
         i += (int64_t)gridDim.x * kThreads) {
        const float* r = det + i * 6;
        float v[6];
#pragma unroll
        for (int q = 0; q < 6; ++q) v[q] = r[q];
        float* o = out + i * 7;
        o[0] = level_of(v[0], v[2], v[3], base_eps, max_k);
#pragma unroll
        for (int q = 0; q < 6; ++q) o[q + 1] = v[q];
    }
}

// ---- a12 ---------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
upsample_boxes_kernel(const float* __restrict__ det, int64_t rows, float rh, float rw,
                      int32_t* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < rows;
         i += (int64_t)gridDim.x * kThreads) {
        upsample_row(det + i * 6, rh, rw, out + i * 6);
    }
}

__global__ void __launch_bounds__(kThreads)
threshold_masks_kernel(const float* __restrict__ m, int64_t n, int32_t* __restrict__ out) {
    const int64_t n4 = n >> 2;
    const float4* m4 = reinterpret_cast<const float4*>(m);
    for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < n4;
         i += (int64_t)gridDim.x * kThreads) {
        float4 v = ldg_stream_f4(m4 + i);
        stg_stream_u4(reinterpret_cast<uint4*>(out) + i,
                      make_uint4(v.x > 0.5f, v.y > 0.5f, v.z > 0.5f, v.w > 0.5f));
    }
    for (int64_t i = (n4 << 2) + blockIdx.x * (int64_t)kThreads + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * kThreads)
        out[i] = m[i] > 0.5f;
}

}  // namespace

// =============================================================== host side ====
extern "C" int mlp_prior_layer(mlp_ctx* ctx, const mlp_prior_config* prior, int batch, int height,
                               int width, int32_t* out_dev, mlp_stream_t stream) {
    MLP_CHECK_ARG(ctx && out_dev, "mlp_prior_layer: NULL argument");
    MLP_CHECK_ARG(batch >= 1, "mlp_prior_layer: batch=%d", batch);
    MLP_CHECK_ARG(mlp_aligned16(out_dev), "mlp_prior_layer: out_dev not 16-byte aligned");
    PriorDev P;
    int rc = mlp_build_prior_dev(prior, height, width, &P);
    if (rc) return rc;
    DeviceGuard g(ctx->device);
    ProfScope prof(ctx, MLP_ST_ELEMENTWISE, (cudaStream_t)stream);
    int64_t total = (int64_t)batch * P.total;
    prior_layer_kernel<<<grid_for(ctx, total), kThreads, 0, (cudaStream_t)stream>>>(
        P, batch, reinterpret_cast<int4*>(out_dev));
    MLP_LAUNCH_CHECK(ctx);
    return MLP_OK;
}

extern "C" int mlp_restore_boxes(mlp_ctx* ctx, const float* loc_dev, const void* prior_dev,
                                 int prior_is_f32, int64_t rows, float* out_dev,
                                 mlp_stream_t stream) {
    MLP_CHECK_ARG(ctx && loc_dev && prior_dev && out_dev, "mlp_restore_boxes: NULL argument");
    MLP_CHECK_ARG(rows >= 0, "mlp_restore_boxes: rows=%lld", (long long)rows);
    MLP_CHECK_ARG(mlp_aligned16(loc_dev) && mlp_aligned16(prior_dev) && mlp_aligned16(out_dev),
                  "mlp_restore_boxes: pointers must be 16-byte aligned");
    if (rows == 0) return MLP_OK;
    DeviceGuard g(ctx->device);
    ProfScope prof(ctx, MLP_ST_ELEMENTWISE, (cudaStream_t)stream);
    int grid = grid_for(ctx, rows);
    if (prior_is_f32)
        restore_boxes_kernel<true><<<grid, kThreads, 0, (cudaStream_t)stream>>>(
            reinterpret_cast<const float4*>(loc_dev), prior_dev, rows,
            reinterpret_cast<float4*>(out_dev));
    else
        restore_boxes_kernel<false><<<grid, kThreads, 0, (cudaStream_t)stream>>>(
            reinterpret_cast<const float4*>(loc_dev), prior_dev, rows,
            reinterpret_cast<float4*>(out_dev));
    MLP_LAUNCH_CHECK(ctx);
    return MLP_OK;
}

extern "C" int mlp_restore_boxes_from_prior(mlp_ctx* ctx, const mlp_prior_config* prior,
                                            const float* loc_dev, int batch, int height, int width,
                                            float* out_dev, mlp_stream_t stream) {
    MLP_CHECK_ARG(ctx && loc_dev && out_dev, "mlp_restore_boxes_from_prior: NULL argument");
    MLP_CHECK_ARG(batch >= 1, "mlp_restore_boxes_from_prior: batch=%d", batch);
    MLP_CHECK_ARG(mlp_aligned16(loc_dev) && mlp_aligned16(out_dev),
                  "mlp_restore_boxes_from_prior: pointers must be 16-byte aligned");
    PriorDev P;
    int rc = mlp_build_prior_dev(prior, height, width, &P);
    if (rc) return rc;
    DeviceGuard g(ctx->device);
    ProfScope prof(ctx, MLP_ST_ELEMENTWISE, (cudaStream_t)stream);
    int64_t total = (int64_t)batch * P.total;
    restore_from_prior_kernel<<<grid_for(ctx, total), kThreads, 0, (cudaStream_t)stream>>>(
        P, reinterpret_cast<const float4*>(loc_dev), batch, reinterpret_cast<float4*>(out_dev));
    MLP_LAUNCH_CHECK(ctx);
    return MLP_OK;
}

extern "C" int mlp_normalize_boxes(mlp_ctx* ctx, const float* boxes_dev, int64_t rows, int row_stride,
                                   float image_h, float image_w, float* out_dev,
                                   mlp_stream_t stream) {
    MLP_CHECK_ARG(ctx && boxes_dev && out_dev, "mlp_normalize_boxes: NULL argument");
    MLP_CHECK_ARG(rows >= 0 && row_stride >= 4, "mlp_normalize_boxes: rows=%lld row_stride=%d",
                  (long long)rows, row_stride);
    MLP_CHECK_ARG(mlp_aligned16(boxes_dev) && mlp_aligned16(out_dev),
                  "mlp_normalize_boxes: pointers must be 16-byte aligned");
    if (rows == 0) return MLP_OK;
    DeviceGuard g(ctx->device);
    ProfScope prof(ctx, MLP_ST_ELEMENTWISE, (cudaStream_t)stream);
    normalize_boxes_kernel<<<grid_for(ctx, rows), kThreads, 0, (cudaStream_t)stream>>>(
        boxes_dev, rows, row_stride, image_h, image_w, reinterpret_cast<float4*>(out_dev));
    MLP_LAUNCH_CHECK(ctx);
    return MLP_OK;
}

extern "C" int mlp_mask_distribute(mlp_ctx* ctx, const float* det_dev, int64_t rows, int max_k,
                                   float base_size, float* out_dev, mlp_stream_t stream) {
    MLP_CHECK_ARG(ctx && det_dev && out_dev, "mlp_mask_distribute: NULL argument");
    MLP_CHECK_ARG(rows >= 0 && max_k >= 0, "mlp_mask_distribute: rows=%lld max_k=%d", (long long)rows,
                  max_k);
    if (rows == 0) return MLP_OK;
    DeviceGuard g(ctx->device);
    ProfScope prof(ctx, MLP_ST_DISTRIBUTE, (cudaStream_t)stream);
    float base_eps = (float)((double)base_size + 1e-7);      // python: base_size + K.epsilon()
    mask_distribute_kernel<<<grid_for(ctx, rows), kThreads, 0, (cudaStream_t)stream>>>(
        det_dev, rows, base_eps, (float)max_k, out_dev);
    MLP_LAUNCH_CHECK(ctx);
    return MLP_OK;
}

extern "C" int mlp_upsample_output(mlp_ctx* ctx, const float* det_dev, int64_t rows, float ratio_h,
                                   float ratio_w, int32_t* det_i32_dev, const float* masks_dev,
                                   int64_t mask_elems, int32_t* masks_i32_dev, mlp_stream_t stream) {
    MLP_CHECK_ARG(ctx, "mlp_upsample_output: NULL ctx");
    MLP_CHECK_ARG(rows >= 0 && mask_elems >= 0, "mlp_upsample_output: negative size");
    DeviceGuard g(ctx->device);
    ProfScope prof(ctx, MLP_ST_UPSAMPLE, (cudaStream_t)stream);
    if (rows > 0) {
        MLP_CHECK_ARG(det_dev && det_i32_dev, "mlp_upsample_output: NULL det pointers");
        upsample_boxes_kernel<<<grid_for(ctx, rows), kThreads, 0, (cudaStream_t)stream>>>(
            det_dev, rows, ratio_h, ratio_w, det_i32_dev);
        MLP_LAUNCH_CHECK(ctx);
    }
    if (mask_elems > 0) {
        MLP_CHECK_ARG(masks_dev && masks_i32_dev, "mlp_upsample_output: NULL mask pointers");
        MLP_CHECK_ARG(mlp_aligned16(masks_dev) && mlp_aligned16(masks_i32_dev),
                      "mlp_upsample_output: mask pointers must be 16-byte aligned");
        threshold_masks_kernel<<<grid_for(ctx, mask_elems / 4 + 1), kThreads, 0,
                                 (cudaStream_t)stream>>>(masks_dev, mask_elems, masks_i32_dev);
        MLP_LAUNCH_CHECK(ctx);
    }
    return MLP_OK;
}
