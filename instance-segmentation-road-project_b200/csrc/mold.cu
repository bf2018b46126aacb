// mold.cu — MoldBatch (/root/reference/engine/layers/misc.py:231-286) as a standalone
// operator: rows x[K, row_elems] tagged with an image id are regrouped as
// [B, M, row_elems], input order kept inside every image, -1 padded,
// M = max(1, max_b count_b).  (The fused stages never call this: DetectionProposal,
// PyramidRoiAlign and TrimInstances write their molded layout directly.)
#include "common.cuh"

namespace {

// One warp per image: slot -> source row table (ballot prefix scan over all K rows).
__global__ void __launch_bounds__(32)
mold_plan_kernel(const int32_t* __restrict__ batch_idx, int64_t K, int32_t* __restrict__ src_of,
                 int32_t* __restrict__ counts, int32_t* __restrict__ m_dev) {
    const int b = blockIdx.x, lane = threadIdx.x;
    int32_t* src = src_of + (int64_t)b * K;
    int base = 0;
    for (int64_t i0 = 0; i0 < K; i0 += 32) {
        const int64_t i = i0 + lane;
        const bool hit = (i < K) && (batch_idx[i] == b);
        const unsigned mask = __ballot_sync(0xffffffffu, hit);
        if (hit) src[base + __popc(mask & ((1u << lane) - 1u))] = (int32_t)i;
        base += __popc(mask);
    }
    if (lane == 0) {
        counts[b] = base;
        atomicMax(m_dev, base > 1 ? base : 1);
    }
}

__global__ void __launch_bounds__(256)
mold_run_kernel(const uint32_t* __restrict__ x, int64_t K, int64_t row_elems, int B, uint32_t pad,
                const int32_t* __restrict__ src_of, const int32_t* __restrict__ counts,
                const int32_t* __restrict__ m_dev, uint32_t* __restrict__ out) {
    const int M = *m_dev;
    const int64_t items = (int64_t)B * M;
    const bool vec = (row_elems & 3) == 0;
    for (int64_t item = blockIdx.x; item < items; item += gridDim.x) {
        const int b = (int)(item / M), slot = (int)(item - (int64_t)b * M);
        uint32_t* o = out + item * row_elems;
        if (slot >= counts[b]) {
            if (vec) {
                const uint4 p4 = make_uint4(pad, pad, pad, pad);
                for (int64_t e = threadIdx.x; e < (row_elems >> 2); e += blockDim.x)
                    stg_stream_u4(reinterpret_cast<uint4*>(o) + e, p4);
            } else {
                for (int64_t e = threadIdx.x; e < row_elems; e += blockDim.x) o[e] = pad;
            }
            continue;
        }
        const uint32_t* r = x + (int64_t)src_of[(int64_t)b * K + slot] * row_elems;
        if (vec) {
            for (int64_t e = threadIdx.x; e < (row_elems >> 2); e += blockDim.x) {
                const float4 v = ldg_stream_f4(reinterpret_cast<const float4*>(r) + e);
                stg_stream_f4(reinterpret_cast<float4*>(o) + e, v);
            }
        } else {
            for (int64_t e = threadIdx.x; e < row_elems; e += blockDim.x) o[e] = r[e];
        }
    }
}

}  // namespace

extern "C" int mlp_mold_batch_plan(mlp_ctx* ctx, const int32_t* batch_idx_dev, int64_t rows, int batch,
                                   int32_t* counts_dev, int32_t* m_dev, mlp_stream_t stream) {
    MLP_CHECK_ARG(ctx && counts_dev && m_dev, "mlp_mold_batch_plan: NULL argument");
    MLP_CHECK_ARG(rows >= 0 && batch >= 1, "mlp_mold_batch_plan: bad shape K=%lld B=%d", (long long)rows,
                  batch);
    MLP_CHECK_ARG(rows == 0 || batch_idx_dev, "mlp_mold_batch_plan: NULL batch indices");
    MLP_CHECK_ARG(rows < (1ll << 31), "mlp_mold_batch_plan: too many rows");
    DeviceGuard g(ctx->device);
    int rc = mlp_ensure_scratch(ctx, MLP_ARENA_MOLD, (int64_t)batch * (rows > 0 ? rows : 1) * 4);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope prof(ctx, MLP_ST_MOLD, st);
    MLP_CUDA(cudaMemsetAsync(m_dev, 0, 4, st));
    mold_plan_kernel<<<batch, 32, 0, st>>>(batch_idx_dev, rows,
                                          static_cast<int32_t*>(ctx->arena[MLP_ARENA_MOLD]), counts_dev,
                                          m_dev);
    MLP_LAUNCH_CHECK(ctx);
    return MLP_OK;
}

extern "C" int mlp_mold_batch_run(mlp_ctx* ctx, const void* x_dev, const int32_t* counts_dev,
                                  int64_t rows, int64_t row_elems, int batch, int pad_is_float,
                                  const int32_t* m_dev, void* out_dev, mlp_stream_t stream) {
    MLP_CHECK_ARG(ctx && counts_dev && m_dev && out_dev, "mlp_mold_batch_run: NULL argument");
    MLP_CHECK_ARG(rows >= 0 && row_elems >= 1 && batch >= 1, "mlp_mold_batch_run: bad shape");
    MLP_CHECK_ARG(rows == 0 || x_dev, "mlp_mold_batch_run: NULL rows");
    MLP_CHECK_ARG(mlp_aligned16(out_dev) && mlp_aligned16(x_dev),
                  "mlp_mold_batch_run: pointers must be 16-byte aligned");
    MLP_CHECK_ARG(ctx->arena[MLP_ARENA_MOLD] &&
                      ctx->arena_bytes[MLP_ARENA_MOLD] >= (int64_t)batch * (rows > 0 ? rows : 1) * 4,
                  "mlp_mold_batch_run: call mlp_mold_batch_plan with the same shapes first");
    DeviceGuard g(ctx->device);
    ProfScope prof(ctx, MLP_ST_MOLD, (cudaStream_t)stream);
    const float m1 = -1.0f;
    uint32_t pad = pad_is_float ? *reinterpret_cast<const uint32_t*>(&m1) : 0xffffffffu;
    mold_run_kernel<<<ctx->sm_count * 8, 256, 0, (cudaStream_t)stream>>>(
        static_cast<const uint32_t*>(x_dev), rows, row_elems, batch, pad,
        static_cast<const int32_t*>(ctx->arena[MLP_ARENA_MOLD]), counts_dev, m_dev,
        static_cast<uint32_t*>(out_dev));
    MLP_LAUNCH_CHECK(ctx);
    return MLP_OK;
}
