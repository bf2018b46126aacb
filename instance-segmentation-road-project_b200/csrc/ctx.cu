// ctx.cu — context, error reporting, DLPack validation, prior table helpers.
#include <stdarg.h>

#include "common.cuh"

static thread_local char g_err[512] = "";

void mlp_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* mlp_last_error(void) { return g_err; }
extern "C" int mlp_version(void) { return MLP_VERSION; }

extern "C" int mlp_ctx_create(int device, mlp_ctx** out) {
    MLP_CHECK_ARG(out != nullptr, "mlp_ctx_create: out is NULL");
    *out = nullptr;
    int ndev = 0;
    MLP_CUDA(cudaGetDeviceCount(&ndev));
    MLP_CHECK_ARG(device >= 0 && device < ndev, "mlp_ctx_create: device %d out of range (%d visible)",
                  device, ndev);
    DeviceGuard g(device);
    cudaDeviceProp prop;
    MLP_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        mlp_set_error("mlp_ctx_create: device %d is sm_%d%d; this library is built for sm_100a only",
                      device, prop.major, prop.minor);
        return MLP_EINVAL;
    }
    mlp_ctx* c = new mlp_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->launches = 0;
    for (int i = 0; i < MLP_NUM_ARENAS; ++i) { c->arena[i] = nullptr; c->arena_bytes[i] = 0; }
    c->ctr = nullptr;
    c->frozen = false;
    c->tail_planar = 0;
    c->tail_layout_key = -1;
    c->tail_layout_base = nullptr;
    c->prof_on = false;
    c->prof_used = 0;
    c->prof_ev = nullptr;
    c->prof_stage = nullptr;
    cudaError_t e = cudaMalloc(&c->ctr, MLP_CTR_WORDS * sizeof(int32_t));
    if (e != cudaSuccess) {
        mlp_set_error("mlp_ctx_create: cudaMalloc failed: %s", cudaGetErrorString(e));
        delete c;
        return MLP_ENOMEM;
    }
    cudaMemset(c->ctr, 0, MLP_CTR_WORDS * sizeof(int32_t));
    *out = c;
    return MLP_OK;
}

extern "C" void mlp_ctx_destroy(mlp_ctx* ctx) {
    if (!ctx) return;
    DeviceGuard g(ctx->device);
    for (int i = 0; i < MLP_NUM_ARENAS; ++i)
        if (ctx->arena[i]) cudaFree(ctx->arena[i]);
    if (ctx->ctr) cudaFree(ctx->ctr);
    if (ctx->prof_ev) {
        for (int i = 0; i < 2 * MLP_PROF_CAP; ++i) cudaEventDestroy(ctx->prof_ev[i]);
        delete[] ctx->prof_ev;
        delete[] ctx->prof_stage;
    }
    delete ctx;
}

extern "C" int mlp_ctx_device(const mlp_ctx* ctx) { return ctx ? ctx->device : -1; }
extern "C" int mlp_ctx_sm_count(const mlp_ctx* ctx) { return ctx ? ctx->sm_count : 0; }
extern "C" int64_t mlp_ctx_scratch_bytes(const mlp_ctx* ctx) {
    if (!ctx) return 0;
    int64_t t = 0;
    for (int i = 0; i < MLP_NUM_ARENAS; ++i) t += ctx->arena_bytes[i];
    return t;
}
extern "C" int64_t mlp_ctx_launch_count(const mlp_ctx* ctx) { return ctx ? ctx->launches : 0; }

static const char* kStageNames[MLP_NUM_STAGES] = {
    "threshold_compact", "nms_per_class", "nms_cross_class", "mask_distribute", "roi_plan",
    "roi_align", "trim", "upsample", "paste_threshold", "paste", "elementwise", "mold_batch",
    "tail_fused", "road_scan", "summary", "draw", "resize", "assign", "jpeg", "paste_fill", "", "", "", ""};

extern "C" const char* mlp_stage_name(int stage) {
    return (stage >= 0 && stage < MLP_NUM_STAGES) ? kStageNames[stage] : "";
}

extern "C" int mlp_ctx_profile_enable(mlp_ctx* ctx, int enable) {
    MLP_CHECK_ARG(ctx != nullptr, "mlp_ctx_profile_enable: NULL ctx");
    DeviceGuard g(ctx->device);
    if (enable && !ctx->prof_ev) {
        ctx->prof_ev = new cudaEvent_t[2 * MLP_PROF_CAP];
        ctx->prof_stage = new int[MLP_PROF_CAP];
        for (int i = 0; i < 2 * MLP_PROF_CAP; ++i) MLP_CUDA(cudaEventCreate(&ctx->prof_ev[i]));
    }
    if (enable) ctx->prof_used = 0;
    ctx->prof_on = enable != 0;
    return MLP_OK;
}

extern "C" int mlp_ctx_profile_read(mlp_ctx* ctx, double* ms_out, int64_t* count_out) {
    MLP_CHECK_ARG(ctx && ms_out && count_out, "mlp_ctx_profile_read: NULL argument");
    DeviceGuard g(ctx->device);
    for (int i = 0; i < MLP_NUM_STAGES; ++i) { ms_out[i] = 0.0; count_out[i] = 0; }
    MLP_CUDA(cudaDeviceSynchronize());
    for (int i = 0; i < ctx->prof_used; ++i) {
        float ms = 0.f;
        if (ctx->prof_stage[i] < 0) continue;
        MLP_CUDA(cudaEventElapsedTime(&ms, ctx->prof_ev[2 * i], ctx->prof_ev[2 * i + 1]));
        ms_out[ctx->prof_stage[i]] += ms;
        count_out[ctx->prof_stage[i]] += 1;
    }
    return MLP_OK;
}

extern "C" int mlp_ctx_freeze_scratch(mlp_ctx* ctx, int freeze) {
    MLP_CHECK_ARG(ctx != nullptr, "mlp_ctx_freeze_scratch: NULL ctx");
    ctx->frozen = freeze != 0;
    return MLP_OK;
}

// Grow-only.  Regrowth frees the old arena after a device synchronise, so it must not happen
// while anything still references it: earlier work of this ctx (hence the synchronise) or a
// captured CUDA graph (hence mlp_ctx_freeze_scratch, which PostProcessPipeline.capture sets: a
// frozen ctx answers MLP_EFROZEN instead of freeing memory a graph replay would touch).  Callers
// size the arena from the call's shapes, which are stable after the first call of a given shape.
int mlp_ensure_scratch(mlp_ctx* ctx, int which, int64_t bytes) {
    if (bytes <= ctx->arena_bytes[which]) return MLP_OK;
    if (ctx->frozen && ctx->arena[which]) {      // a first allocation frees nothing and stays allowed
        mlp_set_error("scratch arena %d would have to grow from %lld to %lld bytes, but the ctx is frozen: a "
                      "captured CUDA graph references its arenas.  Use a separate ctx for other shapes, or "
                      "drop the graph and call mlp_ctx_freeze_scratch(ctx, 0).",
                      which, (long long)ctx->arena_bytes[which], (long long)bytes);
        return MLP_EFROZEN;
    }
    DeviceGuard g(ctx->device);
    if (ctx->arena[which]) {
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            // cudaErrorStreamCaptureUnsupported lands here: growth was attempted inside a stream capture
            mlp_set_error("scratch arena %d cannot grow now: cudaDeviceSynchronize failed: %s%s", which,
                          cudaGetErrorString(e),
                          e == cudaErrorStreamCaptureUnsupported
                              ? " (a stream capture is active: run the call once with these shapes before capturing)"
                              : "");
            cudaGetLastError();
            return MLP_ECUDA;
        }
        e = cudaFree(ctx->arena[which]);
        if (e != cudaSuccess) {
            mlp_set_error("scratch cudaFree failed: %s", cudaGetErrorString(e));
            cudaGetLastError();
            return MLP_ECUDA;
        }
        ctx->arena[which] = nullptr;
        ctx->arena_bytes[which] = 0;
    }
    int64_t want = (bytes + (1 << 20) - 1) & ~((int64_t)(1 << 20) - 1);
    cudaError_t e = cudaMalloc(&ctx->arena[which], (size_t)want);
    if (e != cudaSuccess) {
        mlp_set_error("scratch cudaMalloc(%lld bytes) failed: %s", (long long)want,
                      cudaGetErrorString(e));
        cudaGetLastError();
        return e == cudaErrorMemoryAllocation ? MLP_ENOMEM : MLP_ECUDA;
    }
    ctx->arena_bytes[which] = want;
    return MLP_OK;
}

// ------------------------------------------------------------- DLPack --------
// Minimal mirror of dlpack.h's legacy structs (ABI-stable since v0.2).
namespace {
struct DLDeviceM { int32_t device_type; int32_t device_id; };
struct DLDataTypeM { uint8_t code; uint8_t bits; uint16_t lanes; };
struct DLTensorM {
    void* data;
    DLDeviceM device;
    int32_t ndim;
    DLDataTypeM dtype;
    int64_t* shape;
    int64_t* strides;
    uint64_t byte_offset;
};
struct DLManagedTensorM {
    DLTensorM dl_tensor;
    void* manager_ctx;
    void (*deleter)(DLManagedTensorM*);
};
constexpr int kDLCUDA = 2;
}  // namespace

extern "C" int mlp_dlpack_view(const mlp_ctx* ctx, const void* dl_managed_tensor, int expect_dtype,
                               mlp_tensor_view* out) {
    if (!ctx || !dl_managed_tensor || !out) {
        mlp_set_error("mlp_dlpack_view: NULL argument");
        return MLP_EINVAL;
    }
    const DLTensorM& t = static_cast<const DLManagedTensorM*>(dl_managed_tensor)->dl_tensor;
    if (t.device.device_type != kDLCUDA) {
        mlp_set_error("DLPack tensor is not on a CUDA device (device_type=%d); there is no CPU path",
                      t.device.device_type);
        return MLP_EDLPACK;
    }
    if (t.device.device_id != ctx->device) {
        mlp_set_error("DLPack tensor is on cuda:%d but the ctx is bound to cuda:%d",
                      t.device.device_id, ctx->device);
        return MLP_EDLPACK;
    }
    if (t.ndim < 0 || t.ndim > 8) {
        mlp_set_error("DLPack tensor has ndim=%d (supported: 0..8)", t.ndim);
        return MLP_EDLPACK;
    }
    if (t.dtype.lanes != 1) {
        mlp_set_error("DLPack tensor has vector lanes=%d", (int)t.dtype.lanes);
        return MLP_EDLPACK;
    }
    int code = t.dtype.code, bits = t.dtype.bits;
    if (expect_dtype >= 0) {
        int wc = 0, wb = 0;
        switch (expect_dtype) {
            case MLP_F32: wc = 2; wb = 32; break;
            case MLP_I32: wc = 0; wb = 32; break;
            case MLP_U8:  wc = 1; wb = 8;  break;
            case MLP_I64: wc = 0; wb = 64; break;
            default:
                mlp_set_error("mlp_dlpack_view: unknown expect_dtype %d", expect_dtype);
                return MLP_EINVAL;
        }
        if (code != wc || bits != wb) {
            mlp_set_error("DLPack tensor dtype (code=%d,bits=%d) != expected (code=%d,bits=%d)", code,
                          bits, wc, wb);
            return MLP_EDLPACK;
        }
    }
    int64_t numel = 1;
    for (int i = 0; i < t.ndim; ++i) {
        if (t.shape[i] < 0) {
            mlp_set_error("DLPack tensor has negative extent");
            return MLP_EDLPACK;
        }
        numel *= t.shape[i];
    }
    if (t.strides != nullptr && numel > 0) {
        int64_t expect = 1;
        for (int i = t.ndim - 1; i >= 0; --i) {
            if (t.shape[i] != 1 && t.strides[i] != expect) {
                mlp_set_error("DLPack tensor is not dense row-major (dim %d stride %lld, expected %lld)",
                              i, (long long)t.strides[i], (long long)expect);
                return MLP_EDLPACK;
            }
            expect *= t.shape[i];
        }
    }
    char* p = static_cast<char*>(t.data) + t.byte_offset;
    if (numel > 0 && (reinterpret_cast<uintptr_t>(p) & 15u) != 0) {
        mlp_set_error("DLPack tensor data pointer is not 16-byte aligned");
        return MLP_EDLPACK;
    }
    out->data = p;
    out->device_id = t.device.device_id;
    out->ndim = t.ndim;
    out->dtype_code = code;
    out->dtype_bits = bits;
    for (int i = 0; i < 8; ++i) out->shape[i] = (i < t.ndim) ? t.shape[i] : 1;
    out->numel = numel;
    return MLP_OK;
}

// ------------------------------------------------------------- priors --------
static int level_extent(int size, int stride, int same) {
    return same ? (size + stride - 1) / stride : size / stride;
}

int mlp_build_prior_dev(const mlp_prior_config* prior, int height, int width, PriorDev* out) {
    MLP_CHECK_ARG(prior != nullptr, "prior config is NULL");
    MLP_CHECK_ARG(prior->num_levels >= 1 && prior->num_levels <= MLP_MAX_LEVELS,
                  "prior.num_levels=%d out of range [1,%d]", prior->num_levels, MLP_MAX_LEVELS);
    MLP_CHECK_ARG(height > 0 && width > 0, "image size %dx%d must be positive", height, width);
    memset(out, 0, sizeof(*out));
    out->num_levels = prior->num_levels;
    int64_t total = 0;
    for (int l = 0; l < prior->num_levels; ++l) {
        int s = prior->stride[l], A = prior->num_anchors[l];
        MLP_CHECK_ARG(s > 0, "prior.stride[%d]=%d must be positive", l, s);
        MLP_CHECK_ARG(l == 0 || s > prior->stride[l - 1],
                      "prior strides must be strictly ascending (groupby('stride'))");
        MLP_CHECK_ARG(A >= 1 && A <= MLP_MAX_ANCHORS, "prior.num_anchors[%d]=%d out of range [1,%d]", l,
                      A, MLP_MAX_ANCHORS);
        out->stride[l] = s;
        out->na[l] = A;
        out->hf[l] = level_extent(height, s, prior->padding_same);
        out->wf[l] = level_extent(width, s, prior->padding_same);
        out->start[l] = (int)total;
        for (int a = 0; a < A; ++a) {
            int w = prior->anchor_w[l][a], h = prior->anchor_h[l][a];
            MLP_CHECK_ARG(w >= -32768 && w <= 32767 && h >= -32768 && h <= 32767,
                          "anchor size (%d,%d) does not fit int16", w, h);
            out->aw[l][a] = (short)w;
            out->ah[l][a] = (short)h;
        }
        total += (int64_t)out->hf[l] * out->wf[l] * A;
        MLP_CHECK_ARG(total < (1ll << 30), "too many anchors (%lld)", (long long)total);
    }
    for (int l = prior->num_levels; l <= MLP_MAX_LEVELS; ++l) out->start[l] = (int)total;
    out->total = (int)total;
    return MLP_OK;
}

extern "C" int64_t mlp_prior_count(const mlp_prior_config* prior, int height, int width) {
    PriorDev P;
    int rc = mlp_build_prior_dev(prior, height, width, &P);
    if (rc != MLP_OK) return rc;
    return P.total;
}
