// common.cuh — shared host/device helpers for libmasklab_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "masklab_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libmasklab_b200 is written for sm_100a (B200) only"
#endif

// ---------------------------------------------------------------- context ----
#define MLP_NUM_ARENAS 12
struct mlp_ctx {
    int device;
    int sm_count;
    int64_t launches;
    // grow-only scratch arenas, one per stage family so that stages never alias each
    // other's work space (cudaMalloc'd; regrown only between calls)
    void* arena[MLP_NUM_ARENAS];
    int64_t arena_bytes[MLP_NUM_ARENAS];
    // frozen: a captured CUDA graph holds pointers into the arenas, so growth (= free + malloc) is
    // refused with MLP_EFROZEN instead of pulling the memory from under the graph
    bool frozen;
    int tail_planar;       // mask layout of the last mlp_trim_paste (consumed by mlp_tile_summary / mlp_draw_tiles)
    int64_t tail_layout_key;   // shapes / arena the fused tail's scratch was last laid out for (arrival counter zeroed)
    void* tail_layout_base;
    // small device block for counters / dims (always allocated)
    int32_t* ctr;          // [MLP_CTR_WORDS]
    // optional per-stage CUDA-event timing (bench.py roofline): pairs recorded on the
    // launching stream around each stage's kernels
    bool prof_on;
    int prof_used;
    cudaEvent_t* prof_ev;  // [2 * MLP_PROF_CAP]
    int* prof_stage;       // [MLP_PROF_CAP]
};

#define MLP_PROF_CAP 16384
enum {
    MLP_ST_THRESHOLD = 0, MLP_ST_NMS_CLASS, MLP_ST_NMS_CROSS, MLP_ST_DISTRIBUTE, MLP_ST_ROI_PLAN,
    MLP_ST_ROI_ALIGN, MLP_ST_TRIM, MLP_ST_UPSAMPLE, MLP_ST_PASTE_THR, MLP_ST_PASTE, MLP_ST_ELEMENTWISE,
    MLP_ST_MOLD, MLP_ST_TAIL_FUSED, MLP_ST_ROAD_SCAN, MLP_ST_SUMMARY, MLP_ST_DRAW, MLP_ST_RESIZE, MLP_ST_ASSIGN, MLP_ST_JPEG,
    MLP_ST_PASTE_FILL
};

// RAII: records a start event now and a stop event when it goes out of scope.
struct ProfScope {
    mlp_ctx* c; cudaStream_t st; int slot;
    ProfScope(mlp_ctx* ctx, int stage, cudaStream_t stream) : c(ctx), st(stream), slot(-1) {
        if (c->prof_on && c->prof_used < MLP_PROF_CAP) {
            // events recorded inside a stream capture become graph nodes and cannot be read with
            // cudaEventElapsedTime: no brackets while the stream is capturing
            cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
            if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) {
                cudaGetLastError();
                return;
            }
            slot = c->prof_used;
            if (cudaEventRecord(c->prof_ev[2 * slot], st) != cudaSuccess) {
                cudaGetLastError();
                slot = -1;
                return;
            }
            c->prof_stage[slot] = stage;
            c->prof_used++;
        }
    }
    ~ProfScope() {
        if (slot >= 0 && cudaEventRecord(c->prof_ev[2 * slot + 1], st) != cudaSuccess) {
            cudaGetLastError();
            c->prof_stage[slot] = -1;          // unreadable bracket: skipped by mlp_ctx_profile_read
        }
    }
};

#define MLP_CTR_WORDS 1024
#define MLP_ARENA_DETECT 0
#define MLP_ARENA_ROI    1
#define MLP_ARENA_TRIM   2
#define MLP_ARENA_PASTE  3
#define MLP_ARENA_MOLD   4
#define MLP_ARENA_FUSED  5
#define MLP_ARENA_SUMMARY 6
#define MLP_ARENA_BOXACC 7
#define MLP_ARENA_DRAW 8
#define MLP_ARENA_SMOOTH 9
#define MLP_ARENA_ASSIGN 10
#define MLP_ARENA_JPEG 11

void mlp_set_error(const char* fmt, ...);
int mlp_ensure_scratch(mlp_ctx* ctx, int which, int64_t bytes);

#define MLP_CHECK_ARG(cond, ...)                      \
    do {                                              \
        if (!(cond)) {                                \
            mlp_set_error(__VA_ARGS__);               \
            return MLP_EINVAL;                        \
        }                                             \
    } while (0)

#define MLP_CUDA(call)                                                               \
    do {                                                                             \
        cudaError_t e__ = (call);                                                    \
        if (e__ != cudaSuccess) {                                                    \
            mlp_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__),   \
                          __FILE__, __LINE__);                                       \
            return MLP_ECUDA;                                                        \
        }                                                                            \
    } while (0)

#define MLP_LAUNCH_CHECK(ctx)                                                        \
    do {                                                                             \
        (ctx)->launches++;                                                           \
        cudaError_t e__ = cudaGetLastError();                                        \
        if (e__ != cudaSuccess) {                                                    \
            mlp_set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), \
                          __FILE__, __LINE__);                                       \
            return MLP_ECUDA;                                                        \
        }                                                                            \
    } while (0)

static inline bool mlp_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

struct DeviceGuard {
    int prev;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// ---- debug build: bounds checks on scratch / shared-memory indices ------------------------------
// compute-sanitizer is closed on this GPU pool, so memory safety is checked by other means: building with
// -DMLP_DEBUG_BOUNDS (python -m masklab_b200.build --debug-bounds -> libmasklab_b200_dbg.so) turns every
// MLP_BOUND(index, extent) into a test that prints the site and traps; the small-shape GPU tests are then
// run against that library (MASKLAB_B200_LIB=..., tools/run_debug_bounds.sh).  Release builds compile it away.
#ifdef MLP_DEBUG_BOUNDS
#define MLP_BOUND(i, n)                                                                                   \
    do {                                                                                                  \
        if (!((long long)(i) >= 0 && (long long)(i) < (long long)(n))) {                                  \
            printf("MLP_BOUND violated at %s:%d: index %lld not in [0,%lld)\n", __FILE__, __LINE__,       \
                   (long long)(i), (long long)(n));                                                       \
            __trap();                                                                                     \
        }                                                                                                 \
    } while (0)
#else
#define MLP_BOUND(i, n) do { } while (0)
#endif

// PyramidRoiAlign's plan (roi.cu, and the fused tail of the cross-class NMS in detect.cu): the record of
// (level, image, slot) = the source row j of the detection and its box, 32 bytes - what the RoIAlign kernel needs to
// set a RoI up, in one load instead of the chain slot -> j -> row.
struct RoiRec { int32_t j; float box[6]; int32_t pad; };   // box = (cx, cy, w, h, class, conf)
static_assert(sizeof(RoiRec) == 32, "RoiRec layout");

// ------------------------------------------------------- device helpers ------
// Streaming 128-bit global accesses: read-once inputs bypass L1, write-once
// outputs do not allocate in L1 (guideline 13/14 of the Blackwell playbook).
__device__ __forceinline__ float4 ldg_stream_f4(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream_f4(float4* p, const float4& v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x),
                 "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}
__device__ __forceinline__ void stg_stream_u4(uint4* p, const uint4& v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x),
                 "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
}

// ---- TMA 1-D bulk copy global -> shared with mbarrier completion (cp.async.bulk, SASS UBLKCP) ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// Waits for the given phase; traps instead of hanging the GPU if the copy never lands.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    for (int spin = 0; spin < (1 << 22) && !ok; ++spin) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok)
                     : "r"(smem_u32(bar)), "r"(parity)
                     : "memory");
    }
    if (!ok) __trap();
}

// Packed f32x2 arithmetic (Blackwell FADD2): two IEEE-rounded float adds per instruction.  Only
// add/sub are used packed - ptxas contracts a packed multiply feeding a packed add into FFMA2
// even with .rn and -fmad=false, which would change the rounding, so multiplies stay scalar.
__device__ __forceinline__ uint64_t pack2(float a, float b) {
    uint64_t r;
    asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& a, float& b) {
    asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ uint64_t add2_rn(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t sub2_rn(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
// a + (b - a) * l on two lanes, every operation rounded separately (TF / NumPy order)
__device__ __forceinline__ uint64_t lerp2_rn(uint64_t a, uint64_t b, float l) {
    float d0, d1;
    unpack2(sub2_rn(b, a), d0, d1);
    return add2_rn(a, pack2(__fmul_rn(d0, l), __fmul_rn(d1, l)));
}

// Correctly rounded float32 exp / log (evaluate in fp64, round once).  The oracle
// does the same (oracle/masklab_oracle.py exp_f32/log_f32), which makes the
// decoded boxes and FPN levels agree bit for bit between CPU and GPU.
__device__ __forceinline__ float exp_cr(float x) { return (float)exp((double)x); }
__device__ __forceinline__ float log_cr(float x) { return (float)log((double)x); }

// Order-preserving map float -> uint32 (ascending), -0.0 folded onto +0.0.
__device__ __forceinline__ uint32_t float_ordered(float f) {
    f = f + 0.0f;
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float float_from_ordered(uint32_t u) {
    u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
    return __uint_as_float(u);
}

// Anchor lookup table passed by value to kernels (built from mlp_prior_config
// for one image size).
struct PriorDev {
    int num_levels;
    int total;                            // N
    int stride[MLP_MAX_LEVELS];
    int hf[MLP_MAX_LEVELS];
    int wf[MLP_MAX_LEVELS];
    int na[MLP_MAX_LEVELS];
    int start[MLP_MAX_LEVELS + 1];        // first anchor index of each level
    short aw[MLP_MAX_LEVELS][MLP_MAX_ANCHORS];
    short ah[MLP_MAX_LEVELS][MLP_MAX_ANCHORS];
};

int mlp_build_prior_dev(const mlp_prior_config* prior, int height, int width, PriorDev* out);

// anchor n -> (cx, cy, w, h) as ints (engine/layers/detection.py:272-295)
__device__ __forceinline__ int4 prior_anchor(const PriorDev& P, int n) {
    int l = 0;
#pragma unroll
    for (int i = 1; i < MLP_MAX_LEVELS; ++i)
        if (i < P.num_levels && n >= P.start[i]) l = i;
    int r = n - P.start[l];
    int A = P.na[l];
    int cell = r / A;
    int a = r - cell * A;
    int y = cell / P.wf[l];
    int x = cell - y * P.wf[l];
    int s = P.stride[l];
    return make_int4(s / 2 + x * s, s / 2 + y * s, (int)P.aw[l][a], (int)P.ah[l][a]);
}

// RestoreBoxes arithmetic (engine/layers/detection.py:333-341); compiled with
// -fmad=false so multiply and add round separately like TF/NumPy.
__device__ __forceinline__ float4 restore_box(const float4 loc, const int4 pr) {
    float pcx = (float)pr.x, pcy = (float)pr.y, pw = (float)pr.z, ph = (float)pr.w;
    float4 o;
    o.x = __fadd_rn(__fmul_rn(loc.x, pw), pcx);
    o.y = __fadd_rn(__fmul_rn(loc.y, ph), pcy);
    o.z = __fmul_rn(exp_cr(loc.z), pw);
    o.w = __fmul_rn(exp_cr(loc.w), ph);
    return o;
}

// MaskDistribute (engine/layers/instance.py:53-62):
// k = clip(floor(log((sqrt(w*h)+eps)/(base+eps)) / log(2)), 0, max_k); -1 where cx == -1.
__device__ __forceinline__ float level_of(float cx, float w, float h, float base_eps, float max_k) {
    float size = __fsqrt_rn(__fmul_rn(w, h));
    float ratio = __fdiv_rn(__fadd_rn(size, 1e-7f), base_eps);
    float dk = __fdiv_rn(log_cr(ratio), log_cr(2.0f));
    float k = floorf(dk);
    k = fminf(fmaxf(k, 0.0f), max_k);
    return (cx == -1.0f) ? cx : k;
}

// UpSampleOutput box arithmetic (engine/layers/misc.py:179-186): cx,w scale with ratio[0]=PH/hs and
// cy,h with ratio[1]=PW/ws (sic), tf.cast truncates toward zero.
__device__ __forceinline__ void upsample_row(const float* r, float rh, float rw, int32_t* o) {
    o[0] = __float2int_rz(__fmul_rn(r[0], rh));
    o[1] = __float2int_rz(__fmul_rn(r[1], rw));
    o[2] = __float2int_rz(__fmul_rn(r[2], rh));
    o[3] = __float2int_rz(__fmul_rn(r[3], rw));
    o[4] = __float2int_rz(r[4]);
    o[5] = __float2int_rz(__fmul_rn(r[5], 100.0f));
}
