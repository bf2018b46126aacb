// roi.cu — PyramidRoiAlign (a10) and TrimInstances (a11).
//
// PyramidRoiAlign (/root/reference/engine/layers/instance.py:109-139): every kept box
// is cropped from the FPN level MaskDistribute assigned it to with TF crop_and_resize
// arithmetic (bilinear, extrapolation 0; restated in oracle/tf_ops.py), then MoldBatch
// regroups crops per level as [B,Mf,ch,cw,Cf] padded with -1.  Here:
//   plan : one CTA per image ranks its boxes per level (ballot prefix scan, order j)
//          -> slot -> source-row table, per-level counts and Mf = max(1, max_b count).
//   run  : one short-lived CTA per RoI (level, image, slot).  Warp 0 builds the RoI's tables
//          (source coordinates, column table, row schedule) and TMA-copies the RoI's source
//          window into shared memory (cp.async.bulk per window row, mbarrier completion);
//          then every warp walks one output column top to bottom with the separable
//          form of the bilinear lerp (see below), lanes own a float4 of channels, results leave
//          as streaming 128-bit stores.  Padded slots are filled with -1 the same way.
// The stage is write-dominated (B*M*196*Cf*4 bytes out vs <= the FPN maps in).
#include <limits.h>
#include <stdlib.h>

#include "common.cuh"

namespace {

#ifndef MLP_ROI_PAR_SCHED
#define MLP_ROI_PAR_SCHED 1              // row schedule of a RoI by 32 lanes (0: one lane, serially)
#endif
#ifndef MLP_ROI_MAXNREG
#define MLP_ROI_MAXNREG 64               // 4 CTAs x 7 warps per SM (56 / 5 CTAs spills since the row schedule went parallel; 64 and 72 measured equal)
#endif
constexpr int kRoiThreads = 224;         // upper bound (7 warps x 72 registers x 4 CTAs fill an SM); the launch uses 3..7 warps
constexpr int kMaxCrop = 64;          // crop_h, crop_w <= 64

struct RoiLevels {
    const float* fmap[MLP_MAX_LEVELS];
    float* crops[MLP_MAX_LEVELS];
    int fh[MLP_MAX_LEVELS];
    int fw[MLP_MAX_LEVELS];
};

// ---- plan ---------------------------------------------------------------------
// roi_rec [L][B][m_rows] : record of (level, image, slot) = the source row j and its box, 32 bytes - what the run
// kernel needs to set a RoI up, in one load instead of the chain slot -> j -> row of dist (RoiRec, common.cuh);
// counts [L][B]; level_m [L].
__global__ void __launch_bounds__(32 * MLP_MAX_LEVELS)
roi_plan_kernel(const float* __restrict__ dist, int B, int m_rows, int m_stride,
                const int32_t* __restrict__ m_dev, int L, RoiRec* __restrict__ roi_rec,
                int32_t* __restrict__ counts, int32_t* __restrict__ level_m) {
    const int b = blockIdx.x;
    const int f = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (f >= L) return;
    int M = m_dev ? *m_dev : m_rows;
    if (M > m_rows) M = m_rows;
    const float* rows = dist + (int64_t)b * m_stride * 7;
    RoiRec* rec = roi_rec + ((int64_t)f * B + b) * m_rows;
    int base = 0;
    for (int j0 = 0; j0 < M; j0 += 32) {
        const int j = j0 + lane;
        const bool hit = (j < M) && (rows[(int64_t)j * 7] == (float)f);
        const unsigned mask = __ballot_sync(0xffffffffu, hit);
        if (hit) {
            RoiRec r;
            r.j = j; r.pad = 0;
#pragma unroll
            for (int q = 0; q < 6; ++q) r.box[q] = rows[(int64_t)j * 7 + 1 + q];
            rec[base + __popc(mask & ((1u << lane) - 1u))] = r;
        }
        base += __popc(mask);
    }
    if (lane == 0) {
        counts[f * B + b] = base;
        atomicMax(level_m + f, base > 1 ? base : 1);
    }
}

// ---- run ----------------------------------------------------------------------
// crop_and_resize is separable: out(y,x) = T + (Bo - T) * ly with T = H(top_y, x), Bo = H(bot_y, x)
// and H(r, x) = img[r][left_x] + (img[r][right_x] - img[r][left_x]) * lx - a value that does not
// depend on y.  Up-sampled RoIs (the common case: 14 samples over fewer source rows) reuse every
// H(r, x) for several output rows, so each warp walks ONE output column top to bottom, keeps the
// last two horizontal lerps in registers (hp, hc) and loads every source row once; bit-identical
// to evaluating four corners per pixel because every H(r, x) is the same three rounded operations.
// Per RoI warp 0 builds the schedule once: the list of source rows to load, and after each row
// the outputs (row offset, ly) that become computable as lerp(hp, hc, ly).
constexpr int kMaxRows = 2 * kMaxCrop;       // worst case two new source rows per output row
constexpr int kRoiWindowBytes = 48 * 1024;   // FPN window staged in shared memory per RoI (4 CTAs per SM)

struct RoiOut { uint32_t yoff; float ly; };  // byte offset of the output row inside the crop, y weight
struct RoiSched {                            // per-RoI tables, built once per CTA by warp 0
    float iny[kMaxCrop], inx[kMaxCrop];          // crop_and_resize source coordinates
    uint32_t xl[kMaxCrop], xr[kMaxCrop];         // byte offsets of the left/right source pixel in a row
    float lx[kMaxCrop];                          // x lerp weight; < 0: column outside -> 0
    uint32_t rowoff[kMaxRows];                   // byte offsets of the source rows, in load order
    unsigned short obeg[kMaxRows + 1];           // outputs [obeg[k], obeg[k+1]) follow the load of row k
    RoiOut out[kMaxCrop];
    uint32_t zoff[kMaxCrop];                     // output rows outside the map -> 0
    int nrows, nzero, staged;
};

__device__ __forceinline__ float4 lerp_v(const float4& a, const float4& b, float l) {
    float4 r;
    unpack2(lerp2_rn(pack2(a.x, a.y), pack2(b.x, b.y), l), r.x, r.y);
    unpack2(lerp2_rn(pack2(a.z, a.w), pack2(b.z, b.w), l), r.z, r.w);
    return r;
}
__device__ __forceinline__ float lerp_v(float a, float b, float l) {
    return __fadd_rn(a, __fmul_rn(__fsub_rn(b, a), l));
}
__device__ __forceinline__ void zero_v(float4& v) { v = make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ void zero_v(float& v) { v = 0.f; }
__device__ __forceinline__ void store_v(unsigned char* p, const float4& v) { stg_stream_f4(reinterpret_cast<float4*>(p), v); }
__device__ __forceinline__ void store_v(unsigned char* p, float v) { *reinterpret_cast<float*>(p) = v; }
template <bool kStaged> __device__ __forceinline__ void load_v(const unsigned char* p, float4& v) {
    if (kStaged) v = *reinterpret_cast<const float4*>(p);
    else v = __ldg(reinterpret_cast<const float4*>(p));
}
template <bool kStaged> __device__ __forceinline__ void load_v(const unsigned char* p, float& v) {
    if (kStaged) v = *reinterpret_cast<const float*>(p);
    else v = __ldg(reinterpret_cast<const float*>(p));
}

// One source row (its left and right pixel) in registers.
template <typename Vec> struct RowRegs { Vec l, r; };

template <typename Vec, bool kStaged>
__device__ __forceinline__ void row_load(RowRegs<Vec>& q, const RoiSched& S, const unsigned char* pl,
                                         const unsigned char* pr, int k, int K) {
    if (k < K) {
        const uint32_t ro = S.rowoff[k];
        load_v<kStaged>(pl + ro, q.l);
        load_v<kStaged>(pr + ro, q.r);
    }
}
template <typename Vec>
__device__ __forceinline__ void row_consume(const RowRegs<Vec>& q, const RoiSched& S, unsigned char* o,
                                            float lx, int k, int K, Vec& hp, Vec& hc) {
    if (k < K) {
        hp = hc;
        hc = lerp_v(q.l, q.r, lx);
        const int j1 = S.obeg[k + 1];
#pragma unroll 1
        for (int j = S.obeg[k]; j < j1; ++j) {
            const RoiOut e = S.out[j];
            store_v(o + e.yoff, lerp_v(hp, hc, e.ly));
        }
    }
}

// One RoI: warp-task = (output column x, group of 32 channel vectors); Vec = float4 (Cf % 4 == 0)
// or float.  kStaged: source rows come from the TMA-staged window in shared memory.  The row loads
// are double-buffered: the loads of row k+1 are in flight while row k is used.
template <typename Vec, bool kStaged>
__device__ __forceinline__ void roi_columns(const unsigned char* __restrict__ src, const RoiSched& S,
                                            float* __restrict__ out, int ch, int cw, int Cf, int warp,
                                            int lane, int nwarps) {
    constexpr int V = sizeof(Vec) / 4;
    const int CV = Cf / V;                                 // channel vectors per pixel
    const int groups = (CV + 31) >> 5;
    const int K = S.nrows, nz = S.nzero;
    const uint32_t ystride = (uint32_t)(cw * Cf) * 4u;
#pragma unroll 1
    for (int task = warp; task < cw * groups; task += nwarps) {
        int g = 0, x = task;
        if (groups > 1) { g = task / cw; x = task - g * cw; }      // Cf <= 128: one group, no division
        const int c = (g << 5) + lane;
        if (c >= CV) continue;
        unsigned char* o = reinterpret_cast<unsigned char*>(out + (size_t)x * Cf + (size_t)c * V);
        const float lx = S.lx[x];
        Vec z;
        zero_v(z);
        if (!(lx >= 0.0f)) {                                // column outside: extrapolation_value = 0
            for (int y = 0; y < ch; ++y) store_v(o + y * ystride, z);
            continue;
        }
        for (int i = 0; i < nz; ++i) store_v(o + S.zoff[i], z);
        const unsigned char* pl = src + S.xl[x] + c * (V * 4);
        const unsigned char* pr = src + S.xr[x] + c * (V * 4);
        Vec hp = z, hc = z;                                 // H of the previous / the current source row
        RowRegs<Vec> qa, qb;
        row_load<Vec, kStaged>(qa, S, pl, pr, 0, K);
#pragma unroll 1
        for (int k0 = 0; k0 < K; k0 += 2) {
            row_load<Vec, kStaged>(qb, S, pl, pr, k0 + 1, K);
            row_consume<Vec>(qa, S, o, lx, k0, K, hp, hc);
            row_load<Vec, kStaged>(qa, S, pl, pr, k0 + 2, K);
            row_consume<Vec>(qb, S, o, lx, k0 + 1, K, hp, hc);
        }
    }
}

__global__ void __maxnreg__(MLP_ROI_MAXNREG)
roi_align_kernel(const RoiLevels lv, int L, int Cf, const float* __restrict__ dist, int B,
                 int m_rows, int m_stride, float image_h, float image_w, int ch, int cw,
                 const RoiRec* __restrict__ roi_rec, const int32_t* __restrict__ counts,
                 int32_t* __restrict__ level_m, float* __restrict__ roi_boxes, int window_cap) {
    extern __shared__ __align__(128) unsigned char s_window[];   // TMA-staged FPN window of this RoI
    __shared__ RoiSched S;
    __shared__ __align__(8) uint64_t s_bar;   // mbarrier the TMA row copies complete on
    const int nthreads = blockDim.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nwarps = nthreads >> 5;
    // every thread reads the device-side level sizes itself (one round trip, no barrier): CTAs
    // beyond the B * R real items exit at once
    int R = 0;
#pragma unroll 1
    for (int f = 0; f < L; ++f) R += level_m[f];
    if (blockIdx.x == 0 && threadIdx.x == 0) level_m[L] = R;   // R = sum of Mf, for TrimInstances
    const unsigned items = (unsigned)B * (unsigned)R;          // <= L * B * m_rows < 2^31 (host check)
    if (blockIdx.x >= items) return;
    if (threadIdx.x == 0) mbar_init(&s_bar, 1);
    const int npix = ch * cw;
    const bool vec = (Cf & 3) == 0;
    uint32_t phase = 0;                        // parity of the next mbarrier phase to wait for

    // item = (image b, row r of roi_boxes) -> (level f, slot): the output order of the reference
    for (unsigned item = blockIdx.x; item < items; item += gridDim.x) {
        const int b = (int)(item / (unsigned)R), r = (int)(item - (unsigned)b * (unsigned)R);
        int f = 0, foff = 0, mf = 0;
        {
            int o = 0;
#pragma unroll 1
            for (int q = 0; q < L; ++q) {
                const int m = level_m[q];                    // L1/L2 hits after the first read
                if (r >= o) { f = q; foff = o; mf = m; }
                o += m;
            }
        }
        const int slot = r - foff;
        const int cnt = counts[f * B + b];
        // the slot's record travels beside the count (warp 0, eight lanes: one 32-byte sector); what it holds for a
        // padded slot is never used
        float recv = 0.0f;
        if (warp == 0 && lane < 8)
            recv = __ldg(reinterpret_cast<const float*>(roi_rec + ((int64_t)f * B + b) * m_rows + slot) + lane);
        float* out = lv.crops[f] + ((int64_t)b * mf + slot) * npix * Cf;
        float* rb = roi_boxes + (int64_t)item * 6;

        if (slot >= cnt) {                                   // MoldBatch padding
            if (threadIdx.x < 6) rb[threadIdx.x] = -1.0f;
            const int n = npix * Cf;
            if (vec) {
                float4* o4 = reinterpret_cast<float4*>(out);
                const float4 m1 = make_float4(-1.f, -1.f, -1.f, -1.f);
                for (int i = threadIdx.x; i < (n >> 2); i += nthreads) stg_stream_f4(o4 + i, m1);
            } else {
                for (int i = threadIdx.x; i < n; i += nthreads) out[i] = -1.0f;
            }
            continue;
        }
        const int Hf = lv.fh[f], Wf = lv.fw[f];
        const float* img = lv.fmap[f] + (int64_t)b * Hf * Wf * Cf;
        __syncthreads();                                     // previous RoI done with tables + window
        if (warp == 0) {
            // ---- the whole per-RoI setup in one warp, one barrier for everybody else
            MLP_BOUND(__float_as_int(__shfl_sync(0xffffffffu, recv, 0)), m_rows);
            const float bcx = __shfl_sync(0xffffffffu, recv, 1), bcy = __shfl_sync(0xffffffffu, recv, 2);
            const float bw = __shfl_sync(0xffffffffu, recv, 3), bh = __shfl_sync(0xffffffffu, recv, 4);
            const float hm1 = (float)(Hf - 1), wm1 = (float)(Wf - 1);
            if (lane >= 1 && lane < 7) rb[lane - 1] = recv;
            int ylo = INT_MAX, yhi = -1, xlo = INT_MAX, xhi = -1;
            for (int i = lane; i < ch + cw; i += 32) {
                // NormalizeBoxes(shape=image) then crop_and_resize source coordinates
                const bool is_y = i < ch;
                const int idx = is_y ? i : i - ch;
                const float c = is_y ? bcy : bcx;
                const float s = is_y ? bh : bw;
                const float dim = is_y ? image_h : image_w;
                const float half = __fdiv_rn(s, 2.0f);
                const float lo = __fdiv_rn(__fsub_rn(c, half), dim);      // y1 : x1
                const float hi = __fdiv_rn(__fadd_rn(c, half), dim);      // y2 : x2
                const int nout = is_y ? ch : cw;
                const float fm1 = is_y ? hm1 : wm1;
                float in;
                if (nout > 1) {
                    const float scale = __fdiv_rn(__fmul_rn(__fsub_rn(hi, lo), fm1), (float)(nout - 1));
                    in = __fadd_rn(__fmul_rn(lo, fm1), __fmul_rn((float)idx, scale));
                } else {
                    in = __fmul_rn(__fmul_rn(0.5f, __fadd_rn(lo, hi)), fm1);
                }
                // TF: in < 0 || in > size-1 -> extrapolation value; NaN counts as outside
                const bool ok = in >= 0.0f && in <= fm1;
                if (is_y) {
                    S.iny[idx] = in;
                    if (ok) { ylo = min(ylo, (int)floorf(in)); yhi = max(yhi, (int)ceilf(in)); }
                } else {
                    S.inx[idx] = in;
                    if (ok) { xlo = min(xlo, (int)floorf(in)); xhi = max(xhi, (int)ceilf(in)); }
                }
            }
            for (int o = 16; o > 0; o >>= 1) {
                ylo = min(ylo, __shfl_xor_sync(0xffffffffu, ylo, o));
                yhi = max(yhi, __shfl_xor_sync(0xffffffffu, yhi, o));
                xlo = min(xlo, __shfl_xor_sync(0xffffffffu, xlo, o));
                xhi = max(xhi, __shfl_xor_sync(0xffffffffu, xhi, o));
            }
            // source window of the valid samples; if it fits, TMA it into shared memory (one
            // cp.async.bulk per window row - rows are contiguous in NHWC), else rows come from global
            const bool any = yhi >= 0 && xhi >= 0;
            const int rows = any ? yhi - ylo + 1 : 0, cols = any ? xhi - xlo + 1 : 0;
            const int64_t row_bytes = (int64_t)cols * Cf * 4;
            const bool staged = any && vec && row_bytes * rows <= window_cap;
            if (staged) {
                if (lane == 0) mbar_expect_tx(&s_bar, (uint32_t)(row_bytes * rows));
                __syncwarp();
                // generic-proxy reads of the previous RoI's window (ordered by the barrier above)
                // before the async proxy overwrites it
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                MLP_BOUND(row_bytes * rows - 1, window_cap);
                for (int q = lane; q < rows; q += 32)
                    tma_load_1d(s_window + (size_t)q * row_bytes,
                                img + ((int64_t)(ylo + q) * Wf + xlo) * Cf, (uint32_t)row_bytes, &s_bar);
            }
            const int top0 = staged ? ylo : 0, left0 = staged ? xlo : 0, pitch = staged ? cols : Wf;
            __syncwarp();
            for (int i = lane; i < cw; i += 32) {            // column table
                const float in_x = S.inx[i];
                uint32_t xl = 0, xr = 0;
                float lx = -1.0f;
                if (in_x >= 0.0f && in_x <= wm1) {
                    const float fx = floorf(in_x);
                    lx = __fsub_rn(in_x, fx);
                    xl = (uint32_t)(((int)fx - left0) * Cf) * 4u;
                    xr = (uint32_t)(((int)ceilf(in_x) - left0) * Cf) * 4u;
                }
                S.xl[i] = xl; S.xr[i] = xr; S.lx[i] = lx;
            }
            const uint32_t pitch_bytes = (uint32_t)(pitch * Cf) * 4u;
            const uint32_t ystride = (uint32_t)(cw * Cf) * 4u;
            if (MLP_ROI_PAR_SCHED && ch <= 32) {
                // row schedule, one output row per lane: the source rows (t, bo) a row needs against those the previous
                // valid row left in (hp, hc) decide how many new rows it pushes; positions by ballot prefix counts
                const float v = lane < ch ? S.iny[lane] : -1.0f;
                const bool inr = lane < ch;
                const bool valid = inr && v >= 0.0f && v <= hm1;
                const float fy = floorf(v);
                const int t = (int)fy, bo = (int)ceilf(v);
                const unsigned lt = (1u << lane) - 1u;
                const unsigned vmask = __ballot_sync(0xffffffffu, valid);
                const unsigned before = vmask & lt;
                const int src = before ? 31 - __clz(before) : 0;      // the previous valid row, wherever it is
                const int pt = __shfl_sync(0xffffffffu, t, src), pbo = __shfl_sync(0xffffffffu, bo, src);
                int npush = 0;
                if (valid) {
                    if (before && pt == t && pbo == bo) npush = 0;
                    else if (before && pbo == t) npush = 1;
                    else npush = 2;
                }
                const unsigned m1 = __ballot_sync(0xffffffffu, npush >= 1), m2 = __ballot_sync(0xffffffffu, npush == 2);
                const unsigned zmask = __ballot_sync(0xffffffffu, inr && !valid);
                const int kb = __popc(m1 & lt) + __popc(m2 & lt), no = __popc(before);
                MLP_BOUND(kb + npush, kMaxRows + 1);
                if (npush == 2) {
                    S.obeg[kb] = (unsigned short)no;
                    S.rowoff[kb] = (uint32_t)(t - top0) * pitch_bytes;
                    S.obeg[kb + 1] = (unsigned short)no;
                    S.rowoff[kb + 1] = (uint32_t)(bo - top0) * pitch_bytes;
                } else if (npush == 1) {
                    S.obeg[kb] = (unsigned short)no;
                    S.rowoff[kb] = (uint32_t)(bo - top0) * pitch_bytes;
                }
                if (valid) {
                    S.out[no].yoff = (uint32_t)lane * ystride;
                    S.out[no].ly = __fsub_rn(v, fy);
                } else if (inr) {
                    S.zoff[__popc(zmask & lt)] = (uint32_t)lane * ystride;
                }
                if (lane == 0) {
                    const int K = __popc(m1) + __popc(m2);
                    S.obeg[K] = (unsigned short)__popc(vmask);
                    S.nrows = K; S.nzero = __popc(zmask); S.staged = staged;
                }
            } else if (lane == 0) {                          // crops taller than a warp: the same schedule, serially
                int K = 0, no = 0, nz = 0;
                int prev = INT_MIN, cur = INT_MIN;           // source rows held in hp / hc
                for (int y = 0; y < ch; ++y) {
                    const float v = S.iny[y];
                    if (!(v >= 0.0f && v <= hm1)) { S.zoff[nz++] = (uint32_t)y * ystride; continue; }
                    const float fy = floorf(v);
                    const int t = (int)fy, bo = (int)ceilf(v);
                    // make hp = H(t), hc = H(bo) with as few new rows as possible
                    int npush = 0, first = bo;
                    if (prev == t && cur == bo) npush = 0;
                    else if (cur == t) npush = 1;
                    else { npush = 2; first = t; }
                    MLP_BOUND(K + npush, kMaxRows + 1);
                    MLP_BOUND(no, kMaxCrop);
                    if (npush == 2) {
                        S.obeg[K] = (unsigned short)no;
                        S.rowoff[K++] = (uint32_t)(first - top0) * pitch_bytes;
                    }
                    if (npush >= 1) {
                        S.obeg[K] = (unsigned short)no;
                        S.rowoff[K++] = (uint32_t)(bo - top0) * pitch_bytes;
                    }
                    prev = t; cur = bo;
                    S.out[no].yoff = (uint32_t)y * ystride;
                    S.out[no].ly = __fsub_rn(v, fy);
                    ++no;
                }
                S.obeg[K] = (unsigned short)no;
                S.nrows = K; S.nzero = nz; S.staged = staged;
            }
        }
        __syncthreads();
        if (S.staged) {
            mbar_wait(&s_bar, phase);
            phase ^= 1u;
            roi_columns<float4, true>(s_window, S, out, ch, cw, Cf, warp, lane, nwarps);
        } else if (vec) {
            roi_columns<float4, false>(reinterpret_cast<const unsigned char*>(img), S, out, ch, cw, Cf, warp, lane,
                                       nwarps);
        } else {
            roi_columns<float, false>(reinterpret_cast<const unsigned char*>(img), S, out, ch, cw, Cf, warp, lane,
                                      nwarps);
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
    int rc = mlp_ensure_scratch(ctx, MLP_ARENA_ROI, (int64_t)num_levels * batch * m_rows * (int64_t)sizeof(RoiRec));
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope prof(ctx, MLP_ST_ROI_PLAN, st);
    MLP_CUDA(cudaMemsetAsync(level_m_dev, 0, (size_t)(num_levels + 1) * 4, st));
    roi_plan_kernel<<<batch, 32 * MLP_MAX_LEVELS, 0, st>>>(
        dist_dev, batch, m_rows, m_stride, m_dev, num_levels,
        static_cast<RoiRec*>(ctx->arena[MLP_ARENA_ROI]), level_counts_dev, level_m_dev);
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
                      crop_h + crop_w <= 2 * kMaxCrop,
                  "mlp_roi_align_run: crop size %dx%d out of range [1,%d]", crop_h, crop_w, kMaxCrop);
    MLP_CHECK_ARG(ctx->arena[MLP_ARENA_ROI] &&
                      ctx->arena_bytes[MLP_ARENA_ROI] >= (int64_t)num_levels * batch * m_rows * (int64_t)sizeof(RoiRec),
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
    // The real item count B * R (R = sum over levels of max_b count, <= L * m_rows) is only known on the device.
    // Every image's counts add up to <= m_rows, so R is typically a little above m_rows: the grid is sized
    // for B * (m_rows + m_rows / 4 + L) items and the kernel's grid-stride loop covers the (rare) rest; a
    // grid for the worst case would launch ~6,000 CTAs at cfg-2 that only read R and exit.
    const int64_t worst = (int64_t)num_levels * batch * m_rows;
    int64_t items = (int64_t)batch * (m_rows + m_rows / 4 + num_levels);
    if (items > worst) items = worst;
    if (const char* e = getenv("MLP_ROI_FULL_GRID")) { if (atoi(e)) items = worst; }
    const int grid = (int)(items < (1ll << 30) ? items : (1ll << 30));
    int window_cap = kRoiWindowBytes;
    if (const char* e = getenv("MLP_ROI_WINDOW_KB")) window_cap = atoi(e) * 1024;    // tuning knob; 0 = never stage
    if (window_cap < 0) window_cap = 0;
    if (window_cap > 160 * 1024) window_cap = 160 * 1024;
    // warp-tasks per RoI = crop_w columns x groups of 32 channel vectors; pick the warp count
    // (<= 8) that leaves no idle warp in the last round
    const int cv = (channels & 3) ? channels : channels / 4;
    const int tasks = crop_w * ((cv + 31) / 32);
    int nwarps = 7;
    double best = 2.0;
    for (int w = 7; w >= 4; --w) {
        const int rounds = (tasks + w - 1) / w;
        const double idle = 1.0 - (double)tasks / (rounds * w);
        if (idle < best - 1e-9) { best = idle; nwarps = w; }
    }
    if (const char* e = getenv("MLP_ROI_WARPS")) nwarps = atoi(e);
    if (nwarps < 1) nwarps = 1;
    if (nwarps > kRoiThreads / 32) nwarps = kRoiThreads / 32;
    const size_t smem = (size_t)window_cap;
    MLP_CUDA(cudaFuncSetAttribute(roi_align_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    roi_align_kernel<<<grid, nwarps * 32, smem, (cudaStream_t)stream>>>(
        lv, num_levels, channels, dist_dev, batch, m_rows, m_stride, image_h, image_w, crop_h, crop_w,
        static_cast<const RoiRec*>(ctx->arena[MLP_ARENA_ROI]), level_counts_dev,
        level_m_dev, roi_boxes_dev, window_cap);
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
