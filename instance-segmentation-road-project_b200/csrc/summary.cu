// summary.cu — the first consumer of the pasted masks: SummaryOutput with
// CrackToInstance, CalculateInstanceSize and IncludeMyRoad.
//
// Reference: /root/reference/engine/layers/misc.py:506-543 (CrackToInstance), :546-591
// (SummaryOutput), :594-625 (IncludeMyRoad), :628-724 (CalculateInstanceSize); wiring
// /root/reference/road_project/setup/serving.py:45-48.  Restated in oracle/summary_oracle.py, whose
// header defines the floating-point contract: every reduction is the sum of the float32 terms
// accumulated in float64 and rounded once; the road-border line fit is the closed-form normal
// equation in float64 on exact integer moments.
//
//   road_scan_kernel      one CTA per image, one pass over the semantic map [PH,PW,S] int32: per-row
//                         road extent (tf.segment_min/max), the 15 % trimmed least-squares fit of
//                         both road borders -> metres per pixel on every frame row (unit[PH]), the
//                         my_road bitmap (1 bit/pixel, stays in L2 for the reduce kernel) and the
//                         batch-wide bounding box of the crack channel.
//   instance_reduce_kernel one CTA per instance: ONE streaming read of its [PH,PW] mask (float32 or
//                         uint8; the crack pseudo-instance reads the semantic channel) produces all
//                         five reductions at once - pixel count, instance / horizontal / vertical
//                         size, my_road overlap - and writes the 11-column summary row.
#include <limits.h>

#include "paste_common.cuh"

namespace {

constexpr int kScanThreads = 512;
constexpr int kMaxFrameRows = 4096;
constexpr int kReduceThreads = 256;

__device__ __forceinline__ long long shfl_xor_ll(long long v, int o) {
    return __shfl_xor_sync(0xffffffffu, v, o);
}
__device__ __forceinline__ double shfl_xor_d(double v, int o) { return __shfl_xor_sync(0xffffffffu, v, o); }

// theta of x = theta0 * y + theta1 through the selected rows (misc.py:706-718); zeros when
// det(X^T X) <= 0.  Moments are exact integers (< 2^53).
__device__ __forceinline__ void fit_line(long long n, long long sy, long long syy, long long sx,
                                         long long sxy, float& t0, float& t1) {
    const double dn = (double)n, dsy = (double)sy, dsyy = (double)syy, dsx = (double)sx, dsxy = (double)sxy;
    const double det = __dsub_rn(__dmul_rn(dsyy, dn), __dmul_rn(dsy, dsy));
    t0 = 0.0f; t1 = 0.0f;
    if (det > 0.0) {
        t0 = (float)__ddiv_rn(__dsub_rn(__dmul_rn(dn, dsxy), __dmul_rn(dsy, dsx)), det);
        t1 = (float)__ddiv_rn(__dsub_rn(__dmul_rn(dsyy, dsx), __dmul_rn(dsy, dsxy)), det);
    }
}

__global__ void __launch_bounds__(kScanThreads)
road_scan_kernel(const int32_t* __restrict__ seg, int PH, int PW, int S, int road_ch, int crack_ch,
                 float road_size, float* __restrict__ unit, uint32_t* __restrict__ road_bits,
                 int32_t* __restrict__ crack_box) {
    __shared__ int s_xmin[kMaxFrameRows], s_xmax[kMaxFrameRows];
    __shared__ int s_scan[kScanThreads / 32];
    __shared__ long long s_red[kScanThreads / 32][7];
    __shared__ float s_theta[4];
    const int b = blockIdx.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int nwarps = kScanThreads / 32;
    const int words = (PW + 31) >> 5;
    const int32_t* img = seg + (int64_t)b * PH * PW * S;
    int cy0 = INT_MAX, cy1 = -1, cx0 = INT_MAX, cx1 = -1;            // crack box of this warp's rows
    for (int y = warp; y < PH; y += nwarps) {
        const int32_t* rowp = img + (int64_t)y * PW * S;
        int xmin = INT_MAX, xmax = -1;
        for (int x0 = 0; x0 < PW; x0 += 32) {
            const int x = x0 + lane;
            const bool in = x < PW;
            const int road = in ? __ldg(rowp + (int64_t)x * S + road_ch) : 0;
            const bool on = road > 0;                                 // tf.where(image > 0), misc.py:661
            const unsigned m = __ballot_sync(0xffffffffu, on);
            if (lane == 0) road_bits[((int64_t)b * PH + y) * words + (x0 >> 5)] = m;
            if (m) {
                xmin = min(xmin, x0 + __ffs(m) - 1);
                xmax = max(xmax, x0 + 31 - __clz(m));
            }
            if (crack_ch >= 0) {
                const int cr = in ? __ldg(rowp + (int64_t)x * S + crack_ch) : 0;
                const unsigned c = __ballot_sync(0xffffffffu, cr != 0);   // tf.where(inputs), misc.py:516
                if (c) {
                    cx0 = min(cx0, x0 + __ffs(c) - 1);
                    cx1 = max(cx1, x0 + 31 - __clz(c));
                    cy0 = min(cy0, y);
                    cy1 = max(cy1, y);
                }
            }
        }
        if (lane == 0) {                                              // tf.segment_min/max: 0 for empty rows
            s_xmin[y] = xmax >= 0 ? xmin : 0;
            s_xmax[y] = xmax >= 0 ? xmax : 0;
        }
    }
    if (crack_ch >= 0 && lane == 0 && cy1 >= 0) {
        atomicMin(crack_box + 0, cy0); atomicMin(crack_box + 1, cx0);
        atomicMax(crack_box + 2, cy1); atomicMax(crack_box + 3, cx1);
    }
    __syncthreads();
    // rank the rows with x_min != x_max (misc.py:689-694): contiguous chunk per thread + block scan
    const int chunk = (PH + kScanThreads - 1) / kScanThreads;
    const int ya = min(tid * chunk, PH), yb = min(ya + chunk, PH);
    int mine = 0;
    for (int y = ya; y < yb; ++y) mine += s_xmin[y] != s_xmax[y];
    int incl = mine;
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) s_scan[warp] = incl;
    __syncthreads();
    int base = 0, total = 0;
    for (int w = 0; w < nwarps; ++w) {
        const int v = s_scan[w];
        if (w < warp) base += v;
        total += v;
    }
    int rank = base + incl - mine;
    // drop 15 % (at least one row) at both ends: marginal[drop:-drop] (misc.py:697-703)
    const int drop = max(1, __float2int_rz(__fmul_rn((float)total, 0.15f)));
    long long acc[7] = {0, 0, 0, 0, 0, 0, 0};           // n, Sy, Syy, Sxl, Sxl*y, Sxr, Sxr*y
    for (int y = ya; y < yb; ++y) {
        const int xl = s_xmin[y], xr = s_xmax[y];
        if (xl == xr) continue;
        if (rank >= drop && rank < total - drop) {
            acc[0] += 1; acc[1] += y; acc[2] += (long long)y * y;
            acc[3] += xl; acc[4] += (long long)xl * y;
            acc[5] += xr; acc[6] += (long long)xr * y;
        }
        ++rank;
    }
#pragma unroll
    for (int q = 0; q < 7; ++q) {
        for (int o = 16; o > 0; o >>= 1) acc[q] += shfl_xor_ll(acc[q], o);
        if (lane == 0) s_red[warp][q] = acc[q];
    }
    __syncthreads();
    if (tid == 0) {
        long long t[7] = {0, 0, 0, 0, 0, 0, 0};
        for (int w = 0; w < nwarps; ++w)
            for (int q = 0; q < 7; ++q) t[q] += s_red[w][q];
        fit_line(t[0], t[1], t[2], t[3], t[4], s_theta[0], s_theta[1]);
        fit_line(t[0], t[1], t[2], t[5], t[6], s_theta[2], s_theta[3]);
    }
    __syncthreads();
    // metres per pixel on every frame row (misc.py:669-678)
    const float l0 = s_theta[0], l1 = s_theta[1], r0 = s_theta[2], r1 = s_theta[3];
    for (int y = tid; y < PH; y += kScanThreads) {
        const float fy = (float)y;
        const float pl = __fadd_rn(__fmul_rn(fy, l0), l1);
        const float pr = __fadd_rn(__fmul_rn(fy, r0), r1);
        const float width = fmaxf(__fsub_rn(pr, pl), 1.0f);           // clip_by_value(.., 1, inf)
        unit[(int64_t)b * PH + y] = __fdiv_rn(road_size, width);
    }
}

// CrackToInstance row (misc.py:521-533) from the batch-wide box; false when the region is empty
// or has zero area (conf = clip(100*h*w, 0, 100) must be > 0, misc.py:562).
__device__ __forceinline__ bool crack_row(const int32_t* crack_box, int32_t* row) {
    int y0 = crack_box[0], x0 = crack_box[1], y1 = crack_box[2], x1 = crack_box[3];
    if (y1 < 0) { y0 = x0 = y1 = x1 = 0; }                            // tf.cond: no pixel -> [[0,0,0]]
    const int h = y1 - y0, w = x1 - x0;
    const long long c = 100ll * h * w;
    row[0] = x0 + w / 2; row[1] = y0 + h / 2; row[2] = w; row[3] = h; row[4] = 5;
    row[5] = (int)(c < 0 ? 0 : (c > 100 ? 100 : c));
    return row[5] > 0;
}

// four adjacent mask pixels x..x+3 of one frame row as float (vector load when the row is aligned)
template <typename T>
__device__ __forceinline__ void load4(const T* __restrict__ rowp, int x, int PW, bool aligned, float* v);
template <>
__device__ __forceinline__ void load4<float>(const float* __restrict__ rowp, int x, int PW, bool aligned, float* v) {
    if (aligned && x + 3 < PW) {
        const float4 f = ldg_stream_f4(reinterpret_cast<const float4*>(rowp + x));
        v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
    } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = (x + q < PW) ? __ldg(rowp + x + q) : 0.0f;
    }
}
template <>
__device__ __forceinline__ void load4<uint8_t>(const uint8_t* __restrict__ rowp, int x, int PW, bool aligned,
                                               float* v) {
    if (aligned && x + 3 < PW) {
        const uchar4 u = __ldg(reinterpret_cast<const uchar4*>(rowp + x));
        v[0] = (float)u.x; v[1] = (float)u.y; v[2] = (float)u.z; v[3] = (float)u.w;
    } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = (x + q < PW) ? (float)__ldg(rowp + x + q) : 0.0f;
    }
}

// Block-wide reduction of the per-thread partial results and the 11-column summary row
// (SummaryOutput.call, misc.py:569-583); called by every thread of the CTA.
__device__ __forceinline__ void finish_row(double pix, double size, double vert, double colmax, int cnt,
                                           int inter, double (*s_d)[4], int (*s_i)[2], const int32_t* row,
                                           float threshold, float* __restrict__ o) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int k = 16; k > 0; k >>= 1) {
        pix = __dadd_rn(pix, shfl_xor_d(pix, k));
        size = __dadd_rn(size, shfl_xor_d(size, k));
        vert = __dadd_rn(vert, shfl_xor_d(vert, k));
        colmax = fmax(colmax, shfl_xor_d(colmax, k));
        cnt += __shfl_xor_sync(0xffffffffu, cnt, k);
        inter += __shfl_xor_sync(0xffffffffu, inter, k);
    }
    __syncthreads();                                       // s_d / s_i free (previous item)
    if (lane == 0) {
        s_d[warp][0] = pix; s_d[warp][1] = size; s_d[warp][2] = vert; s_d[warp][3] = colmax;
        s_i[warp][0] = cnt; s_i[warp][1] = inter;
    }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < kReduceThreads / 32; ++w) {
            pix = __dadd_rn(pix, s_d[w][0]); size = __dadd_rn(size, s_d[w][1]);
            vert = __dadd_rn(vert, s_d[w][2]); colmax = fmax(colmax, s_d[w][3]);
            cnt += s_i[w][0]; inter += s_i[w][1];
        }
        // (class, cx, cy, w, h, conf, pixel_counts, instance, horizontal, vertical, include_my_road)
        o[0] = (float)row[4]; o[1] = (float)row[0]; o[2] = (float)row[1]; o[3] = (float)row[2];
        o[4] = (float)row[3]; o[5] = (float)row[5];
        o[6] = (float)pix; o[7] = (float)size; o[8] = (float)colmax; o[9] = (float)vert;
        const float ioi = __fdiv_rn((float)inter, __fadd_rn((float)cnt, 1e-5f));      // misc.py:616
        o[10] = ioi > threshold ? 1.0f : 0.0f;
    }
}

struct SummaryArgs {
    const int32_t* det;        // [B, m_stride, 6] int32
    const void* masks;         // [B, M, PH, PW] MaskT (dense over the device-side M)
    const int32_t* seg;        // [B, PH, PW, S] int32 (crack pseudo-instance, may be NULL)
    const float* unit;         // [B, PH]
    const uint32_t* road_bits; // [B, PH, words]
    const int32_t* crack_box;  // [4] or NULL: no crack instance is appended
    const int32_t* m_dev;      // [1] or NULL
    int B, m_rows, m_stride, PH, PW, S, crack_ch;
    float threshold;           // IncludeMyRoad threshold
    float* out;                // [B, M', 11]
    int32_t* m_out;            // [1] M'
    int only_crack;            // 1: only the crack pseudo-rows (the others come from tile_summary_kernel)
};

template <typename MaskT>
__global__ void __launch_bounds__(kReduceThreads)
instance_reduce_kernel(const SummaryArgs A) {
    __shared__ float s_unit[kMaxFrameRows];
    __shared__ unsigned s_rowany[kMaxFrameRows / 32];
    __shared__ double s_d[kReduceThreads / 32][4];
    __shared__ int s_i[kReduceThreads / 32][2];
    const int tid = threadIdx.x, lane = tid & 31;
    int M = A.m_dev ? *A.m_dev : A.m_rows;
    if (M > A.m_rows) M = A.m_rows;
    const int m_stride = A.m_stride ? A.m_stride : M;
    int32_t crow[6];
    const bool has_crack = A.crack_box && crack_row(A.crack_box, crow);
    const int Mo = M + (has_crack ? 1 : 0);
    if (blockIdx.x == 0 && tid == 0 && A.m_out) *A.m_out = Mo;
    const int PH = A.PH, PW = A.PW;
    const int words = (PW + 31) >> 5;
    const bool aligned = (PW & 3) == 0;
    const int64_t items = A.only_crack ? (has_crack ? A.B : 0) : (int64_t)A.B * Mo;
    for (int64_t item = blockIdx.x; item < items; item += gridDim.x) {
        const int b = A.only_crack ? (int)item : (int)(item / Mo);
        const int j = A.only_crack ? M : (int)(item - (int64_t)b * Mo);
        const bool crack = j >= M;
        __syncthreads();                                   // previous item done with shared memory
        for (int y = tid; y < PH; y += kReduceThreads) s_unit[y] = A.unit[(int64_t)b * PH + y];
        for (int i = tid; i < (PH + 31) / 32; i += kReduceThreads) s_rowany[i] = 0u;
        __syncthreads();
        const MaskT* mask = static_cast<const MaskT*>(A.masks) + ((int64_t)b * M + (crack ? 0 : j)) * PH * PW;
        const int32_t* cseg = A.seg + (int64_t)b * PH * PW * A.S + A.crack_ch;
        const uint32_t* rbits = A.road_bits + (int64_t)b * PH * words;
        double pix = 0.0, size = 0.0, colmax = 0.0;
        int cnt = 0, inter = 0;
        for (int x0 = 0; x0 < PW; x0 += kReduceThreads * 4) {             // 1024-column tiles
            const int x = x0 + tid * 4;                                   // this thread's 4 columns
            const bool live = x < PW;
            double col[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 4
            for (int y = 0; y < PH; ++y) {
                float v[4] = {0.f, 0.f, 0.f, 0.f};
                if (live) {
                    if (crack) {
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            v[q] = (x + q < PW) ? (float)__ldg(cseg + ((int64_t)y * PW + x + q) * A.S) : 0.0f;
                    } else {
                        load4<MaskT>(mask + (int64_t)y * PW, x, PW, aligned, v);
                    }
                }
                unsigned on = 0u;
                if (v[0] != 0.0f || v[1] != 0.0f || v[2] != 0.0f || v[3] != 0.0f) {   // frames are mostly zeros
                    const float u = s_unit[y];
                    const double du = (double)u, du2 = (double)__fmul_rn(u, u);      // unit ** 2 in float32
                    const unsigned road = (__ldg(rbits + (int64_t)y * words + (x >> 5)) >> (x & 31)) & 0xFu;
                    double rs = 0.0;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const double dv = (double)v[q];
                        rs = __dadd_rn(rs, dv);
                        col[q] = __dadd_rn(col[q], __dmul_rn(du, dv));
                        on |= (v[q] > 0.5f ? 1u : 0u) << q;
                    }
                    pix = __dadd_rn(pix, rs);
                    size = __dadd_rn(size, __dmul_rn(du2, rs));
                    cnt += __popc(on);
                    inter += __popc(on & road);
                }
                const unsigned anyw = __ballot_sync(0xffffffffu, on != 0u);
                if (lane == 0 && anyw) atomicOr(&s_rowany[y >> 5], 1u << (y & 31));
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) colmax = fmax(colmax, col[q]);
        }
        __syncthreads();                                   // row flags complete
        double vert = 0.0;
        for (int y = tid; y < PH; y += kReduceThreads)
            if ((s_rowany[y >> 5] >> (y & 31)) & 1u) vert = __dadd_rn(vert, (double)s_unit[y]);
        finish_row(pix, size, vert, colmax, cnt, inter, s_d, s_i,
                   crack ? crow : A.det + ((int64_t)b * m_stride + j) * 6, A.threshold,
                   A.out + ((int64_t)b * Mo + j) * 11);
    }
}

// ---- the same reductions straight from the mask tiles -------------------------------------
// One CTA per instance: the float32 paste values exist only inside the clipped box, so only the box
// is evaluated (two-stage lerp from the tile in shared memory, the values CropAndPadMask would
// write); a thread owns a box column, walks it top to bottom and carries the column sum, warps
// vote the per-row "any pixel > 0.5" flags.  Nothing of size [PH,PW] is read or written.
struct TileSummaryArgs {
    const int32_t* det;        // [B, m_stride, 6] int32 (UpSampleOutput rows)
    PasteSrc src;              // tile source (standalone int32 tiles or the fused tail)
    const float* unit;         // [B, PH]
    const uint32_t* road_bits; // [B, PH, words]
    const int32_t* crack_box;  // [4] or NULL
    int B, m_rows, m_stride, mh, mw, PH, PW;
    float threshold;
    float* out;                // [B, M', 11], rows j < M
    int32_t* m_out;            // [1] M'
};

__global__ void __launch_bounds__(kReduceThreads)
tile_summary_kernel(const TileSummaryArgs A) {
    __shared__ float s_tile[kMaxTile];
    __shared__ float s_unit[kMaxFrameRows];
    __shared__ unsigned s_rowany[kMaxFrameRows / 32];
    __shared__ double s_d[kReduceThreads / 32][4];
    __shared__ int s_i[kReduceThreads / 32][2];
    const int tid = threadIdx.x, lane = tid & 31;
    int M, thr;
    paste_scalars(A.src, A.B, A.m_rows, M, thr);
    const int m_stride = A.m_stride ? A.m_stride : M;
    int32_t crow[6];
    const bool has_crack = A.crack_box && crack_row(A.crack_box, crow);
    const int Mo = M + (has_crack ? 1 : 0);
    if (blockIdx.x == 0 && tid == 0 && A.m_out) *A.m_out = Mo;
    const int PH = A.PH, PW = A.PW, mh = A.mh, mw = A.mw;
    const int words = (PW + 31) >> 5;
    const int px = mh * mw;
    const int64_t items = (int64_t)A.B * M;
    for (int64_t item = blockIdx.x; item < items; item += gridDim.x) {
        const int b = (int)(item / M), j = (int)(item - (int64_t)b * M);
        const int32_t* row = A.det + ((int64_t)b * m_stride + j) * 6;
        const PasteGeom g = paste_geometry(row, thr, mh, mw, PH, PW);
        double pix = 0.0, size = 0.0, colmax = 0.0, vert = 0.0;
        int cnt = 0, inter = 0;
        __syncthreads();                                   // previous item done with shared memory
        if (g.active) {
            const TileRef tref = tile_ref(A.src, b, j, m_stride, px, row[4], mh, mw);
            for (int i = tid; i < px; i += kReduceThreads) s_tile[i] = (float)tref.at(i);
            for (int y = g.ymin + tid; y < g.ymax; y += kReduceThreads) s_unit[y] = A.unit[(int64_t)b * PH + y];
            for (int i = tid; i < (PH + 31) / 32; i += kReduceThreads) s_rowany[i] = 0u;
            __syncthreads();
            const uint32_t* rbits = A.road_bits + (int64_t)b * PH * words;
            const int bw = g.xmax - g.xmin;
            for (int c0 = 0; c0 < bw; c0 += kReduceThreads) {             // 256 box columns per pass
                const int oxl = c0 + tid;                                 // column inside the box
                const bool live = oxl < bw;
                const int ox = g.xmin + oxl;
                // x terms of the lerp are per column (paste_value, paste_common.cuh)
                const float p = __fmul_rn((float)oxl, g.sx);
                const float fl = floorf(p);
                const int xlo = max((int)fl, 0), xhi = min((int)ceilf(p), mw - 1);
                const float lx = __fsub_rn(p, fl);
                double col = 0.0;
                for (int oy = g.ymin; oy < g.ymax; ++oy) {
                    bool on = false;
                    if (live) {
                        const float py = __fmul_rn((float)(oy - g.ymin), g.sy);
                        const float fy = floorf(py);
                        const int ylo = max((int)fy, 0), yhi = min((int)ceilf(py), mh - 1);
                        const float ly = __fsub_rn(py, fy);
                        const float tl = s_tile[ylo * mw + xlo], tr = s_tile[ylo * mw + xhi];
                        const float bl = s_tile[yhi * mw + xlo], br = s_tile[yhi * mw + xhi];
                        const float t = __fadd_rn(tl, __fmul_rn(__fsub_rn(tr, tl), lx));
                        const float bo = __fadd_rn(bl, __fmul_rn(__fsub_rn(br, bl), lx));
                        const float v = __fadd_rn(t, __fmul_rn(__fsub_rn(bo, t), ly));
                        if (v != 0.0f) {
                            const float u = s_unit[oy];
                            const double dv = (double)v;
                            pix = __dadd_rn(pix, dv);
                            size = __dadd_rn(size, __dmul_rn((double)__fmul_rn(u, u), dv));
                            col = __dadd_rn(col, __dmul_rn((double)u, dv));
                            on = v > 0.5f;
                            if (on) {
                                ++cnt;
                                inter += (__ldg(rbits + (int64_t)oy * words + (ox >> 5)) >> (ox & 31)) & 1u;
                            }
                        }
                    }
                    const unsigned anyw = __ballot_sync(0xffffffffu, on);
                    if (lane == 0 && anyw) atomicOr(&s_rowany[oy >> 5], 1u << (oy & 31));
                }
                colmax = fmax(colmax, col);
            }
            __syncthreads();                               // row flags complete
            for (int y = g.ymin + tid; y < g.ymax; y += kReduceThreads)
                if ((s_rowany[y >> 5] >> (y & 31)) & 1u) vert = __dadd_rn(vert, (double)s_unit[y]);
        }
        finish_row(pix, size, vert, colmax, cnt, inter, s_d, s_i, row, A.threshold,
                   A.out + ((int64_t)b * Mo + j) * 11);
    }
}

}  // namespace

// ================================================================ host side ===
extern "C" int mlp_road_scan(mlp_ctx* ctx, const int32_t* seg_dev, int batch, int frame_h, int frame_w,
                             int channels, int road_channel, int crack_channel, float default_road_size,
                             float* unit_dev, uint32_t* road_bits_dev, int32_t* crack_box_dev,
                             mlp_stream_t stream) {
    MLP_CHECK_ARG(ctx && seg_dev && unit_dev && road_bits_dev, "mlp_road_scan: NULL argument");
    MLP_CHECK_ARG(batch >= 1 && frame_h >= 1 && frame_w >= 1 && channels >= 1, "mlp_road_scan: bad shape");
    MLP_CHECK_ARG(frame_h <= kMaxFrameRows, "mlp_road_scan: frame height %d > %d", frame_h, kMaxFrameRows);
    MLP_CHECK_ARG(road_channel >= 0 && road_channel < channels, "mlp_road_scan: road channel %d out of range",
                  road_channel);
    MLP_CHECK_ARG(crack_channel < channels && (crack_channel < 0 || crack_box_dev),
                  "mlp_road_scan: crack channel %d needs a crack box / is out of range", crack_channel);
    MLP_CHECK_ARG((int64_t)frame_h * frame_w * channels < (1ll << 31), "mlp_road_scan: frame too large");
    DeviceGuard g(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope prof(ctx, MLP_ST_ROAD_SCAN, st);
    if (crack_channel >= 0) {
        const int32_t init[4] = {INT_MAX, INT_MAX, -1, -1};
        MLP_CUDA(cudaMemcpyAsync(crack_box_dev, init, sizeof(init), cudaMemcpyHostToDevice, st));
    }
    road_scan_kernel<<<batch, kScanThreads, 0, st>>>(seg_dev, frame_h, frame_w, channels, road_channel,
                                                   crack_channel, default_road_size, unit_dev,
                                                   road_bits_dev, crack_box_dev);
    MLP_LAUNCH_CHECK(ctx);
    return MLP_OK;
}

extern "C" int mlp_summary_output(mlp_ctx* ctx, const int32_t* det_i32_dev, const void* masks_dev,
                                  int mask_dtype, const int32_t* seg_dev, const float* unit_dev,
                                  const uint32_t* road_bits_dev, const int32_t* crack_box_dev, int batch,
                                  int m_rows, int m_stride, const int32_t* m_dev, int frame_h, int frame_w,
                                  int channels, int crack_channel, float include_threshold,
                                  float* out_dev, int32_t* m_out_dev, mlp_stream_t stream) {
    MLP_CHECK_ARG(ctx && det_i32_dev && masks_dev && unit_dev && road_bits_dev && out_dev,
                  "mlp_summary_output: NULL argument");
    MLP_CHECK_ARG(batch >= 1 && m_rows >= 1 && frame_h >= 1 && frame_w >= 1, "mlp_summary_output: bad shape");
    MLP_CHECK_ARG(frame_h <= kMaxFrameRows, "mlp_summary_output: frame height %d > %d", frame_h, kMaxFrameRows);
    MLP_CHECK_ARG(mask_dtype == MLP_F32 || mask_dtype == MLP_U8, "mlp_summary_output: masks must be f32 or u8");
    MLP_CHECK_ARG(!crack_box_dev || (seg_dev && crack_channel >= 0 && crack_channel < channels),
                  "mlp_summary_output: the crack instance needs the semantic map and its channel");
    MLP_CHECK_ARG(mlp_aligned16(masks_dev), "mlp_summary_output: masks must be 16-byte aligned");
    DeviceGuard g(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope prof(ctx, MLP_ST_SUMMARY, st);
    SummaryArgs A;
    A.det = det_i32_dev; A.masks = masks_dev; A.seg = seg_dev ? seg_dev : det_i32_dev; A.unit = unit_dev;
    A.road_bits = road_bits_dev; A.crack_box = crack_box_dev; A.m_dev = m_dev;
    A.B = batch; A.m_rows = m_rows; A.m_stride = m_stride; A.PH = frame_h; A.PW = frame_w;
    A.S = channels > 0 ? channels : 1; A.crack_ch = crack_channel >= 0 ? crack_channel : 0;
    A.threshold = include_threshold; A.out = out_dev; A.m_out = m_out_dev; A.only_crack = 0;
    const int64_t items = (int64_t)batch * (m_rows + 1);
    const int grid = (int)(items < (1ll << 30) ? items : (1ll << 30));
    if (mask_dtype == MLP_F32) instance_reduce_kernel<float><<<grid, kReduceThreads, 0, st>>>(A);
    else instance_reduce_kernel<uint8_t><<<grid, kReduceThreads, 0, st>>>(A);
    MLP_LAUNCH_CHECK(ctx);
    return MLP_OK;
}

extern "C" int mlp_tile_summary(mlp_ctx* ctx, const int32_t* det_i32_dev, const int32_t* masks_i32_dev,
                                const float* roi_masks_dev, int r_rows, const int32_t* r_dev,
                                int num_classes, const int32_t* counts_dev, int batch, int m_rows,
                                int m_stride, const int32_t* m_dev, int mask_h, int mask_w,
                                const int32_t* seg_dev, const float* unit_dev, const uint32_t* road_bits_dev,
                                const int32_t* crack_box_dev, int frame_h, int frame_w, int channels,
                                int crack_channel, float include_threshold, float* out_dev,
                                int32_t* m_out_dev, int32_t* m_dev_out, mlp_stream_t stream) {
    MLP_CHECK_ARG(ctx && det_i32_dev && unit_dev && road_bits_dev && out_dev, "mlp_tile_summary: NULL argument");
    MLP_CHECK_ARG(masks_i32_dev || (roi_masks_dev && counts_dev && m_dev_out && num_classes >= 1 && r_rows >= 1),
                  "mlp_tile_summary: neither int32 tiles nor a prepared fused tail");
    MLP_CHECK_ARG(batch >= 1 && m_rows >= 1 && frame_h >= 1 && frame_w >= 1 &&
                      (m_stride >= m_rows || (m_stride == 0 && m_dev)),
                  "mlp_tile_summary: bad shape");
    MLP_CHECK_ARG(mask_h >= 1 && mask_w >= 1 && mask_h * mask_w <= kMaxTile, "mlp_tile_summary: mask tile %dx%d",
                  mask_h, mask_w);
    MLP_CHECK_ARG(frame_h <= kMaxFrameRows, "mlp_tile_summary: frame height %d > %d", frame_h, kMaxFrameRows);
    MLP_CHECK_ARG(!crack_box_dev || (seg_dev && crack_channel >= 0 && crack_channel < channels),
                  "mlp_tile_summary: the crack instance needs the semantic map and its channel");
    DeviceGuard g(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope prof(ctx, MLP_ST_SUMMARY, st);
    TileSummaryArgs T;
    memset(&T, 0, sizeof(T));
    const int32_t* m_for_crack = m_dev;
    if (masks_i32_dev) {
        int32_t* thr_dev = ctx->ctr;        // ctr[0]: paste row-filter threshold
        paste_threshold_kernel<<<1, 1024, 0, st>>>(det_i32_dev, batch, m_rows, m_stride, m_dev, thr_dev);
        MLP_LAUNCH_CHECK(ctx);
        T.src.masks_i32 = masks_i32_dev;
        T.src.m_dev = m_dev;
        T.src.thr_dev = thr_dev;
    } else {
        MLP_CHECK_ARG(m_stride == m_rows, "mlp_tile_summary: the fused tail uses capacity rows (m_stride == m_rows)");
        const FusedTail need = fused_tail_layout(nullptr, batch, m_rows, mask_h, mask_w);
        MLP_CHECK_ARG(ctx->arena[MLP_ARENA_FUSED] && ctx->arena_bytes[MLP_ARENA_FUSED] >= need.bytes,
                      "mlp_tile_summary: call mlp_trim_paste with the same shapes first");
        const FusedTail ft = fused_tail_layout(ctx->arena[MLP_ARENA_FUSED], batch, m_rows, mask_h, mask_w);
        T.src.fused = 1;
        T.src.roi_masks = roi_masks_dev;
        T.src.tail_src = ft.tail_src;
        T.src.tail_bits = ft.tail_bits;
        T.src.r_dev = r_dev;
        T.src.r_rows = r_rows;
        T.src.C = num_classes;
        T.src.counts = counts_dev;
        T.src.confmax = ft.confmax;
        T.src.m_out = m_dev_out;
        m_for_crack = m_dev_out;
    }
    T.det = det_i32_dev; T.unit = unit_dev; T.road_bits = road_bits_dev; T.crack_box = crack_box_dev;
    T.B = batch; T.m_rows = m_rows; T.m_stride = m_stride; T.mh = mask_h; T.mw = mask_w;
    T.PH = frame_h; T.PW = frame_w; T.threshold = include_threshold; T.out = out_dev; T.m_out = m_out_dev;
    const int64_t items = (int64_t)batch * m_rows;
    const int grid = (int)(items < (1ll << 30) ? items : (1ll << 30));
    tile_summary_kernel<<<grid, kReduceThreads, 0, st>>>(T);
    MLP_LAUNCH_CHECK(ctx);
    if (crack_box_dev) {
        SummaryArgs A;
        A.det = det_i32_dev; A.masks = det_i32_dev; A.seg = seg_dev; A.unit = unit_dev;
        A.road_bits = road_bits_dev; A.crack_box = crack_box_dev; A.m_dev = m_for_crack;
        A.B = batch; A.m_rows = m_rows; A.m_stride = m_stride; A.PH = frame_h; A.PW = frame_w;
        A.S = channels; A.crack_ch = crack_channel; A.threshold = include_threshold; A.out = out_dev;
        A.m_out = nullptr; A.only_crack = 1;
        instance_reduce_kernel<float><<<batch, kReduceThreads, 0, st>>>(A);
        MLP_LAUNCH_CHECK(ctx);
    }
    return MLP_OK;
}
