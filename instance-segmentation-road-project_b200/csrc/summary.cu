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

// Reductions of the crack pseudo-instance of one image (bitmap mask), produced by mlp_road_scan and
// stored behind the crack box: crack_box_dev i32 [4 + 8 * B] = box, then one CrackPart per image.
struct CrackPart {
    double size, vert;                   // sum_y unit^2[y] * count[y],  sum_y unit[y] * [count[y] > 0]
    unsigned long long colmax_bits;      // bit pattern of max_x sum_y unit[y] * bit[y,x]  (>= 0)
    int pix, inter;                      // crack pixels, crack pixels on my_road
};
static_assert(sizeof(CrackPart) == 32, "CrackPart is 8 int32 words");

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

// Pass 1 over the semantic map, one warp per frame row: road extent of the row (tf.segment_min /
// segment_max: 0 for empty rows), the my_road and crack bitmaps, the batch-wide crack box.
constexpr int kRowsThreads = 256;
constexpr int kRowsUnroll = 4;            // 32-pixel groups in flight per warp

__global__ void __launch_bounds__(kRowsThreads)
road_rows_kernel(const int32_t* __restrict__ seg, int PH, int PW, int S, int road_ch, int crack_ch,
                 int2* __restrict__ row_ext, uint32_t* __restrict__ road_bits,
                 uint32_t* __restrict__ crack_bits, int2* __restrict__ crack_rows,
                 int32_t* __restrict__ crack_box) {
    const int b = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int nwarps = kRowsThreads / 32;
    const int y = blockIdx.x * nwarps + warp;
    if (y >= PH) return;
    const int words = (PW + 31) >> 5;
    const int32_t* rowp = seg + ((int64_t)b * PH + y) * PW * S;
    uint32_t* rb = road_bits + ((int64_t)b * PH + y) * words;
    uint32_t* cb = crack_bits ? crack_bits + ((int64_t)b * PH + y) * words : nullptr;
    int xmin = INT_MAX, xmax = -1, cx0 = INT_MAX, cx1 = -1;
    int cpix = 0, cinter = 0;                               // crack pixels of the row, and those on my_road
    for (int w0 = 0; w0 < words; w0 += kRowsUnroll) {
        int road[kRowsUnroll], cr[kRowsUnroll];
#pragma unroll
        for (int u = 0; u < kRowsUnroll; ++u) {
            const int x = (w0 + u) * 32 + lane;
            const bool in = x < PW;
            road[u] = in ? __ldg(rowp + (int64_t)x * S + road_ch) : 0;
            cr[u] = (in && cb) ? __ldg(rowp + (int64_t)x * S + crack_ch) : 0;
        }
#pragma unroll
        for (int u = 0; u < kRowsUnroll; ++u) {
            const int x0 = (w0 + u) * 32;
            const unsigned m = __ballot_sync(0xffffffffu, road[u] > 0);     // tf.where(image > 0), misc.py:661
            const unsigned c = __ballot_sync(0xffffffffu, cr[u] != 0);      // tf.where(inputs), misc.py:516
            if (w0 + u < words) {
                if (lane == 0) {
                    rb[w0 + u] = m;
                    if (cb) cb[w0 + u] = c;
                }
                if (m) { xmin = min(xmin, x0 + __ffs(m) - 1); xmax = max(xmax, x0 + 31 - __clz(m)); }
                if (c) {
                    cx0 = min(cx0, x0 + __ffs(c) - 1); cx1 = max(cx1, x0 + 31 - __clz(c));
                    cpix += __popc(c); cinter += __popc(c & m);
                }
            }
        }
    }
    if (lane == 0) {
        row_ext[(int64_t)b * PH + y] = xmax >= 0 ? make_int2(xmin, xmax) : make_int2(0, 0);
        if (crack_rows) crack_rows[(int64_t)b * PH + y] = make_int2(cpix, cinter);
        if (cx1 >= 0) {
            atomicMin(crack_box + 0, y); atomicMin(crack_box + 1, cx0);
            atomicMax(crack_box + 2, y); atomicMax(crack_box + 3, cx1);
        }
    }
}

// The same pass for the serving graph's three-channel map (channels == 3, frame width a multiple of 4, 16-byte aligned
// rows): every lane takes four pixels = three 128-bit loads, a warp 128 pixels = 1.5 KB of contiguous row per step, two
// steps in flight - every byte of the map crosses once, in full 128-byte lines.  The four bits of a lane meet the
// other seven lanes of its 32-pixel word through three xor-shuffles.
__device__ __forceinline__ uint4 ldg_stream_u4(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ unsigned or_over_8(unsigned v) {
    v |= __shfl_xor_sync(0xffffffffu, v, 1);
    v |= __shfl_xor_sync(0xffffffffu, v, 2);
    v |= __shfl_xor_sync(0xffffffffu, v, 4);
    return v;
}

template <int kRoad, int kCrack>
__global__ void __launch_bounds__(kRowsThreads)
road_rows3_kernel(const int32_t* __restrict__ seg, int PH, int PW, int2* __restrict__ row_ext,
                  uint32_t* __restrict__ road_bits, uint32_t* __restrict__ crack_bits, int2* __restrict__ crack_rows,
                  int32_t* __restrict__ crack_box) {
    const int b = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int nwarps = kRowsThreads / 32;
    const int y = blockIdx.x * nwarps + warp;
    if (y >= PH) return;
    const int words = (PW + 31) >> 5;
    const uint4* rowv = reinterpret_cast<const uint4*>(seg + ((int64_t)b * PH + y) * PW * 3);
    uint32_t* rb = road_bits + ((int64_t)b * PH + y) * words;
    uint32_t* cb = crack_bits ? crack_bits + ((int64_t)b * PH + y) * words : nullptr;
    const int grp = lane >> 3, sub = lane & 7;
    int xmin = INT_MAX, xmax = -1, cx0 = INT_MAX, cx1 = -1;
    int cpix = 0, cinter = 0;                               // crack pixels of the row, and those on my_road
    constexpr int kSteps = 2;
    for (int base = 0; base < PW; base += 128 * kSteps) {
        uint4 v[kSteps][3];
#pragma unroll
        for (int u = 0; u < kSteps; ++u) {
            const int x = base + u * 128 + 4 * lane;
#pragma unroll
            for (int k = 0; k < 3; ++k)
                v[u][k] = x < PW ? ldg_stream_u4(rowv + (x >> 2) * 3 + k) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < kSteps; ++u) {
            const int32_t e[12] = {(int32_t)v[u][0].x, (int32_t)v[u][0].y, (int32_t)v[u][0].z, (int32_t)v[u][0].w,
                                   (int32_t)v[u][1].x, (int32_t)v[u][1].y, (int32_t)v[u][1].z, (int32_t)v[u][1].w,
                                   (int32_t)v[u][2].x, (int32_t)v[u][2].y, (int32_t)v[u][2].z, (int32_t)v[u][2].w};
            unsigned mn = 0, cn = 0;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                mn |= (e[3 * q + kRoad] > 0 ? 1u : 0u) << q;             // tf.where(image > 0), misc.py:661
                cn |= (e[3 * q + kCrack] != 0 ? 1u : 0u) << q;           // tf.where(inputs), misc.py:516
            }
            const unsigned m = or_over_8(mn << (4 * sub));
            const unsigned c = cb ? or_over_8(cn << (4 * sub)) : 0u;
            const int wi = ((base + u * 128) >> 5) + grp, x0 = wi * 32;
            if (sub == 0 && x0 < PW) {
                rb[wi] = m;
                if (cb) cb[wi] = c;
                if (m) { xmin = min(xmin, x0 + __ffs(m) - 1); xmax = max(xmax, x0 + 31 - __clz(m)); }
                if (c) {
                    cx0 = min(cx0, x0 + __ffs(c) - 1); cx1 = max(cx1, x0 + 31 - __clz(c));
                    cpix += __popc(c); cinter += __popc(c & m);
                }
            }
        }
    }
    xmin = __reduce_min_sync(0xffffffffu, xmin); xmax = __reduce_max_sync(0xffffffffu, xmax);
    cx0 = __reduce_min_sync(0xffffffffu, cx0); cx1 = __reduce_max_sync(0xffffffffu, cx1);
    cpix = __reduce_add_sync(0xffffffffu, cpix); cinter = __reduce_add_sync(0xffffffffu, cinter);
    if (lane == 0) {
        row_ext[(int64_t)b * PH + y] = xmax >= 0 ? make_int2(xmin, xmax) : make_int2(0, 0);
        if (crack_rows) crack_rows[(int64_t)b * PH + y] = make_int2(cpix, cinter);
        if (cx1 >= 0) {
            atomicMin(crack_box + 0, y); atomicMin(crack_box + 1, cx0);
            atomicMax(crack_box + 2, y); atomicMax(crack_box + 3, cx1);
        }
    }
}

// Pass 2, one CTA per image: the 15 % trimmed least-squares fit of both road borders over the rows
// with x_min != x_max, then metres per pixel on every frame row.
__global__ void __launch_bounds__(kScanThreads)
road_fit_kernel(const int2* __restrict__ row_ext, int PH, float road_size, float* __restrict__ unit,
                const int2* __restrict__ crack_rows, CrackPart* __restrict__ crack_part) {
    __shared__ int s_xmin[kMaxFrameRows], s_xmax[kMaxFrameRows];
    __shared__ int s_scan[kScanThreads / 32];
    __shared__ long long s_red[kScanThreads / 32][7];
    __shared__ float s_theta[4];
    const int b = blockIdx.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int nwarps = kScanThreads / 32;
    for (int y = tid; y < PH; y += kScanThreads) {
        const int2 e = row_ext[(int64_t)b * PH + y];
        s_xmin[y] = e.x; s_xmax[y] = e.y;
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
    // the crack pseudo-instance is a bitmap: everything but its horizontal size follows from the
    // per-row pixel counts of pass 1 (pixel count, instance size, vertical size, my_road overlap)
    double size = 0.0, vert = 0.0;
    long long pix = 0, inter = 0;
    for (int y = tid; y < PH; y += kScanThreads) {
        const float fy = (float)y;
        const float pl = __fadd_rn(__fmul_rn(fy, l0), l1);
        const float pr = __fadd_rn(__fmul_rn(fy, r0), r1);
        const float width = fmaxf(__fsub_rn(pr, pl), 1.0f);           // clip_by_value(.., 1, inf)
        const float u = __fdiv_rn(road_size, width);
        unit[(int64_t)b * PH + y] = u;
        if (crack_rows) {
            const int2 c = crack_rows[(int64_t)b * PH + y];
            if (c.x > 0) {
                pix += c.x; inter += c.y;
                size = __dadd_rn(size, __dmul_rn((double)__fmul_rn(u, u), (double)c.x));
                vert = __dadd_rn(vert, (double)u);
            }
        }
    }
    if (crack_rows == nullptr) return;
    __syncthreads();                                        // s_red free
    for (int o = 16; o > 0; o >>= 1) {
        size = __dadd_rn(size, shfl_xor_d(size, o));
        vert = __dadd_rn(vert, shfl_xor_d(vert, o));
        pix += shfl_xor_ll(pix, o);
        inter += shfl_xor_ll(inter, o);
    }
    if (lane == 0) {
        s_red[warp][0] = __double_as_longlong(size); s_red[warp][1] = __double_as_longlong(vert);
        s_red[warp][2] = pix; s_red[warp][3] = inter;
    }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < nwarps; ++w) {
            size = __dadd_rn(size, __longlong_as_double(s_red[w][0]));
            vert = __dadd_rn(vert, __longlong_as_double(s_red[w][1]));
            pix += s_red[w][2]; inter += s_red[w][3];
        }
        CrackPart& P = crack_part[b];
        P.size = size; P.vert = vert; P.colmax_bits = 0ull; P.pix = (int)pix; P.inter = (int)inter;
    }
}

// Pass 3 (crack only): horizontal size of the crack pseudo-instance = max over columns of
// sum_y unit[y] * bit[y,x].  One CTA per 32-column word of the batch-wide crack box: its eight warps take the box
// rows in interleaved runs of eight (eight word loads in flight, zero words skipped), lane = column; the warps'
// column sums meet in shared memory, the maximum goes to the image's CrackPart with an atomicMax on the bit
// pattern (sums are >= 0, so the order is preserved).
constexpr int kColsThreads = 256;

__global__ void __launch_bounds__(kColsThreads)
crack_cols_kernel(const uint32_t* __restrict__ crack_bits, const float* __restrict__ unit,
                  const int32_t* __restrict__ crack_box, int PH, int words, CrackPart* __restrict__ crack_part) {
    __shared__ double s_col[kColsThreads / 32][32];
    const int y0 = crack_box[0], x0 = crack_box[1], y1 = crack_box[2], x1 = crack_box[3];
    if (y1 < 0) return;                                     // no crack pixel in the batch
    const int b = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int nwarps = kColsThreads / 32;
    const int wi = (x0 >> 5) + blockIdx.x;
    if (wi > (x1 >> 5)) return;
    const uint32_t* cb = crack_bits + (int64_t)b * PH * words + wi;
    const float* un = unit + (int64_t)b * PH;
    double col = 0.0;
    for (int oy = y0 + 8 * warp; oy <= y1; oy += 8 * nwarps) {
        uint32_t cw[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) cw[u] = (oy + u <= y1) ? __ldg(cb + (int64_t)(oy + u) * words) : 0u;
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if ((cw[u] >> lane) & 1u) col = __dadd_rn(col, (double)__ldg(un + oy + u));
    }
    s_col[warp][lane] = col;
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int w = 1; w < nwarps; ++w) col = __dadd_rn(col, s_col[w][lane]);
        for (int o = 16; o > 0; o >>= 1) col = fmax(col, shfl_xor_d(col, o));
        if (lane == 0 && col > 0.0)
            atomicMax(&crack_part[b].colmax_bits, (unsigned long long)__double_as_longlong(col));
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

// Block-wide reduction of the per-thread partial results; thread 0 returns with the totals.
// Called by every thread of the CTA.
__device__ __forceinline__ void block_totals(double& pix, double& size, double& vert, double& colmax, int& cnt,
                                             int& inter, double (*s_d)[4], int (*s_i)[2]) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int k = 16; k > 0; k >>= 1) {
        pix = __dadd_rn(pix, shfl_xor_d(pix, k));
        size = __dadd_rn(size, shfl_xor_d(size, k));
        vert = __dadd_rn(vert, shfl_xor_d(vert, k));
        colmax = fmax(colmax, shfl_xor_d(colmax, k));
        cnt += __shfl_xor_sync(0xffffffffu, cnt, k);
        inter += __shfl_xor_sync(0xffffffffu, inter, k);
    }
    __syncthreads();                                       // s_d / s_i free (previous use)
    if (lane == 0) {
        s_d[warp][0] = pix; s_d[warp][1] = size; s_d[warp][2] = vert; s_d[warp][3] = colmax;
        s_i[warp][0] = cnt; s_i[warp][1] = inter;
    }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
            pix = __dadd_rn(pix, s_d[w][0]); size = __dadd_rn(size, s_d[w][1]);
            vert = __dadd_rn(vert, s_d[w][2]); colmax = fmax(colmax, s_d[w][3]);
            cnt += s_i[w][0]; inter += s_i[w][1];
        }
    }
}

// The 11-column summary row (SummaryOutput.call, misc.py:569-583):
// (class, cx, cy, w, h, conf, pixel_counts, instance, horizontal, vertical, include_my_road)
__device__ __forceinline__ void write_row(double pix, double size, double vert, double colmax, int cnt, int inter,
                                          const int32_t* row, float threshold, float* __restrict__ o) {
    o[0] = (float)row[4]; o[1] = (float)row[0]; o[2] = (float)row[1]; o[3] = (float)row[2];
    o[4] = (float)row[3]; o[5] = (float)row[5];
    o[6] = (float)pix; o[7] = (float)size; o[8] = (float)colmax; o[9] = (float)vert;
    const float ioi = __fdiv_rn((float)inter, __fadd_rn((float)cnt, 1e-5f));      // misc.py:616
    o[10] = ioi > threshold ? 1.0f : 0.0f;
}

__device__ __forceinline__ void finish_row(double pix, double size, double vert, double colmax, int cnt,
                                           int inter, double (*s_d)[4], int (*s_i)[2], const int32_t* row,
                                           float threshold, float* __restrict__ o) {
    block_totals(pix, size, vert, colmax, cnt, inter, s_d, s_i);
    if (threadIdx.x == 0) write_row(pix, size, vert, colmax, cnt, inter, row, threshold, o);
}

struct SummaryArgs {
    const int32_t* det;        // [B, m_stride, 6] int32
    const void* masks;         // [B, M, PH, PW] MaskT (dense over the device-side M)
    const float* unit;         // [B, PH]
    const uint32_t* road_bits; // [B, PH, words]
    const int32_t* crack_box;  // [4] or NULL: no crack instance is appended
    const int32_t* m_dev;      // [1] or NULL
    int B, m_rows, m_stride, PH, PW;
    float threshold;           // IncludeMyRoad threshold
    float* out;                // [B, M', 11]
    int32_t* m_out;            // [1] M'
};

constexpr int kRowsInFlight = 8;          // 16-byte row loads in flight per thread

// One CTA per instance, one thread per 16-byte column group (4 float32 / 16 uint8 pixels), walking
// all frame rows with kRowsInFlight independent loads in flight: a pure streaming read whose
// arithmetic only runs on the (few) non-zero groups.  blockDim = min(256, groups per row).
template <typename MaskT>
__global__ void __launch_bounds__(kReduceThreads)
instance_reduce_kernel(const SummaryArgs A) {
    using Vec = MaskVec<MaskT>;
    constexpr int kPx = Vec::kPx;
    __shared__ unsigned s_rowany[kMaxFrameRows / 32];
    __shared__ double s_d[kReduceThreads / 32][4];
    __shared__ int s_i[kReduceThreads / 32][2];
    const int tid = threadIdx.x, nthreads = blockDim.x;
    int M = A.m_dev ? *A.m_dev : A.m_rows;
    if (M > A.m_rows) M = A.m_rows;
    const int m_stride = A.m_stride ? A.m_stride : M;
    int32_t crow[6];
    const bool has_crack = A.crack_box && crack_row(A.crack_box, crow);
    const int Mo = M + (has_crack ? 1 : 0);
    if (blockIdx.x == 0 && tid == 0 && A.m_out) *A.m_out = Mo;
    const int PH = A.PH, PW = A.PW;
    const int words = (PW + 31) >> 5;
    const bool aligned = (PW % kPx) == 0;                   // every row starts 16-byte aligned
    const int groups = (PW + kPx - 1) / kPx;
    const int64_t items = (int64_t)A.B * M;                 // the crack rows come from box_summary_kernel
    for (int64_t item = blockIdx.x; item < items; item += gridDim.x) {
        const int b = (int)(item / M), j = (int)(item - (int64_t)b * M);
        __syncthreads();                                   // previous item done with shared memory
        for (int i = tid; i < (PH + 31) / 32; i += nthreads) s_rowany[i] = 0u;
        __syncthreads();
        const MaskT* mask = static_cast<const MaskT*>(A.masks) + ((int64_t)b * M + j) * PH * PW;
        const uint32_t* rbits = A.road_bits + (int64_t)b * PH * words;
        const float* unit = A.unit + (int64_t)b * PH;
        double pix = 0.0, size = 0.0, colmax = 0.0;
        int cnt = 0, inter = 0;
        for (int g = tid; g < groups; g += nthreads) {
            const int x = g * kPx;
            const int n = min(kPx, PW - x);
            const bool vec = aligned && n == kPx;
            double col[kPx];
#pragma unroll
            for (int q = 0; q < kPx; ++q) col[q] = 0.0;
            for (int y0 = 0; y0 < PH; y0 += kRowsInFlight) {
                Vec v[kRowsInFlight];
#pragma unroll
                for (int u = 0; u < kRowsInFlight; ++u) {
                    const int y = min(y0 + u, PH - 1);
                    if (vec) v[u].load(mask + (int64_t)y * PW + x);
                    else v[u].load_tail(mask + (int64_t)y * PW + x, n);
                }
#pragma unroll
                for (int u = 0; u < kRowsInFlight; ++u) {
                    const int y = y0 + u;
                    if (y >= PH || !v[u].any()) continue;                 // frames are mostly zeros
                    const float un = __ldg(unit + y);
                    const double du = (double)un;
                    double rs = 0.0;
                    unsigned on = 0u;
#pragma unroll
                    for (int q = 0; q < kPx; ++q) {
                        const float f = v[u].at(q);
                        const double dv = (double)f;
                        rs = __dadd_rn(rs, dv);
                        col[q] = __dadd_rn(col[q], __dmul_rn(du, dv));
                        on |= (f > 0.5f ? 1u : 0u) << q;
                    }
                    pix = __dadd_rn(pix, rs);
                    size = __dadd_rn(size, __dmul_rn((double)__fmul_rn(un, un), rs));   // unit ** 2 in float32
                    if (on) {
                        atomicOr(&s_rowany[y >> 5], 1u << (y & 31));
                        cnt += __popc(on);
                        // my_road bits of pixels x .. x+kPx-1 (kPx divides 32: inside one word)
                        const unsigned road = (__ldg(rbits + (int64_t)y * words + (x >> 5)) >> (x & 31)) &
                                              (kPx == 16 ? 0xffffu : 0xfu);
                        inter += __popc(on & road);
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < kPx; ++q) colmax = fmax(colmax, col[q]);
        }
        __syncthreads();                                   // row flags complete
        double vert = 0.0;
        for (int y = tid; y < PH; y += nthreads)
            if ((s_rowany[y >> 5] >> (y & 31)) & 1u) vert = __dadd_rn(vert, (double)__ldg(unit + y));
        finish_row(pix, size, vert, colmax, cnt, inter, s_d, s_i,
                   A.det + ((int64_t)b * m_stride + j) * 6, A.threshold, A.out + ((int64_t)b * Mo + j) * 11);
    }
}

// ---- the same reductions inside a box, without any [PH,PW] tensor ---------------------------
// One CTA per work item = (instance, chunk of 128 columns of its clipped box).  An instance's float32 paste values
// exist only inside the box and are evaluated there from its tile (two-stage lerp, the values CropAndPadMask would
// write).  Lane layout: a chunk of width <= 32 / <= 64 / <= 128 columns is walked with 8 / 16 / 32 lanes per row, so a
// warp covers 4 / 2 / 1 box rows per step and every lane owns FOUR columns (c, c+cw, c+2cw, c+3cw) whose x terms it
// computes once (median box of the benchmark: 47 columns - with a fixed 128-column layout 63 % of the lane slots were
// idle).  The 8 warps take the row groups round-robin; lanes carry their column sums, which meet across the row
// groups of a warp by shuffles and across the warps in shared memory for the horizontal maximum.  Row groups vote
// their "any pixel > 0.5" flag.  Boxes wider than one chunk are split over several CTAs whose partial results meet
// in a per-instance accumulator; the last CTA to arrive writes the row.  The crack pseudo-instance of every image
// was reduced by mlp_road_scan and is only copied here.
// Tiles that arrive as bit rows (the fused tail, mw <= 32) are {0,1} by construction: the x lerp of a row,
// tl + (tr - tl) * lx, is then one of 0, lx, 1 - lx (rounded once, as fadd(1, -lx)) or 1 - selected by the two
// corner bits with the same float32 results as the arithmetic, without the four shared-memory loads per pixel.
constexpr int kBoxWarps = kReduceThreads / 32;
constexpr int kBoxCols = 128;             // columns per chunk = 4 per lane x 32 lanes at most
constexpr int kBoxQ = 4;                  // columns per lane
constexpr int kBitRows = 64;              // tile rows the bit path holds

struct BoxAcc {                           // per-instance accumulator of a box split over several CTAs
    double pix, size;
    unsigned long long colmax_bits;
    int cnt, inter, done, pad;
};
static_assert(sizeof(BoxAcc) == 40, "BoxAcc layout");

struct BoxItem {                          // one work item: instance, chunk and the instance's clipped box (32 bytes)
    uint32_t code;                        // (b * m_rows + j) | chunk << 20
    int xmin, xmax, ymin, ymax;
    float sx, sy;
    int active;
};
static_assert(sizeof(BoxItem) == 32, "BoxItem layout");

struct BoxSummaryArgs {
    const int32_t* det;        // [B, m_stride, 6] int32 (UpSampleOutput rows)
    PasteSrc src;              // tile source (standalone int32 tiles or the fused tail)
    int has_tiles;             // 0: only the crack pseudo-instances (rows j < M come from elsewhere)
    const float* unit;         // [B, PH]
    const uint32_t* road_bits; // [B, PH, words]
    const int32_t* crack_box;  // [4 + 8 * B] box + CrackPart per image (mlp_road_scan), or NULL
    const int32_t* m_dev;      // M when there are no tiles
    BoxAcc* acc;               // [B * m_rows], zeroed before the launch
    uint32_t* acc_rowany;      // [B * m_rows, ceil(PH / 32)], zeroed before the launch
    int B, m_rows, m_stride, mh, mw, PH, PW, chunks;      // chunks = ceil(PW / 128): most per instance
    float threshold;
    float* out;                // [B, M', 11]
    int32_t* m_out;            // [1] M'
    BoxItem* items;            // work items of box_plan_kernel
    int32_t* n_items;          // [1]
    int item_cap;
};

// One CTA per image, one thread per instance row (rounds of 256): which 128-column chunks of its clipped box exist.
// Only those become work items (an atomicAdd per round and CTA), widest chunk index first within an instance; the
// summary kernel walks the list instead of testing all B * M * ceil(PW / 128) combinations, 7 of 8 of which are empty.
// The box geometry (two float divisions, four ceils) is computed here once per instance and travels in the item.
__global__ void __launch_bounds__(kReduceThreads)
box_plan_kernel(const BoxSummaryArgs A) {
    __shared__ int s_cnt[kReduceThreads / 32];
    __shared__ int s_base;
    int M, thr = 0;
    paste_scalars(A.src, A.B, A.m_rows, M, thr);
    const int m_stride = A.m_stride ? A.m_stride : M;
    const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int j0 = 0; j0 < M; j0 += kReduceThreads) {
        const int j = j0 + tid;
        int n = 0;
        BoxItem it;
        it.active = 0;
        if (j < M) {
            const PasteGeom g = paste_geometry(A.det + ((int64_t)b * m_stride + j) * 6, thr, A.mh, A.mw, A.PH, A.PW);
            n = g.active ? (g.xmax - g.xmin + kBoxCols - 1) / kBoxCols : 1;      // an all-zero mask still gets its row
            it.xmin = g.xmin; it.xmax = g.xmax; it.ymin = g.ymin; it.ymax = g.ymax; it.sx = g.sx; it.sy = g.sy;
            it.active = g.active ? 1 : 0;
        }
        int incl = n;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) s_cnt[warp] = incl;
        __syncthreads();
        int pre = 0, tot = 0;
#pragma unroll
        for (int q = 0; q < kReduceThreads / 32; ++q) {
            if (q < warp) pre += s_cnt[q];
            tot += s_cnt[q];
        }
        if (tid == 0) s_base = atomicAdd(A.n_items, tot);
        __syncthreads();
        int at = s_base + pre + incl - n;
        for (int c = n - 1; c >= 0; --c, ++at)
            if (at < A.item_cap) {
                it.code = (uint32_t)(b * A.m_rows + j) | ((uint32_t)c << 20);
                A.items[at] = it;
            }
        __syncthreads();
    }
}

struct BoxSmem {
    float tile[kMaxTile];                 // general path: the tile as floats
    uint32_t rows[kBitRows];              // bit path: the tile's rows
    float unit[kMaxFrameRows];
    unsigned rowany[kMaxFrameRows / 32];
    double col[kBoxWarps][kBoxCols];
    double d[kBoxWarps][4];
    int i[kBoxWarps][2];
    int last;
};

// One chunk (columns [c0, c0 + 4 * cw) of the box, cw = 1 << cwl lanes per row); every thread of the CTA calls it.
template <bool kBits>
__device__ __forceinline__ void box_reduce(BoxSmem& S, const PasteGeom& g, int c0, int cwl, int mh, int mw,
                                           const uint32_t* __restrict__ rbits, int words, double& pix,
                                           double& size, double& colmax, int& cnt, int& inter) {
    constexpr int Q = kBoxQ;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int cw = 1 << cwl, c = lane & (cw - 1), r = lane >> cwl, rpw = 32 >> cwl;
    const int bw = g.xmax - g.xmin;
    // this lane's Q columns of the chunk and their x lerp terms (paste_value, paste_common.cuh)
    int xlo[Q], xhi[Q];                   // kBits: single-bit masks of the two source columns (0 outside the box)
    float lx[Q];
    bool live[Q];
    double col[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        const int oxl = c0 + c + cw * q;
        live[q] = oxl < bw;
        const float p = __fmul_rn((float)oxl, g.sx);
        const float fl = floorf(p);
        xlo[q] = max((int)fl, 0);
        xhi[q] = min((int)ceilf(p), mw - 1);
        lx[q] = __fsub_rn(p, fl);
        col[q] = 0.0;
        if (kBits) {
            xlo[q] = live[q] ? (int)(1u << xlo[q]) : 0;
            xhi[q] = live[q] ? (int)(1u << xhi[q]) : 0;
        }
    }
    const int xw0 = g.xmin + c0 + c;                        // frame column of q = 0; q adds cw
    const unsigned gmask = (cw == 32 ? 0xffffffffu : ((1u << cw) - 1u)) << (r * cw);   // lanes of this row group
    for (int oyb = g.ymin + warp * rpw; oyb < g.ymax; oyb += kBoxWarps * rpw) {        // warp-uniform
        const int oy = oyb + r;
        float v[Q];
        bool nz = false;
#pragma unroll
        for (int q = 0; q < Q; ++q) v[q] = 0.0f;
        if (oy < g.ymax) {
            const float py = __fmul_rn((float)(oy - g.ymin), g.sy);
            const float fy = floorf(py);
            const int ylo = max((int)fy, 0), yhi = min((int)ceilf(py), mh - 1);
            const float ly = __fsub_rn(py, fy);
            if (kBits) {
                const uint32_t w0 = S.rows[ylo], w1 = S.rows[yhi];
#pragma unroll
                for (int q = 0; q < Q; ++q) {
                    const bool tl = (w0 & (uint32_t)xlo[q]) != 0u, tr = (w0 & (uint32_t)xhi[q]) != 0u;
                    const bool bl = (w1 & (uint32_t)xlo[q]) != 0u, br = (w1 & (uint32_t)xhi[q]) != 0u;
                    const float oml = __fsub_rn(1.0f, lx[q]);                  // fadd(1, (0 - 1) * lx)
                    const float t = tl ? (tr ? 1.0f : oml) : (tr ? lx[q] : 0.0f);
                    const float bo = bl ? (br ? 1.0f : oml) : (br ? lx[q] : 0.0f);
                    v[q] = __fadd_rn(t, __fmul_rn(__fsub_rn(bo, t), ly));
                    nz |= v[q] != 0.0f;
                }
            } else {
                const float* r0 = S.tile + ylo * mw;
                const float* r1 = S.tile + yhi * mw;
#pragma unroll
                for (int q = 0; q < Q; ++q) {
                    if (live[q]) {
                        const float tl = r0[xlo[q]], tr = r0[xhi[q]], bl = r1[xlo[q]], br = r1[xhi[q]];
                        const float t = __fadd_rn(tl, __fmul_rn(__fsub_rn(tr, tl), lx[q]));
                        const float bo = __fadd_rn(bl, __fmul_rn(__fsub_rn(br, bl), lx[q]));
                        v[q] = __fadd_rn(t, __fmul_rn(__fsub_rn(bo, t), ly));
                    }
                    nz |= v[q] != 0.0f;
                }
            }
        }
        unsigned on = 0u;
        if (nz) {
            const float u = S.unit[oy];
            const double du = (double)u;
            double d[Q];
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                d[q] = (double)v[q];
                col[q] = __fma_rn(du, d[q], col[q]);       // float32 x float32 is exact in float64: same bits as mul + add
                on |= (v[q] > 0.5f ? 1u : 0u) << q;
            }
            const double rs = __dadd_rn(__dadd_rn(d[0], d[1]), __dadd_rn(d[2], d[3]));
            pix = __dadd_rn(pix, rs);
            size = __dadd_rn(size, __dmul_rn((double)__fmul_rn(u, u), rs));          // unit ** 2 in float32
        }
        const unsigned vote = __ballot_sync(0xffffffffu, on != 0u);
        if (vote) {                                          // warp-uniform: some row of the step has a pixel > 0.5
            if ((vote & gmask) && c == 0) atomicOr(&S.rowany[oy >> 5], 1u << (oy & 31));
            if (on) {
                cnt += __popc(on);
                // my_road bits of the lane's lit columns (a word-wise variant - ballots of the four column runs, one
                // funnel-shifted AND per run - was measured slower: 135 against 108 us)
                const uint32_t* rw = rbits + (int64_t)oy * words;
#pragma unroll
                for (int q = 0; q < Q; ++q)
                    if ((on >> q) & 1u) {
                        const int xw = xw0 + cw * q;
                        inter += (int)((__ldg(rw + (xw >> 5)) >> (xw & 31)) & 1u);
                    }
            }
        }
    }
    // horizontal size: the row groups of a warp first (shuffles), then the 8 warps in shared memory
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        for (int o = cw; o < 32; o <<= 1) col[q] = __dadd_rn(col[q], shfl_xor_d(col[q], o));
        if (r == 0) S.col[warp][q * cw + c] = col[q];
    }
    __syncthreads();
    if (tid < Q * cw) {
        double t = S.col[0][tid];
#pragma unroll
        for (int w = 1; w < kBoxWarps; ++w) t = __dadd_rn(t, S.col[w][tid]);
        colmax = fmax(colmax, t);
    }
}

__global__ void __launch_bounds__(kReduceThreads, 4)
box_summary_kernel(const BoxSummaryArgs A) {
    __shared__ BoxSmem S;
    const int tid = threadIdx.x;
    int M, thr = 0;
    if (A.has_tiles) {
        paste_scalars(A.src, A.B, A.m_rows, M, thr);
    } else {
        M = A.m_dev ? *A.m_dev : A.m_rows;
        if (M > A.m_rows) M = A.m_rows;
    }
    const int m_stride = A.m_stride ? A.m_stride : M;
    int32_t crow[6];
    const bool has_crack = A.crack_box && crack_row(A.crack_box, crow);
    const int Mo = M + (has_crack ? 1 : 0);
    if (blockIdx.x == 0 && tid == 0 && A.m_out) *A.m_out = Mo;
    const int PH = A.PH, PW = A.PW, mh = A.mh, mw = A.mw;
    const int words = (PW + 31) >> 5, ywords = (PH + 31) >> 5;
    const int n_crack = has_crack ? A.B : 0;
    const int n_list = A.has_tiles ? min(*A.n_items, A.item_cap) : 0;
    const int items = n_crack + n_list;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
        if (item < n_crack) {                               // reduced by mlp_road_scan already
            if (tid == 0) {
                const int b = (int)item;
                const CrackPart P = reinterpret_cast<const CrackPart*>(A.crack_box + 4)[b];
                write_row((double)P.pix, P.size, P.vert, __longlong_as_double((long long)P.colmax_bits), P.pix,
                          P.inter, crow, A.threshold, A.out + ((int64_t)b * Mo + M) * 11);
            }
            continue;
        }
        // the item: two 16-byte loads, the same address for the whole CTA
        const uint4 i0 = __ldg(reinterpret_cast<const uint4*>(A.items + (item - n_crack)));
        const uint4 i1 = __ldg(reinterpret_cast<const uint4*>(A.items + (item - n_crack)) + 1);
        PasteGeom g;
        g.xmin = (int)i0.y; g.xmax = (int)i0.z; g.ymin = (int)i0.w; g.ymax = (int)i1.x;
        g.sx = __uint_as_float(i1.y); g.sy = __uint_as_float(i1.z); g.active = i1.w != 0u;
        const int chunk = (int)(i0.x >> 20);
        const int inst = (int)(i0.x & 0xfffffu);
        const int b = inst / A.m_rows, j = inst - b * A.m_rows;
        const int32_t* row = A.det + ((int64_t)b * m_stride + j) * 6;
        const int bw = g.xmax - g.xmin;
        const int nchunks = g.active ? (bw + kBoxCols - 1) / kBoxCols : 1;
        float* o = A.out + ((int64_t)b * Mo + j) * 11;
        double pix = 0.0, size = 0.0, colmax = 0.0, vert = 0.0;
        int cnt = 0, inter = 0;
        if (!g.active) {                                    // all-zero mask
            if (tid == 0) write_row(pix, size, vert, colmax, cnt, inter, row, A.threshold, o);
            continue;
        }
        __syncthreads();                                   // previous item done with shared memory
        const TileRef tref = tile_ref(A.src, b, j, m_stride, mh * mw, row[4], mh, mw);
        const bool bits = tref.bits && !tref.mi && mh <= kBitRows;          // CTA-uniform
        if (bits) {
            if (tid < mh) S.rows[tid] = tref.valid ? __ldg(tref.bits + tid) : 0u;
        } else {
            tref.fill(S.tile, mh, tid, kReduceThreads);
        }
        for (int y = g.ymin + tid; y < g.ymax; y += kReduceThreads) S.unit[y] = A.unit[(int64_t)b * PH + y];
        for (int i = (g.ymin >> 5) + tid; i <= ((g.ymax - 1) >> 5); i += kReduceThreads) S.rowany[i] = 0u;
        __syncthreads();
        const int c0 = chunk * kBoxCols;
        const int width = min(kBoxCols, bw - c0);
        const int cwl = width <= 32 ? 3 : (width <= 64 ? 4 : 5);            // 8 / 16 / 32 lanes per box row
        const uint32_t* rbits = A.road_bits + (int64_t)b * PH * words;
        if (bits) box_reduce<true>(S, g, c0, cwl, mh, mw, rbits, words, pix, size, colmax, cnt, inter);
        else box_reduce<false>(S, g, c0, cwl, mh, mw, rbits, words, pix, size, colmax, cnt, inter);
        __syncthreads();                                   // row flags complete
        if (nchunks == 1) {                                 // the whole box: finish here
            for (int y = g.ymin + tid; y < g.ymax; y += kReduceThreads)
                if ((S.rowany[y >> 5] >> (y & 31)) & 1u) vert = __dadd_rn(vert, (double)S.unit[y]);
            finish_row(pix, size, vert, colmax, cnt, inter, S.d, S.i, row, A.threshold, o);
            continue;
        }
        // partial results of this chunk meet the other chunks' in the instance's accumulator
        BoxAcc* acc = A.acc + ((int64_t)b * A.m_rows + j);
        uint32_t* grow = A.acc_rowany + ((int64_t)b * A.m_rows + j) * ywords;
        for (int i = (g.ymin >> 5) + tid; i <= ((g.ymax - 1) >> 5); i += kReduceThreads)
            if (S.rowany[i]) atomicOr(grow + i, S.rowany[i]);
        block_totals(pix, size, vert, colmax, cnt, inter, S.d, S.i);
        if (tid == 0) {
            atomicAdd(&acc->pix, pix);
            atomicAdd(&acc->size, size);
            atomicMax(&acc->colmax_bits, (unsigned long long)__double_as_longlong(colmax));
            atomicAdd(&acc->cnt, cnt);
            atomicAdd(&acc->inter, inter);
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) S.last = (atomicAdd(&acc->done, 1) == nchunks - 1);
        __syncthreads();
        if (S.last) {                                       // every chunk has arrived
            __threadfence();
            vert = 0.0;
            for (int y = g.ymin + tid; y < g.ymax; y += kReduceThreads)
                if ((__ldcg(grow + (y >> 5)) >> (y & 31)) & 1u) vert = __dadd_rn(vert, (double)S.unit[y]);
            double z0 = 0.0, z1 = 0.0, z3 = 0.0;
            int i0 = 0, i1 = 0;
            block_totals(z0, z1, vert, z3, i0, i1, S.d, S.i);
            if (tid == 0)
                write_row(__ldcg(&acc->pix), __ldcg(&acc->size), vert,
                          __longlong_as_double((long long)__ldcg(&acc->colmax_bits)), __ldcg(&acc->cnt),
                          __ldcg(&acc->inter), row, A.threshold, o);
        }
    }
}

}  // namespace

// ================================================================ host side ===
extern "C" int mlp_road_scan(mlp_ctx* ctx, const int32_t* seg_dev, int batch, int frame_h, int frame_w,
                             int channels, int road_channel, int crack_channel, float default_road_size,
                             float* unit_dev, uint32_t* road_bits_dev, uint32_t* crack_bits_dev,
                             int32_t* crack_box_dev, mlp_stream_t stream) {
    MLP_CHECK_ARG(ctx && seg_dev && unit_dev && road_bits_dev, "mlp_road_scan: NULL argument");
    MLP_CHECK_ARG(batch >= 1 && frame_h >= 1 && frame_w >= 1 && channels >= 1, "mlp_road_scan: bad shape");
    MLP_CHECK_ARG(frame_h <= kMaxFrameRows, "mlp_road_scan: frame height %d > %d", frame_h, kMaxFrameRows);
    MLP_CHECK_ARG(road_channel >= 0 && road_channel < channels, "mlp_road_scan: road channel %d out of range",
                  road_channel);
    MLP_CHECK_ARG(crack_channel < 0 || mlp_aligned16(crack_box_dev), "mlp_road_scan: crack_box_dev must be 16-byte aligned");
    MLP_CHECK_ARG(crack_channel < channels && (crack_channel < 0 || (crack_box_dev && crack_bits_dev)),
                  "mlp_road_scan: crack channel %d needs crack_bits_dev and crack_box_dev / is out of range",
                  crack_channel);
    MLP_CHECK_ARG((int64_t)frame_h * frame_w * channels < (1ll << 31), "mlp_road_scan: frame too large");
    DeviceGuard g(ctx->device);
    int rc = mlp_ensure_scratch(ctx, MLP_ARENA_SUMMARY, (int64_t)batch * frame_h * 16);
    if (rc) return rc;
    int2* row_ext = static_cast<int2*>(ctx->arena[MLP_ARENA_SUMMARY]);
    int2* crack_rows = crack_channel >= 0 ? row_ext + (int64_t)batch * frame_h : nullptr;
    CrackPart* crack_part = crack_channel >= 0 ? reinterpret_cast<CrackPart*>(crack_box_dev + 4) : nullptr;
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope prof(ctx, MLP_ST_ROAD_SCAN, st);
    if (crack_channel >= 0) {
        static const int32_t init[4] = {INT_MAX, INT_MAX, -1, -1};
        MLP_CUDA(cudaMemcpyAsync(crack_box_dev, init, sizeof(init), cudaMemcpyHostToDevice, st));
    }
    const int rows_per_cta = kRowsThreads / 32;
    const dim3 rgrid((frame_h + rows_per_cta - 1) / rows_per_cta, batch);
    const int crack_ch = crack_channel >= 0 ? crack_channel : 0;
    uint32_t* cbits = crack_channel >= 0 ? crack_bits_dev : nullptr;
    if (channels == 3 && frame_w % 4 == 0 && mlp_aligned16(seg_dev)) {
#define MLP_ROAD3(R, C)                                                                                     \
    road_rows3_kernel<R, C><<<rgrid, kRowsThreads, 0, st>>>(seg_dev, frame_h, frame_w, row_ext, road_bits_dev, \
                                                            cbits, crack_rows, crack_box_dev)
        switch (road_channel * 3 + crack_ch) {
            case 0: MLP_ROAD3(0, 0); break;
            case 1: MLP_ROAD3(0, 1); break;
            case 2: MLP_ROAD3(0, 2); break;
            case 3: MLP_ROAD3(1, 0); break;
            case 4: MLP_ROAD3(1, 1); break;
            case 5: MLP_ROAD3(1, 2); break;
            case 6: MLP_ROAD3(2, 0); break;
            case 7: MLP_ROAD3(2, 1); break;
            default: MLP_ROAD3(2, 2); break;
        }
#undef MLP_ROAD3
    } else {
        road_rows_kernel<<<rgrid, kRowsThreads, 0, st>>>(seg_dev, frame_h, frame_w, channels, road_channel, crack_ch,
                                                         row_ext, road_bits_dev, cbits, crack_rows, crack_box_dev);
    }
    MLP_LAUNCH_CHECK(ctx);
    road_fit_kernel<<<batch, kScanThreads, 0, st>>>(row_ext, frame_h, default_road_size, unit_dev, crack_rows,
                                                  crack_part);
    MLP_LAUNCH_CHECK(ctx);
    if (crack_channel >= 0) {
        const int words = (frame_w + 31) / 32;
        crack_cols_kernel<<<dim3(words, batch), kColsThreads, 0, st>>>(
            crack_bits_dev, unit_dev, crack_box_dev, frame_h, words, crack_part);
        MLP_LAUNCH_CHECK(ctx);
    }
    return MLP_OK;
}

extern "C" int mlp_summary_output(mlp_ctx* ctx, const int32_t* det_i32_dev, const void* masks_dev,
                                  int mask_dtype, const float* unit_dev, const uint32_t* road_bits_dev,
                                  const int32_t* crack_box_dev, int batch,
                                  int m_rows, int m_stride, const int32_t* m_dev, int frame_h, int frame_w,
                                  float include_threshold, float* out_dev, int32_t* m_out_dev,
                                  mlp_stream_t stream) {
    MLP_CHECK_ARG(ctx && det_i32_dev && masks_dev && unit_dev && road_bits_dev && out_dev,
                  "mlp_summary_output: NULL argument");
    MLP_CHECK_ARG(batch >= 1 && m_rows >= 1 && frame_h >= 1 && frame_w >= 1, "mlp_summary_output: bad shape");
    MLP_CHECK_ARG(frame_h <= kMaxFrameRows, "mlp_summary_output: frame height %d > %d", frame_h, kMaxFrameRows);
    MLP_CHECK_ARG(mask_dtype == MLP_F32 || mask_dtype == MLP_U8, "mlp_summary_output: masks must be f32 or u8");
    MLP_CHECK_ARG(mlp_aligned16(masks_dev), "mlp_summary_output: masks must be 16-byte aligned");
    DeviceGuard g(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope prof(ctx, MLP_ST_SUMMARY, st);
    SummaryArgs A;
    A.det = det_i32_dev; A.masks = masks_dev; A.unit = unit_dev; A.road_bits = road_bits_dev;
    A.crack_box = crack_box_dev; A.m_dev = m_dev;
    A.B = batch; A.m_rows = m_rows; A.m_stride = m_stride; A.PH = frame_h; A.PW = frame_w;
    A.threshold = include_threshold; A.out = out_dev; A.m_out = m_out_dev;
    const int64_t items = (int64_t)batch * m_rows;
    const int grid = (int)(items < (1ll << 30) ? items : (1ll << 30));
    // one thread per 16-byte column group of a frame row
    const int groups = mask_dtype == MLP_F32 ? (frame_w + 3) / 4 : (frame_w + 15) / 16;
    int threads = ((groups + 31) / 32) * 32;
    threads = threads < 32 ? 32 : (threads > kReduceThreads ? kReduceThreads : threads);
    if (mask_dtype == MLP_F32) instance_reduce_kernel<float><<<grid, threads, 0, st>>>(A);
    else instance_reduce_kernel<uint8_t><<<grid, threads, 0, st>>>(A);
    MLP_LAUNCH_CHECK(ctx);
    if (crack_box_dev) {                    // the crack pseudo-instance of every image, over the bitmaps
        BoxSummaryArgs X;
        memset(&X, 0, sizeof(X));
        X.det = det_i32_dev; X.unit = unit_dev; X.road_bits = road_bits_dev;
        X.crack_box = crack_box_dev; X.m_dev = m_dev; X.B = batch; X.m_rows = m_rows; X.m_stride = m_stride;
        X.mh = 1; X.mw = 1; X.PH = frame_h; X.PW = frame_w; X.chunks = 1; X.threshold = include_threshold;
        X.out = out_dev;
        box_summary_kernel<<<batch, kReduceThreads, 0, st>>>(X);
        MLP_LAUNCH_CHECK(ctx);
    }
    return MLP_OK;
}

extern "C" int mlp_tile_summary(mlp_ctx* ctx, const int32_t* det_i32_dev, const int32_t* masks_i32_dev,
                                const float* roi_masks_dev, int r_rows, const int32_t* r_dev,
                                int num_classes, const int32_t* counts_dev, int batch, int m_rows,
                                int m_stride, const int32_t* m_dev, int mask_h, int mask_w,
                                const float* unit_dev, const uint32_t* road_bits_dev,
                                const int32_t* crack_box_dev, int frame_h,
                                int frame_w, float include_threshold, float* out_dev, int32_t* m_out_dev,
                                int32_t* m_dev_out, mlp_stream_t stream) {
    MLP_CHECK_ARG(ctx && det_i32_dev && unit_dev && road_bits_dev && out_dev, "mlp_tile_summary: NULL argument");
    MLP_CHECK_ARG(masks_i32_dev || (roi_masks_dev && counts_dev && m_dev_out && num_classes >= 1 && r_rows >= 1),
                  "mlp_tile_summary: neither int32 tiles nor a prepared fused tail");
    MLP_CHECK_ARG(batch >= 1 && m_rows >= 1 && frame_h >= 1 && frame_w >= 1 &&
                      (m_stride >= m_rows || (m_stride == 0 && m_dev)),
                  "mlp_tile_summary: bad shape");
    MLP_CHECK_ARG(mask_h >= 1 && mask_w >= 1 && mask_h * mask_w <= kMaxTile, "mlp_tile_summary: mask tile %dx%d",
                  mask_h, mask_w);
    MLP_CHECK_ARG(frame_h <= kMaxFrameRows, "mlp_tile_summary: frame height %d > %d", frame_h, kMaxFrameRows);
    DeviceGuard g(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope prof(ctx, MLP_ST_SUMMARY, st);
    BoxSummaryArgs T;
    memset(&T, 0, sizeof(T));
    T.has_tiles = 1;
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
        T.src.planar = ctx->tail_planar;
        T.src.counts = counts_dev;
        T.src.confmax = ft.confmax;
        T.src.scalars = ft.scalars;
        T.src.m_out = m_dev_out;
    }
    T.det = det_i32_dev; T.unit = unit_dev; T.road_bits = road_bits_dev;
    T.crack_box = crack_box_dev; T.m_dev = m_dev;
    T.B = batch; T.m_rows = m_rows; T.m_stride = m_stride; T.mh = mask_h; T.mw = mask_w;
    T.PH = frame_h; T.PW = frame_w; T.threshold = include_threshold; T.out = out_dev; T.m_out = m_out_dev;
    // per-instance accumulators for boxes split over several CTAs (zeroed every call)
    const int ywords = (frame_h + 31) / 32;
    // ... then the work-item counter (zeroed with them) and the list
    T.chunks = (frame_w + kBoxCols - 1) / kBoxCols;
    MLP_CHECK_ARG((int64_t)batch * m_rows < (1 << 20) && T.chunks < (1 << 12), "mlp_tile_summary: too many instances / chunks");
    const int64_t acc_bytes = ((int64_t)batch * m_rows * (sizeof(BoxAcc) + (int64_t)ywords * 4) + 15) / 16 * 16 + 16;
    const int64_t item_cap = (int64_t)batch * m_rows * T.chunks;
    int rc = mlp_ensure_scratch(ctx, MLP_ARENA_BOXACC, acc_bytes + item_cap * (int64_t)sizeof(BoxItem));
    if (rc) return rc;
    MLP_CUDA(cudaMemsetAsync(ctx->arena[MLP_ARENA_BOXACC], 0, (size_t)acc_bytes, st));
    T.acc = static_cast<BoxAcc*>(ctx->arena[MLP_ARENA_BOXACC]);
    T.acc_rowany = reinterpret_cast<uint32_t*>(T.acc + (int64_t)batch * m_rows);
    T.n_items = reinterpret_cast<int32_t*>(static_cast<char*>(ctx->arena[MLP_ARENA_BOXACC]) + acc_bytes - 16);
    T.items = reinterpret_cast<BoxItem*>(static_cast<char*>(ctx->arena[MLP_ARENA_BOXACC]) + acc_bytes);
    T.item_cap = (int)item_cap;
    box_plan_kernel<<<batch, kReduceThreads, 0, st>>>(T);
    MLP_LAUNCH_CHECK(ctx);
    // one CTA per work item for the usual load (about one item per instance, a few more for wide boxes); the
    // grid-stride loop of the kernel takes whatever exceeds it
    const int64_t want = (int64_t)batch * m_rows + (int64_t)batch * m_rows / 4 + batch;
    const int64_t cap = item_cap + batch;
    const int grid = (int)(want < cap ? want : cap);
    box_summary_kernel<<<grid, kReduceThreads, 0, st>>>(T);
    MLP_LAUNCH_CHECK(ctx);
    return MLP_OK;
}
