// detect.cu — DetectionProposal (rows a5-a8 of SURVEY.md §8):
//   K1 threshold_compact : one streaming pass over cls_pred [B,N,C] (128-bit loads),
//                          candidates appended per (image,class) as 64-bit keys
//                          (descending score | anchor index)  -> a total order, so the
//                          append order does not matter.
//   K2 nms_per_class     : one CTA per (image,class): sort keys in shared memory,
//                          greedy NMS in chunks (kept list + intra-chunk bitmask), boxes
//                          fetched (or decoded from loc_pred + anchors) only for candidates.
//   K3 nms_cross_class   : one CTA per image over the concatenated per-class survivors
//                          in tf.unique first-appearance order; writes the -1 padded
//                          [B,K,6] rows, kept (n,c), counts and M.
// Reference: /root/reference/engine/layers/detection.py:482-567; TF NonMaxSuppressionV3
// semantics restated in oracle/tf_ops.py.  Compiled with -fmad=false: IoU and decode
// round every multiply/add separately so kept indices match the CPU oracle bit for bit.
#include <stdlib.h>

#include "common.cuh"

#ifdef MLP_NMS_TIMING
// tuning build only (-DMLP_NMS_TIMING): %globaltimer at the phase boundaries of the NMS kernels, CTA 0
__device__ unsigned long long g_nms_t[2][32];
__device__ __forceinline__ void nms_mark(int kernel, int slot) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        g_nms_t[kernel][slot] = t;
    }
}
extern "C" int mlp_debug_nms_timing(unsigned long long* out_host) {
    return cudaMemcpyFromSymbol(out_host, g_nms_t, sizeof(g_nms_t)) == cudaSuccess ? 0 : -2;
}
#define NMS_MARK(k, s) nms_mark(k, s)
#else
#define NMS_MARK(k, s) do { } while (0)
#endif

namespace {

constexpr int kClassThreads = 512;          // K2: several CTAs per SM (B*C CTAs in one wave)
constexpr int kCrossThreads = 1024;         // K3: one CTA per image
constexpr int kChunk = 64;                  // candidates resolved per round (one 64-bit mask row each)
constexpr int kMaskWords = kChunk / 32;
constexpr int kBlock = 512;                 // candidates whose boxes are fetched at once (8 chunks): ONE memory round
                                            // trip per block instead of one per chunk - the kernels are latency-bound
constexpr uint64_t kKeyPad = ~0ull;

struct BoxC { float ymin, xmin, ymax, xmax, area; };

// NormalizeBoxes with shape = ones (detection.py:367-374, :488) then the min/max
// canonicalisation and area of TF's IOU().
__device__ __forceinline__ BoxC corners_of(const float4 b) {
    float hw = __fmul_rn(b.z, 0.5f), hh = __fmul_rn(b.w, 0.5f);
    float y1 = __fsub_rn(b.y, hh), x1 = __fsub_rn(b.x, hw);
    float y2 = __fadd_rn(b.y, hh), x2 = __fadd_rn(b.x, hw);
    BoxC c;
    c.ymin = fminf(y1, y2); c.xmin = fminf(x1, x2);
    c.ymax = fmaxf(y1, y2); c.xmax = fmaxf(x1, x2);
    c.area = __fmul_rn(__fsub_rn(c.ymax, c.ymin), __fsub_rn(c.xmax, c.xmin));
    return c;
}

// IoU(a,b) > thr, TF operation order, true division.  (A fast approximate division with an exact fall-back near
// the threshold was measured: no gain - the chunk phases are not bound by the division, profiles/nms_phases_r02.txt.)
__device__ __forceinline__ bool iou_exceeds(float aymin, float axmin, float aymax, float axmax,
                                            float aarea, float bymin, float bxmin, float bymax,
                                            float bxmax, float barea, float thr) {
    if (aarea <= 0.0f || barea <= 0.0f) return 0.0f > thr;
    float ih = fmaxf(__fsub_rn(fminf(aymax, bymax), fmaxf(aymin, bymin)), 0.0f);
    float iw = fmaxf(__fsub_rn(fminf(axmax, bxmax), fmaxf(axmin, bxmin)), 0.0f);
    float inter = __fmul_rn(ih, iw);
    if (inter == 0.0f) return 0.0f > thr;
    float iou = __fdiv_rn(inter, __fsub_rn(__fadd_rn(aarea, barea), inter));
    return iou > thr;
}

__device__ __forceinline__ uint64_t make_key(float score, uint32_t low) {
    return ((uint64_t)(~float_ordered(score)) << 32) | low;
}
__device__ __forceinline__ float key_score(uint64_t key) {
    return float_from_ordered(~(uint32_t)(key >> 32));
}

// Optional fusion of MaskDistribute (a9) and the PyramidRoiAlign plan (a10) into the epilogue of
// the cross-class NMS kernel: the kept rows of an image are still on chip there.
struct FusedPlan {
    int enabled;
    int L;                  // pyramid levels (max_k + 1)
    float base_eps;         // base_size + K.epsilon()
    float max_k;
    float* dist;            // [B,K,7]
    RoiRec* roi_rec;        // [L,B,K] slot -> row and box (RoiRec, common.cuh)
    int32_t* level_counts;  // [L,B]
    int32_t* level_m;       // [L+1]
};

// ------------------------------------------------------------------ K1 -------
// cand_keys [G][cap] (G = B*C), cand_count [G].  Every CTA streams one contiguous slice of
// cls_pred with 4 independent 128-bit loads in flight per thread.  Hits (0.2-2 % of the
// scores) are parked in a shared-memory queue as (flat index, score) and turned into
// per-(image,class) appends once per CTA, so the streaming loop carries no global atomics,
// no 64-bit divisions and almost no divergence.
constexpr int kK1Threads = 256;
constexpr int kK1Queue = 2048;
constexpr int kK1Groups = 512;           // (image,class) groups a CTA slice may touch on the fast flush path

__device__ __forceinline__ void k1_append(uint32_t e, float s, int N, int C, uint64_t* cand_keys,
                                          int32_t* cand_count, int64_t cap) {
    const uint32_t bn = e / (uint32_t)C;
    const int c = (int)(e - bn * (uint32_t)C);
    const uint32_t b = bn / (uint32_t)N;
    const uint32_t n = bn - b * (uint32_t)N;
    const int g = (int)b * C + c;
    const int pos = atomicAdd(cand_count + g, 1);
    MLP_BOUND(n, N);
    if (pos < cap) cand_keys[(int64_t)g * cap + pos] = make_key(s, n);
}

__global__ void __launch_bounds__(kK1Threads)
threshold_compact_kernel(const float* __restrict__ cls, uint32_t total, int N, int C, float thr,
                         uint64_t* __restrict__ cand_keys, int32_t* __restrict__ cand_count,
                         int64_t cap, int32_t* __restrict__ m_dev, int32_t* __restrict__ zero_ptr,
                         int zero_n) {
    __shared__ uint2 s_q[kK1Queue];
    __shared__ int s_hist[kK1Groups], s_cursor[kK1Groups], s_gbase[kK1Groups];
    __shared__ int s_qn;
    if (blockIdx.x == 0 && threadIdx.x == 0 && m_dev) *m_dev = 1;
    if (blockIdx.x == 0 && (int)threadIdx.x < zero_n) zero_ptr[threadIdx.x] = 0;   // level_m for the fused plan
    if (threadIdx.x == 0) s_qn = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    // Warp-aggregated append of one tile's hits: every lane holds the 16-bit mask of its 16
    // scores; one shuffle scan + ONE shared atomic per warp reserves queue slots for all of them.
    auto flush_tile = [&](const float4 (&v)[4], uint32_t base) {
        uint32_t mask = 0;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            mask |= (uint32_t)(v[u].x >= thr) << (4 * u);
            mask |= (uint32_t)(v[u].y >= thr) << (4 * u + 1);
            mask |= (uint32_t)(v[u].z >= thr) << (4 * u + 2);
            mask |= (uint32_t)(v[u].w >= thr) << (4 * u + 3);
        }
        if (__ballot_sync(0xffffffffu, mask != 0) == 0) return;        // common: no hit in 512 scores
        const int n = __popc(mask);
        int incl = n;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        int qbase = 0;
        if (lane == 31) qbase = atomicAdd(&s_qn, total);
        qbase = __shfl_sync(0xffffffffu, qbase, 31);
        int pos = qbase + incl - n;
        if (mask == 0) return;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float comp[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if ((mask >> (4 * u + c)) & 1u) {
                    const uint32_t e = ((base + u * kK1Threads + threadIdx.x) << 2) + c;
                    if (pos < kK1Queue) s_q[pos] = make_uint2(e, __float_as_uint(comp[c]));
                    else k1_append(e, comp[c], N, C, cand_keys, cand_count, cap);   // queue full
                    ++pos;
                }
            }
        }
    };
    auto hit = [&](uint32_t e, float sc) {                             // scalar tail only
        const int pos = atomicAdd(&s_qn, 1);
        if (pos < kK1Queue) s_q[pos] = make_uint2(e, __float_as_uint(sc));
        else k1_append(e, sc, N, C, cand_keys, cand_count, cap);
    };
    const uint32_t total4 = total >> 2;
    const float4* cls4 = reinterpret_cast<const float4*>(cls);
    const float kNaN = __int_as_float(0x7fc00000);     // NaN >= thr is false for every thr
    // contiguous slice per CTA, rounded to whole tiles of kK1Threads*4 vectors
    const uint32_t tile = kK1Threads * 4;
    const uint32_t tiles = (total4 + tile - 1) / tile;
    const uint32_t tiles_per_cta = (tiles + gridDim.x - 1) / gridDim.x;
    uint32_t v0 = blockIdx.x * tiles_per_cta * tile;
    uint32_t v1 = v0 + tiles_per_cta * tile;
    if (v1 > total4) v1 = total4;
    if (v0 > v1) v0 = v1;
    // software pipeline: the next tile's four 128-bit loads are in flight while the current
    // tile is compared against the threshold
    float4 cur[4], nxt[4];
    auto load_tile = [&](float4 (&v)[4], uint32_t base) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint32_t i = base + u * kK1Threads + threadIdx.x;
            v[u] = (i < v1) ? ldg_stream_f4(cls4 + i) : make_float4(kNaN, kNaN, kNaN, kNaN);
        }
    };
    if (v0 < v1) load_tile(cur, v0);
    for (uint32_t base = v0; base < v1; base += tile) {
        if (base + tile < v1) load_tile(nxt, base + tile);
        flush_tile(cur, base);
#pragma unroll
        for (int u = 0; u < 4; ++u) cur[u] = nxt[u];
    }
    if (blockIdx.x == gridDim.x - 1) {             // scalar tail (total not a multiple of 4)
        for (uint32_t e = (total4 << 2) + threadIdx.x; e < total; e += kK1Threads) {
            const float s = cls[e];
            if (s >= thr) hit(e, s);
        }
    }
    __syncthreads();
    const int qn = s_qn < kK1Queue ? s_qn : kK1Queue;
    // Flush: a CTA's slice covers at most a couple of images, i.e. a handful of (image,class)
    // groups.  Hits are counted per group in shared memory first so that each CTA issues ONE
    // global atomicAdd per group instead of one per hit: 626 k same-address atomics on 192
    // counters (stress config) serialise in L2 and used to cost more than the streaming pass.
    const uint32_t NC = (uint32_t)N * (uint32_t)C;
    const uint32_t e_first = v0 << 2;
    const uint32_t e_last = (blockIdx.x == gridDim.x - 1) ? total - 1 : (v1 > v0 ? (v1 << 2) - 1 : e_first);
    const uint32_t b_first = e_first / NC;
    const uint32_t nb = e_last / NC - b_first + 1;
    if (qn > 0 && nb * (uint32_t)C <= (uint32_t)kK1Groups) {
        const int ng = (int)nb * C;
        for (int i = threadIdx.x; i < ng; i += kK1Threads) { s_hist[i] = 0; s_cursor[i] = 0; }
        __syncthreads();
        for (int i = threadIdx.x; i < qn; i += kK1Threads) {
            const uint32_t e = s_q[i].x;
            const uint32_t bn = e / (uint32_t)C;
            const int c = (int)(e - bn * (uint32_t)C);
            const int gl = (int)(bn / (uint32_t)N - b_first) * C + c;
            atomicAdd(&s_hist[gl], 1);
        }
        __syncthreads();
        for (int i = threadIdx.x; i < ng; i += kK1Threads)
            if (s_hist[i] > 0) s_gbase[i] = atomicAdd(cand_count + (int)b_first * C + i, s_hist[i]);
        __syncthreads();
        for (int i = threadIdx.x; i < qn; i += kK1Threads) {
            const uint32_t e = s_q[i].x;
            const uint32_t bn = e / (uint32_t)C;
            const int c = (int)(e - bn * (uint32_t)C);
            const uint32_t b = bn / (uint32_t)N;
            const uint32_t n = bn - b * (uint32_t)N;
            const int gl = (int)(b - b_first) * C + c;
            const int pos = s_gbase[gl] + atomicAdd(&s_cursor[gl], 1);
            if (pos < cap)
                cand_keys[(int64_t)((int)b * C + c) * cap + pos] = make_key(__uint_as_float(s_q[i].y), n);
        }
    } else {
        for (int i = threadIdx.x; i < qn; i += kK1Threads)
            k1_append(s_q[i].x, __uint_as_float(s_q[i].y), N, C, cand_keys, cand_count, cap);
    }
}

// --------------------------------------------------------- sort helpers ------
__device__ void bitonic_sort_smem(uint64_t* s, int n_pow2) {
    for (int k = 2; k <= n_pow2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n_pow2; i += blockDim.x) {
                int ixj = i ^ j;
                if (ixj > i) {
                    uint64_t a = s[i], b = s[ixj];
                    bool asc = (i & k) == 0;
                    if ((a > b) == asc) { s[i] = b; s[ixj] = a; }
                }
            }
            __syncthreads();
        }
    }
}

// The same network with ONE key per thread held in a register (n_pow2 <= blockDim.x, the usual case: a few
// hundred candidates per group): compare-exchange partners less than 32 apart meet by warp shuffle, only
// the strides >= 32 go through shared memory - 10 block-wide exchanges for 512 keys instead of 45, which was
// 40 % of the run time of both NMS kernels (profiles/ncu_summary_r02*.md).
__device__ void bitonic_sort_reg(uint64_t* s, int n_pow2) {
    const int t = threadIdx.x;
    const bool in = t < n_pow2;
    const bool warp_in = (t & ~31) < n_pow2;           // warps wholly past the keys only keep the barriers company
    const uint64_t v0 = in ? s[t] : kKeyPad;
    uint32_t hi = (uint32_t)(v0 >> 32), lo = (uint32_t)v0;     // two 32-bit halves: shuffles and selects are 32-bit
    for (int k = 2; k <= n_pow2; k <<= 1) {
        const bool asc = (t & k) == 0;
        for (int j = k >> 1; j > 0; j >>= 1) {
            uint32_t ohi = hi, olo = lo;
            if (j >= 32) {
                __syncthreads();                       // partners have read the previous exchange
                if (in) s[t] = ((uint64_t)hi << 32) | lo;
                __syncthreads();
                if (in) {
                    const uint64_t o = s[t ^ j];
                    ohi = (uint32_t)(o >> 32); olo = (uint32_t)o;
                }
            } else if (warp_in) {
                ohi = __shfl_xor_sync(0xffffffffu, hi, j);
                olo = __shfl_xor_sync(0xffffffffu, lo, j);
            }
            if (warp_in) {
                // the lower index of a pair keeps the minimum in an ascending run, the maximum in a descending one
                const bool take_min = asc == ((t & j) == 0);
                const bool lt = hi < ohi || (hi == ohi && lo < olo);
                const bool keep = lt == take_min;      // keys are unique: never equal (pad lanes: equal, keep either)
                hi = keep ? hi : ohi;
                lo = keep ? lo : olo;
            }
        }
    }
    __syncthreads();
    if (in) s[t] = ((uint64_t)hi << 32) | lo;
    __syncthreads();
}

// Smallest pivot P such that #{key in gkeys : lo_excl < key <= P (or any key when
// !have_lo)} == want.  Keys are unique, 1 <= want <= number of such keys.  MSB-first
// 8-bit radix select; hist is a 256-int shared array, bcast a 2-word shared array.
__device__ uint64_t radix_select_pivot(const uint64_t* gkeys, int cnt, bool have_lo,
                                       uint64_t lo_excl, int want, int* hist,
                                       unsigned long long* bcast) {
    uint64_t prefix = 0, prefix_mask = 0;
    int remaining = want;
    for (int shift = 56; shift >= 0; shift -= 8) {
        for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
        __syncthreads();
        for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
            uint64_t k = gkeys[i];
            if ((!have_lo || k > lo_excl) && (k & prefix_mask) == prefix)
                atomicAdd(&hist[(int)((k >> shift) & 0xff)], 1);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int acc = 0, d = 0;
            for (; d < 256; ++d) {
                if (acc + hist[d] >= remaining) break;
                acc += hist[d];
            }
            bcast[0] = (unsigned long long)d;
            bcast[1] = (unsigned long long)acc;
        }
        __syncthreads();
        int d = (int)bcast[0];
        remaining -= (int)bcast[1];
        prefix |= ((uint64_t)d) << shift;
        prefix_mask |= 0xffull << shift;
        __syncthreads();
    }
    return prefix;     // the want-th smallest qualifying key itself
}

// ------------------------------------------------------------ NMS core -------
struct NmsSmem {
    uint64_t* skeys;        // [sort_cap]
    float4* k_crn; float* k_area;   // [max_out] kept boxes (ymin,xmin,ymax,xmax) + area: one LDS.128 + one LDS.32 per test
    float4* c_crn; float* c_area;   // [kBlock]  candidates of the current block (chunks index into it)
    float4* c_box;          // [kBlock] cx,cy,w,h of the block's candidates
    uint32_t* c_mask;       // [2][kChunk][kMaskWords]  (double-buffered per chunk)
    int* c_supp;            // [2][kChunk]
    uint32_t* keptw;        // [kMaskWords]
    int* hist;              // [256]
    unsigned long long* bcast;   // [2]
    int* misc;              // [4]: 0 kept, 1 gather counter
};

__host__ __device__ inline size_t nms_smem_bytes(int sort_cap, int max_out) {
    size_t b = 0;
    b += (size_t)sort_cap * 8;
    b += (size_t)max_out * 4 * 5;
    b += (size_t)kBlock * 4 * 5;
    b += (size_t)kBlock * 16;
    b += (size_t)2 * kChunk * kMaskWords * 4;
    b += (size_t)2 * kChunk * 4;
    b += kMaskWords * 4;
    b += 256 * 4;
    b += 2 * 8;
    b += 4 * 4;
    return b + 64;
}

__device__ inline NmsSmem carve_smem(unsigned char* base, int sort_cap, int max_out) {
    NmsSmem S;
    unsigned char* p = base;
    S.skeys = reinterpret_cast<uint64_t*>(p); p += (size_t)sort_cap * 8;
    S.c_box = reinterpret_cast<float4*>(p);   p += (size_t)kBlock * 16;
    S.bcast = reinterpret_cast<unsigned long long*>(p); p += 16;
    S.k_crn = reinterpret_cast<float4*>(p); p += (size_t)max_out * 16;     // 16-byte aligned (after c_box/bcast)
    S.c_crn = reinterpret_cast<float4*>(p); p += (size_t)kBlock * 16;
    S.k_area = reinterpret_cast<float*>(p); p += (size_t)max_out * 4;
    S.c_area = reinterpret_cast<float*>(p); p += kBlock * 4;
    S.c_mask = reinterpret_cast<uint32_t*>(p); p += (size_t)2 * kChunk * kMaskWords * 4;
    S.c_supp = reinterpret_cast<int*>(p); p += 2 * kChunk * 4;
    S.keptw = reinterpret_cast<uint32_t*>(p); p += kMaskWords * 4;
    S.hist = reinterpret_cast<int*>(p); p += 256 * 4;
    S.misc = reinterpret_cast<int*>(p); p += 16;
    return S;
}

// Greedy NMS over `cnt` unique keys in global memory (any order).  Pops keys in
// ascending key order (= descending score, ties -> lower low-word), suppresses a
// candidate iff IoU > thr with an earlier kept box, stops at max_out.  fetch(low)
// returns the (cx,cy,w,h) box of a key; emit(rank, key, box) is called once per kept
// box by exactly one thread.  Returns the kept count (block-uniform).
template <int kThreads, class Fetch, class Emit>
__device__ int nms_core(const uint64_t* gkeys, int cnt, int sort_cap, float thr,
                        int max_out, const NmsSmem& S, Fetch fetch, Emit emit, bool staged = false, int tk = 0) {
    (void)tk;
    int mark_chunk = 8;
    constexpr int kLanesPerCand = kThreads / kChunk;   // threads that split the kept list
    static_assert(kThreads % kChunk == 0 && kChunk % kLanesPerCand == 0 && kChunk == 64, "bad NMS geometry");
    const int tid = threadIdx.x;
    if (tid == 0) S.misc[0] = 0;
    if (tid < 2 * kChunk) { S.c_supp[tid] = 0; S.c_mask[tid * 2] = 0u; S.c_mask[tid * 2 + 1] = 0u; }
    __syncthreads();
    int buf = 0;                     // which half of c_supp / c_mask the current chunk uses
    int kept = 0;
    int done = 0;                    // keys consumed so far (in sorted order)
    uint64_t lo = 0;                 // largest key consumed so far
    while (done < cnt && kept < max_out) {
        // ---- stage a super-chunk of up to sort_cap smallest unconsumed keys ----
        const int remaining = cnt - done;
        const int sc = remaining < sort_cap ? remaining : sort_cap;
        const bool have_lo = done > 0;
        if (!have_lo && cnt <= sort_cap) {
            // everything fits: plain copy (or already in S.skeys when the caller staged it), no
            // selection pass and no shared-memory counter
            if (!staged)
                for (int i = tid; i < cnt; i += blockDim.x) S.skeys[i] = gkeys[i];
        } else {
            uint64_t pivot = kKeyPad;
            if (remaining > sort_cap)
                pivot = radix_select_pivot(gkeys, cnt, have_lo, lo, sc, S.hist, S.bcast);
            if (tid == 0) S.misc[1] = 0;
            __syncthreads();
            for (int i = tid; i < cnt; i += blockDim.x) {
                uint64_t k = gkeys[i];
                if ((!have_lo || k > lo) && k <= pivot) {
                    int pos = atomicAdd(&S.misc[1], 1);
                    if (pos < sort_cap) S.skeys[pos] = k;
                }
            }
        }
        int p2 = 1;
        while (p2 < sc) p2 <<= 1;
        __syncthreads();
        for (int i = sc + tid; i < p2; i += blockDim.x) S.skeys[i] = kKeyPad;
        __syncthreads();
        NMS_MARK(tk, 2);
        // (a rank sort - every key counts the smaller ones, one scatter - was measured at 7.4 us against the 4.9 us of
        // the register network for 512 keys: profiles/nms_phases_r02.txt)
        if (p2 <= (int)blockDim.x) bitonic_sort_reg(S.skeys, p2);
        else bitonic_sort_smem(S.skeys, p2);
        NMS_MARK(tk, 3);
        lo = S.skeys[sc - 1];
        done += sc;

        // ---- blocks of kBlock candidates: all their boxes are fetched (decoded) in one round trip ----
        for (int pb = 0; pb < sc && kept < max_out; pb += kBlock) {
            const int pn = (sc - pb) < kBlock ? (sc - pb) : kBlock;
            __syncthreads();                          // previous block's chunks are done with c_*
            for (int i = tid; i < pn; i += blockDim.x) {
                MLP_BOUND(pb + i, sort_cap);
                const float4 bx = fetch((uint32_t)S.skeys[pb + i]);
                const BoxC c = corners_of(bx);
                S.c_box[i] = bx;
                S.c_crn[i] = make_float4(c.ymin, c.xmin, c.ymax, c.xmax);
                S.c_area[i] = c.area;
            }
            __syncthreads();                          // c_* of this block are in shared memory
            NMS_MARK(tk, 4);
            // ---- chunks of kChunk (= 64) candidates, two barriers each ----
            // phase 1 (all threads): candidate t against the kept list (its slice r of it) AND against the
            //   earlier candidates of the chunk - the two do not depend on each other: a mask bit that
            //   points at a candidate which turns out dead is never looked at by the resolve.
            // phase 2: warp 0 resolves the chunk sequentially out of registers and appends the newly kept
            //   boxes; two other warps clear the other half of c_supp / c_mask for the next chunk.
            for (int base = 0; base < pn && kept < max_out; base += kChunk) {
                const int m = (pn - base) < kChunk ? (pn - base) : kChunk;
                int* c_supp = S.c_supp + buf * kChunk;
                uint32_t* c_mask = S.c_mask + buf * kChunk * kMaskWords;
                const int t = tid % kChunk, r = tid / kChunk;
                if (t < m) {
                    const float4 tc = S.c_crn[base + t];
                    const float tymin = tc.x, txmin = tc.y, tymax = tc.z, txmax = tc.w;
                    const float tarea = S.c_area[base + t];
                    for (int j = r; j < kept; j += kLanesPerCand) {
                        const float4 kc = S.k_crn[j];
                        if (iou_exceeds(tymin, txmin, tymax, txmax, tarea, kc.x, kc.y, kc.z, kc.w, S.k_area[j],
                                        thr)) {
                            c_supp[t] = 1;
                            break;
                        }
                    }
                    // bit u of row t set iff u < t and IoU(t,u) > thr
                    constexpr int kSlice = kChunk / kLanesPerCand;
                    const int u0 = r * kSlice;
                    const int u1 = (u0 + kSlice < t) ? (u0 + kSlice) : t;
                    uint32_t lo = 0, hi = 0;
                    for (int u = u0; u < u1; ++u) {
                        const float4 uc = S.c_crn[base + u];
                        if (iou_exceeds(tymin, txmin, tymax, txmax, tarea, uc.x, uc.y, uc.z, uc.w,
                                        S.c_area[base + u], thr)) {
                            if (u < 32) lo |= 1u << u; else hi |= 1u << (u - 32);
                        }
                    }
                    if (lo) atomicOr(&c_mask[t * 2], lo);
                    if (hi) atomicOr(&c_mask[t * 2 + 1], hi);
                }
                __syncthreads();
                if (mark_chunk < 12) NMS_MARK(tk, 16 + 3 * (mark_chunk - 8));
                if (tid < 32) {
                    // sequential resolve.  Rows live in registers (lane l: candidates l and l+32) and reach
                    // every lane by shuffle, so the loop-carried chain is two ANDs, a compare and an OR per
                    // candidate - no shared-memory latency on it.
                    const uint32_t alive_lo = __ballot_sync(0xffffffffu, tid < m && c_supp[tid] == 0);
                    const uint32_t alive_hi = __ballot_sync(0xffffffffu, tid + 32 < m && c_supp[tid + 32] == 0);
                    const uint32_t rowA_lo = c_mask[tid * 2];
                    const uint32_t rowB_lo = c_mask[(tid + 32) * 2];
                    const uint32_t rowB_hi = c_mask[(tid + 32) * 2 + 1];
                    // the cap max_out stays off the loop-carried chain: resolve as if there were none, then keep only
                    // the first (max_out - kept) winners - earlier decisions never depend on later candidates
                    uint32_t kept_lo = 0, kept_hi = 0;
#pragma unroll
                    for (int q = 0; q < 32; ++q) {
                        const uint32_t rl = __shfl_sync(0xffffffffu, rowA_lo, q);
                        const uint32_t bit = alive_lo & (1u << q);
                        kept_lo |= (rl & kept_lo) ? 0u : bit;
                    }
#pragma unroll
                    for (int q = 0; q < 32; ++q) {
                        const uint32_t rl = __shfl_sync(0xffffffffu, rowB_lo, q);
                        const uint32_t rh = __shfl_sync(0xffffffffu, rowB_hi, q);
                        const uint32_t bit = alive_hi & (1u << q);
                        kept_hi |= ((rl & kept_lo) | (rh & kept_hi)) ? 0u : bit;
                    }
                    int room = max_out - kept;
                    if (__popc(kept_lo) > room) kept_lo = room > 0 ? kept_lo & ((2u << __fns(kept_lo, 0, room)) - 1u) : 0u;
                    room -= __popc(kept_lo);
                    if (__popc(kept_hi) > room) kept_hi = room > 0 ? kept_hi & ((2u << __fns(kept_hi, 0, room)) - 1u) : 0u;
                    const int k = kept + __popc(kept_lo) + __popc(kept_hi);
                    if (tid == 0) S.misc[0] = k;
                    if (mark_chunk < 12) NMS_MARK(tk, 17 + 3 * (mark_chunk - 8));
                    // append the newly kept boxes in order and emit them (lane l: candidates l and l+32)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const uint32_t wv = h ? kept_hi : kept_lo;
                        if ((wv >> tid) & 1u) {
                            const int cand = tid + 32 * h;
                            const int rank = kept + (h ? __popc(kept_lo) : 0) + __popc(wv & ((1u << tid) - 1u));
                            MLP_BOUND(rank, max_out);
                            MLP_BOUND(base + cand, kBlock);
                            S.k_crn[rank] = S.c_crn[base + cand];
                            S.k_area[rank] = S.c_area[base + cand];
                            emit(rank, S.skeys[pb + base + cand], S.c_box[base + cand]);
                        }
                    }
                    if (mark_chunk < 12) NMS_MARK(tk, 18 + 3 * (mark_chunk - 8));
                } else if (tid < 32 + kChunk) {
                    const int i = tid - 32;            // clear the other half for the next chunk
                    S.c_supp[(buf ^ 1) * kChunk + i] = 0;
                    S.c_mask[((buf ^ 1) * kChunk + i) * 2] = 0u;
                    S.c_mask[((buf ^ 1) * kChunk + i) * 2 + 1] = 0u;
                }
                __syncthreads();
                kept = S.misc[0];
                buf ^= 1;
                if (mark_chunk < 30) NMS_MARK(tk, mark_chunk++);
            }
        }
    }
    return kept;
}

// ------------------------------------------------------------------ K2 -------
struct FetchBoxes {
    const float4* boxes;     // [N] of this image
    __device__ float4 operator()(uint32_t n) const { return boxes[n]; }
};
struct FetchDecode {
    const PriorDev* P;
    const float4* loc;       // [N] of this image
    __device__ float4 operator()(uint32_t n) const {
        MLP_BOUND(n, P->total);
        return restore_box(loc[n], prior_anchor(*P, (int)n));
    }
};

struct DetScratch {
    uint64_t* cand_keys;     // [G][cap]
    int32_t* cand_count;     // [G]
    int32_t* cls_kept;       // [G]
    int32_t* min_n;          // [G]
    float4* rec_box;         // [G][max_out]
    int4* rec_sn;            // [G][max_out] (score bits, n, 0, 0): 16-byte records, bulk-copyable per group
    uint64_t* cat_keys;      // [B][C*max_out]
    float4* cat_box;         // [B][C*max_out]
    int2* cat_sn;            // [B][C*max_out] (score bits, n)
    int32_t* cat_c;          // [B][C*max_out]
    int32_t* img_done;       // [B] arrival counters of the fused cross-class pass
    int64_t zero_bytes;      // cand_count .. img_done: zeroed with one memset per call
    int64_t cap;
};

// Arguments of the cross-class pass when it is fused behind the per-class NMS: the class CTA of an image that
// finishes last (arrival counter per image) runs the image's cross-class NMS itself, so the second kernel - its
// launch, its drain and its 32-CTA grid - disappears and early images do not wait for late ones.
struct CrossArgs {
    int fuse;
    float thr;
    int sort_cap;
    float* det;
    int32_t* keep;
    int32_t* counts;
    int32_t* m_dev;
    int32_t* img_done;       // [B] arrival counters, zero before the launch; the last CTA resets its image's
    FusedPlan F;
};

template <int kThreads>
__device__ void cross_class_body(unsigned char* smem_raw, int b, int B, int C, float thr, int max_out, int sort_cap,
                                 const DetScratch& D, float* __restrict__ det, int32_t* __restrict__ keep,
                                 int32_t* __restrict__ counts, int32_t* __restrict__ m_dev, const FusedPlan& F,
                                 int use_tma);

template <bool kDecode>
__global__ void __launch_bounds__(kClassThreads, 2)
nms_per_class_kernel(const __grid_constant__ PriorDev P, const float4* __restrict__ boxes_or_loc,
                     int N, int C, float thr, int max_out, int sort_cap, DetScratch D, const CrossArgs X) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    NmsSmem S = carve_smem(smem_raw, sort_cap, max_out);
    NMS_MARK(0, 0);
    const int g = blockIdx.x;
    const int b = g / C;
    int cnt = D.cand_count[g];
    if ((int64_t)cnt > D.cap) cnt = (int)D.cap;
    const uint64_t* gkeys = D.cand_keys + (int64_t)g * D.cap;

    // first anchor index at which this (image,class) appears: orders the groups like
    // tf.unique does in the row-major (b,n,c) scan (detection.py:519-520).  When the group fits the
    // sort buffer (the usual case) the same pass stages the keys, so they are read once.
    const bool staged = cnt <= sort_cap;
    {
        int mn = 0x7fffffff;
        for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
            const uint64_t k = gkeys[i];
            if (staged) S.skeys[i] = k;
            int n = (int)(uint32_t)k;
            mn = n < mn ? n : mn;
        }
        for (int o = 16; o > 0; o >>= 1) {
            int v = __shfl_xor_sync(0xffffffffu, mn, o);
            mn = v < mn ? v : mn;
        }
        if (threadIdx.x == 0) S.hist[0] = 0x7fffffff;
        __syncthreads();
        if ((threadIdx.x & 31) == 0) atomicMin(&S.hist[0], mn);
        __syncthreads();
        if (threadIdx.x == 0) D.min_n[g] = S.hist[0];
        __syncthreads();
    }
    if (cnt == 0) {
        if (threadIdx.x == 0) D.cls_kept[g] = 0;
    } else {
        NMS_MARK(0, 1);
        float4* rec_box = D.rec_box + (int64_t)g * max_out;
        int4* rec_sn = D.rec_sn + (int64_t)g * max_out;
        auto emit = [&](int rank, uint64_t key, const float4& bx) {
            rec_box[rank] = bx;
            rec_sn[rank] = make_int4(__float_as_int(key_score(key)), (int)(uint32_t)key, 0, 0);
        };
        int kept;
        if (kDecode) {
            FetchDecode f{&P, boxes_or_loc + (int64_t)b * N};
            kept = nms_core<kClassThreads>(gkeys, cnt, sort_cap, thr, max_out, S, f, emit, staged);
        } else {
            FetchBoxes f{boxes_or_loc + (int64_t)b * N};
            kept = nms_core<kClassThreads>(gkeys, cnt, sort_cap, thr, max_out, S, f, emit, staged);
        }
        if (threadIdx.x == 0) D.cls_kept[g] = kept;
    }
    NMS_MARK(0, 31);
    if (!X.fuse) return;
    // ---- arrival: the last class of image b to finish runs the image's cross-class pass (CTA-uniform branch)
    __shared__ int s_is_last;
    __threadfence();                                   // this group's records are visible before it arrives
    __syncthreads();
    if (threadIdx.x == 0) s_is_last = atomicAdd(X.img_done + b, 1) == C - 1;
    __syncthreads();
    if (!s_is_last) return;
    __threadfence();
    if (threadIdx.x == 0) X.img_done[b] = 0;           // ready for the next launch
    cross_class_body<kClassThreads>(smem_raw, b, gridDim.x / C, C, X.thr, max_out, X.sort_cap, D, X.det, X.keep,
                                    X.counts, X.m_dev, X.F, 0);
}

// ------------------------------------------------------------------ K3 -------
struct FetchCat {
    const float4* box;
    __device__ float4 operator()(uint32_t p) const { return box[p]; }
};

// Survivors staged in shared memory by bulk copies (the north star's "TMA-tiled IoU blocks", where a contiguous tile
// exists: every class's survivors are one contiguous run of 16-byte records).  kStageCap survivors at most.
constexpr int kStageCap = kBlock;

template <int kThreads>
__device__ void cross_class_body(unsigned char* smem_raw, int b, int B, int C, float thr, int max_out, int sort_cap,
                                 const DetScratch& D, float* __restrict__ det, int32_t* __restrict__ keep,
                                 int32_t* __restrict__ counts, int32_t* __restrict__ m_dev, const FusedPlan& F,
                                 int use_tma) {
    NmsSmem S = carve_smem(smem_raw, sort_cap, max_out);
    // behind the NMS work space: staging area of the bulk copies, boxes then (score, anchor) records
    float4* t_box = reinterpret_cast<float4*>(smem_raw + ((nms_smem_bytes(sort_cap, max_out) + 127) & ~(size_t)127));
    int4* t_sn = reinterpret_cast<int4*>(t_box + kStageCap);
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ unsigned char s_cat_c[kStageCap];
    __shared__ int s_order[256];
    __shared__ int s_off[257];
    const int tid = threadIdx.x;
    NMS_MARK(1, 0);
    // groups of this image in first-appearance order: sort by (min_n, c).  The per-class scalars are
    // fetched by C threads at once (one round trip), thread 0 orders them out of shared memory.
    __shared__ int s_cnt[256], s_min[256], s_kept[256];
    if (tid < C) {
        s_cnt[tid] = __ldcg(D.cand_count + b * C + tid);     // .cg: siblings of the same launch wrote these when fused
        s_min[tid] = __ldcg(D.min_n + b * C + tid);
        s_kept[tid] = __ldcg(D.cls_kept + b * C + tid);
    }
    __syncthreads();
    if (tid == 0) {
        int ng = 0;
        for (int c = 0; c < C; ++c) {
            if (s_cnt[c] > 0) {
                int mn = s_min[c];
                int pos = ng++;
                while (pos > 0 && s_min[s_order[pos - 1]] > mn) {
                    s_order[pos] = s_order[pos - 1];
                    --pos;
                }
                s_order[pos] = c;
            }
        }
        int off = 0;
        for (int i = 0; i < ng; ++i) {
            s_off[i] = off;
            off += s_kept[s_order[i]];
        }
        s_off[ng] = off;
        S.misc[2] = ng;
        S.misc[3] = off;
    }
    __syncthreads();
    const int ng = S.misc[2];
    const int total = S.misc[3];
    const int64_t cat_base = (int64_t)b * C * max_out;
    uint64_t* cat_keys = D.cat_keys + cat_base;
    float4* cat_box = D.cat_box + cat_base;
    int2* cat_sn = D.cat_sn + cat_base;
    int32_t* cat_c = D.cat_c + cat_base;
    NMS_MARK(1, 1);
    const bool staged = total <= sort_cap;
    const bool tma = use_tma && staged && total > 0 && total <= kStageCap;      // CTA-uniform
    float* det_b = det + (int64_t)b * max_out * 6;
    int32_t* keep_b = keep ? keep + (int64_t)b * max_out * 2 : nullptr;
    int kept = 0;
    if (tma) {
        // ---- one cp.async.bulk pair per class run: survivors land in shared memory, nothing is written back
        if (tid == 0) {
            mbar_init(&s_bar, 1);
            mbar_expect_tx(&s_bar, (uint32_t)total * 32u);
        }
        __syncthreads();
        if (tid < ng) {
            const int c = s_order[tid];
            const int n_g = s_off[tid + 1] - s_off[tid];
            if (n_g > 0) {
                const int64_t src = (int64_t)(b * C + c) * max_out;
                tma_load_1d(t_box + s_off[tid], D.rec_box + src, (uint32_t)n_g * 16u, &s_bar);
                tma_load_1d(t_sn + s_off[tid], D.rec_sn + src, (uint32_t)n_g * 16u, &s_bar);
            }
        }
        for (int p = tid; p < total; p += blockDim.x) {          // class of every position while the copies fly
            int gi = 0;
            while (gi + 1 < ng && p >= s_off[gi + 1]) ++gi;
            s_cat_c[p] = (unsigned char)s_order[gi];
        }
        mbar_wait(&s_bar, 0);
        for (int p = tid; p < total; p += blockDim.x)
            S.skeys[p] = make_key(__int_as_float(t_sn[p].x), (uint32_t)p);
        __syncthreads();
        auto emit = [&](int rank, uint64_t key, const float4& bx) {
            const uint32_t p = (uint32_t)key;
            MLP_BOUND(p, kStageCap);
            MLP_BOUND(rank, max_out);
            const int4 sn = t_sn[p];
            const int c = (int)s_cat_c[p];
            float* o = det_b + rank * 6;
            o[0] = bx.x; o[1] = bx.y; o[2] = bx.z; o[3] = bx.w;
            o[4] = (float)c;
            o[5] = __int_as_float(sn.x);
            if (keep_b) { keep_b[rank * 2] = sn.y; keep_b[rank * 2 + 1] = c; }
        };
        FetchCat f{t_box};
        kept = nms_core<kThreads>(nullptr, total, sort_cap, thr, max_out, S, f, emit, true, 1);
    } else {
    // concatenation, flat over the survivors (one round of loads): position p -> (group gi, rank i)
    for (int p = tid; p < total; p += blockDim.x) {
        int gi = 0;
        while (gi + 1 < ng && p >= s_off[gi + 1]) ++gi;
        const int c = s_order[gi];
        const int64_t src = (int64_t)(b * C + c) * max_out + (p - s_off[gi]);
        MLP_BOUND(p, C * max_out);
        MLP_BOUND(p - s_off[gi], max_out);
        const float4 bx = __ldcg(D.rec_box + src);
        const int4 sn4 = __ldcg(D.rec_sn + src);
        const int2 sn = make_int2(sn4.x, sn4.y);
        cat_box[p] = bx;
        cat_sn[p] = sn;
        cat_c[p] = c;
        const uint64_t key = make_key(__int_as_float(sn.x), (uint32_t)p);
        cat_keys[p] = key;
        if (staged) S.skeys[p] = key;
    }
    __syncthreads();
    auto emit = [&](int rank, uint64_t key, const float4& bx) {
        const uint32_t p = (uint32_t)key;
        MLP_BOUND(p, C * max_out);
        MLP_BOUND(rank, max_out);
        const int2 sn = cat_sn[p];
        const int c = cat_c[p];
        float* o = det_b + rank * 6;
        o[0] = bx.x; o[1] = bx.y; o[2] = bx.z; o[3] = bx.w;
        o[4] = (float)c;
        o[5] = __int_as_float(sn.x);
        if (keep_b) { keep_b[rank * 2] = sn.y; keep_b[rank * 2 + 1] = c; }
    };
    if (total > 0) {
        FetchCat f{cat_box};
        kept = nms_core<kThreads>(cat_keys, total, sort_cap, thr, max_out, S, f, emit, staged, 1);
    }
    }
    NMS_MARK(1, 30);
    // -1 padding of the unused rows (MoldBatch, misc.py:276-283)
    for (int i = kept * 6 + tid; i < max_out * 6; i += blockDim.x) det_b[i] = -1.0f;
    if (keep_b)
        for (int i = kept * 2 + tid; i < max_out * 2; i += blockDim.x) keep_b[i] = -1;
    if (tid == 0) {
        counts[b] = kept;
        if (m_dev) atomicMax(m_dev, kept > 1 ? kept : 1);
    }
    if (!F.enabled) return;
    // ---- fused MaskDistribute + RoIAlign plan for this image
    __shared__ int s_lvl[MLP_MAX_KEEP];
    __syncthreads();                               // det_b rows written by other threads
    for (int r = tid; r < max_out; r += blockDim.x) {
        float* drow = F.dist + ((int64_t)b * max_out + r) * 7;
        int lv = -1;
        if (r < kept) {
            const float* o = det_b + r * 6;
            const float k = level_of(o[0], o[2], o[3], F.base_eps, F.max_k);
            drow[0] = k;
#pragma unroll
            for (int q = 0; q < 6; ++q) drow[q + 1] = o[q];
            lv = (int)k;
        } else {
#pragma unroll
            for (int q = 0; q < 7; ++q) drow[q] = -1.0f;
        }
        s_lvl[r] = lv;
    }
    __syncthreads();
    const int f = tid >> 5, lane = tid & 31;
    if (f < F.L) {
        RoiRec* rec = F.roi_rec + ((int64_t)f * B + b) * max_out;
        int base = 0;
        for (int r0 = 0; r0 < kept; r0 += 32) {
            const int r = r0 + lane;
            const bool hit = (r < kept) && (s_lvl[r] == f);
            const unsigned mask = __ballot_sync(0xffffffffu, hit);
            if (hit) {
                MLP_BOUND(base + __popc(mask & ((1u << lane) - 1u)), max_out);
                RoiRec v;
                v.j = r; v.pad = 0;
#pragma unroll
                for (int q = 0; q < 6; ++q) v.box[q] = det_b[r * 6 + q];
                rec[base + __popc(mask & ((1u << lane) - 1u))] = v;
            }
            base += __popc(mask);
        }
        if (lane == 0) {
            F.level_counts[f * B + b] = base;
            atomicMax(F.level_m + f, base > 1 ? base : 1);
        }
    }
}

// The cross-class pass as a kernel of its own: one CTA per image (used when it cannot ride behind the per-class
// kernel: survivors that do not fit that kernel's shared memory, or MLP_NMS_FUSE=0).
__global__ void __launch_bounds__(kCrossThreads, 1)
nms_cross_class_kernel(int C, float thr, int max_out, int sort_cap, DetScratch D,
                       float* __restrict__ det, int32_t* __restrict__ keep,
                       int32_t* __restrict__ counts, int32_t* __restrict__ m_dev, const FusedPlan F, int use_tma) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cross_class_body<kCrossThreads>(smem_raw, blockIdx.x, gridDim.x, C, thr, max_out, sort_cap, D, det, keep, counts,
                                    m_dev, F, use_tma);
}

// ------------------------------------------------------------ host side ------
struct DetPlan {
    DetScratch D;
    int64_t bytes;
};

inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

int plan_scratch(mlp_ctx* ctx, int B, int64_t N, int C, int max_out, DetScratch* out) {
    const int64_t G = (int64_t)B * C;
    const int64_t cap = align_up(N, 2);
    int64_t off = 0;
    auto take = [&](int64_t bytes) { int64_t o = off; off = align_up(off + bytes, 256); return o; };
    const int64_t o_keys = take(G * cap * 8);
    const int64_t o_count = take(G * 4);
    const int64_t o_done = take((int64_t)B * 4);                  // directly behind cand_count: one memset covers both
    const int64_t o_kept = take(G * 4);
    const int64_t o_minn = take(G * 4);
    const int64_t o_rbox = take(G * max_out * 16);
    const int64_t o_rsn = take(G * max_out * 16);
    const int64_t o_ckeys = take(G * max_out * 8);
    const int64_t o_cbox = take(G * max_out * 16);
    const int64_t o_csn = take(G * max_out * 8);
    const int64_t o_cc = take(G * max_out * 4);
    int rc = mlp_ensure_scratch(ctx, MLP_ARENA_DETECT, off);
    if (rc) return rc;
    char* base = static_cast<char*>(ctx->arena[MLP_ARENA_DETECT]);
    out->cand_keys = reinterpret_cast<uint64_t*>(base + o_keys);
    out->cand_count = reinterpret_cast<int32_t*>(base + o_count);
    out->img_done = reinterpret_cast<int32_t*>(base + o_done);
    out->zero_bytes = o_done + (int64_t)B * 4 - o_count;
    out->cls_kept = reinterpret_cast<int32_t*>(base + o_kept);
    out->min_n = reinterpret_cast<int32_t*>(base + o_minn);
    out->rec_box = reinterpret_cast<float4*>(base + o_rbox);
    out->rec_sn = reinterpret_cast<int4*>(base + o_rsn);
    out->cat_keys = reinterpret_cast<uint64_t*>(base + o_ckeys);
    out->cat_box = reinterpret_cast<float4*>(base + o_cbox);
    out->cat_sn = reinterpret_cast<int2*>(base + o_csn);
    out->cat_c = reinterpret_cast<int32_t*>(base + o_cc);
    out->cap = cap;
    return MLP_OK;
}

int pick_sort_cap(int want_items, int max_out, size_t smem_limit) {
    int cap = 256;
    while (cap < want_items && cap < 16384 && nms_smem_bytes(cap * 2, max_out) <= smem_limit) cap <<= 1;
    return cap;
}

int detection_impl(mlp_ctx* ctx, const mlp_prior_config* prior, int height, int width,
                   const float* cls_dev, const float* boxes_or_loc_dev, int B, int64_t N, int C,
                   const mlp_detection_params* p, float* det_dev, int32_t* keep_dev,
                   int32_t* counts_dev, int32_t* m_dev, cudaStream_t stream, const char* who,
                   const FusedPlan& fp) {
    MLP_CHECK_ARG(ctx && cls_dev && boxes_or_loc_dev && p && det_dev && counts_dev,
                  "%s: NULL argument", who);
    MLP_CHECK_ARG(B >= 1 && N >= 1 && C >= 1, "%s: bad shape B=%d N=%lld C=%d", who, B, (long long)N, C);
    MLP_CHECK_ARG(C <= 256, "%s: num_classes=%d > 256 not supported", who, C);
    MLP_CHECK_ARG(N < (1ll << 30) && (int64_t)B * N * C < (1ll << 32) - 4096,
                  "%s: B*N*C must fit 32 bits", who);
    if (p->strict_batch && B > MLP_MAX_BATCH) {
        mlp_set_error("%s: batch %d > 32; the reference's MoldBatch uses tf.dynamic_partition(.., 32) "
                      "(engine/layers/misc.py:275). Pass strict_batch=0 to lift the limit.", who, B);
        return MLP_EBATCH;
    }
    MLP_CHECK_ARG(p->nms_max_output_size >= 1 && p->nms_max_output_size <= MLP_MAX_KEEP,
                  "%s: nms_max_output_size=%d out of range [1,%d]", who, p->nms_max_output_size,
                  MLP_MAX_KEEP);
    MLP_CHECK_ARG(mlp_aligned16(cls_dev) && mlp_aligned16(boxes_or_loc_dev) && mlp_aligned16(det_dev),
                  "%s: pointers must be 16-byte aligned", who);
    PriorDev P;
    memset(&P, 0, sizeof(P));
    if (prior) {
        int rc = mlp_build_prior_dev(prior, height, width, &P);
        if (rc) return rc;
        MLP_CHECK_ARG(P.total == N, "%s: prior config gives %d anchors, tensors have %lld", who, P.total,
                      (long long)N);
    }
    DeviceGuard g(ctx->device);
    const int max_out = p->nms_max_output_size;
    DetScratch D;
    int rc = plan_scratch(ctx, B, N, C, max_out, &D);
    if (rc) return rc;
    const int G = B * C;
    MLP_CUDA(cudaMemsetAsync(D.cand_count, 0, (size_t)D.zero_bytes, stream));

    // K1: stream cls_pred once
    {
        ProfScope prof(ctx, MLP_ST_THRESHOLD, stream);
        const int64_t total = (int64_t)B * N * C;
        int64_t tiles = (total / 4 + kK1Threads * 4 - 1) / (kK1Threads * 4);
        int occ = 0;
        MLP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, threshold_compact_kernel, kK1Threads, 0));
        int64_t capb = (int64_t)ctx->sm_count * (occ < 1 ? 1 : occ);      // one resident wave
        if (const char* e = getenv("MLP_K1_TILES_PER_CTA")) {             // tuning knob
            const int t = atoi(e);
            if (t > 0) capb = (tiles + t - 1) / t;
        }
        int grid = (int)(tiles < capb ? (tiles < 1 ? 1 : tiles) : capb);
        threshold_compact_kernel<<<grid, kK1Threads, 0, stream>>>(cls_dev, (uint32_t)total, (int)N, C,
                                                          p->min_confidence, D.cand_keys,
                                                          D.cand_count, D.cap, m_dev,
                                                          fp.enabled ? fp.level_m : nullptr,
                                                          fp.enabled ? fp.L + 1 : 0);
        MLP_LAUNCH_CHECK(ctx);
    }
    // K2: per (image,class) NMS.  512-thread CTAs, <= 74 KB smem -> 2-3 CTAs per SM, so all
    // B*C groups (160-192 at batch 32) are resident in a single wave on 148 SMs.  When an image's survivors
    // (<= C * max_out) fit the same shared memory, the cross-class pass rides behind it (CrossArgs).
    const int sort_cap2 = pick_sort_cap(4096, max_out, 74 * 1024);
    const size_t smem2 = nms_smem_bytes(sort_cap2, max_out);
    int cross_cap = 256;
    while (cross_cap < C * max_out) cross_cap <<= 1;
    // Measured (profiles/nms_fuse_ab_r02.txt): eager launches save 3 us per batch, but inside a CUDA graph - where a
    // kernel boundary costs next to nothing - the fused form is no faster (425 vs 422 us per batch alone at cfg-2,
    // 108.9 vs 109.2 us at batch 1) because the cross-class pass then runs on 512 threads instead of 1024.  Off unless
    // MLP_NMS_FUSE=1; both forms are covered by the GPU tests.
    bool fuse = false;
    if (const char* e = getenv("MLP_NMS_FUSE")) fuse = atoi(e) != 0;
    fuse = fuse && nms_smem_bytes(cross_cap, max_out) <= smem2 && C <= 256;
    {
        ProfScope prof(ctx, MLP_ST_NMS_CLASS, stream);
        CrossArgs X;
        memset(&X, 0, sizeof(X));
        X.fuse = fuse ? 1 : 0;
        X.thr = p->post_iou_threshold;
        X.sort_cap = cross_cap;
        X.det = det_dev; X.keep = keep_dev; X.counts = counts_dev; X.m_dev = m_dev;
        X.img_done = D.img_done;
        X.F = fp;
        if (prior) {
            MLP_CUDA(cudaFuncSetAttribute(nms_per_class_kernel<true>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
            nms_per_class_kernel<true><<<G, kClassThreads, smem2, stream>>>(
                P, reinterpret_cast<const float4*>(boxes_or_loc_dev), (int)N, C, p->nms_iou_threshold,
                max_out, sort_cap2, D, X);
        } else {
            MLP_CUDA(cudaFuncSetAttribute(nms_per_class_kernel<false>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
            nms_per_class_kernel<false><<<G, kClassThreads, smem2, stream>>>(
                P, reinterpret_cast<const float4*>(boxes_or_loc_dev), (int)N, C, p->nms_iou_threshold,
                max_out, sort_cap2, D, X);
        }
        MLP_LAUNCH_CHECK(ctx);
    }
    // K3: cross-class NMS per image as its own kernel when it could not be fused; sort buffer sized for
    // C*max_out survivors when it fits.
    if (!fuse) {
        ProfScope prof(ctx, MLP_ST_NMS_CROSS, stream);
        const int sort_cap = pick_sort_cap(C * max_out, max_out, 180 * 1024);
        const size_t smem = ((nms_smem_bytes(sort_cap, max_out) + 127) & ~(size_t)127) + (size_t)kStageCap * 32;
        int use_tma = 0;                               // measured: equal at cfg-2, 1.7 us slower at batch 1 (profiles/nms_tma_ab_r02.txt)
        if (const char* e = getenv("MLP_NMS_TMA")) use_tma = atoi(e);
        MLP_CUDA(cudaFuncSetAttribute(nms_cross_class_kernel,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        nms_cross_class_kernel<<<B, kCrossThreads, smem, stream>>>(C, p->post_iou_threshold, max_out,
                                                                sort_cap, D, det_dev, keep_dev,
                                                                counts_dev, m_dev, fp, use_tma);
        MLP_LAUNCH_CHECK(ctx);
    }
    return MLP_OK;
}

}  // namespace

extern "C" int mlp_detection_proposal(mlp_ctx* ctx, const float* cls_dev, const float* boxes_dev,
                                      int batch, int64_t num_boxes, int num_classes,
                                      const mlp_detection_params* params, float* det_dev,
                                      int32_t* keep_dev, int32_t* counts_dev, int32_t* m_dev,
                                      mlp_stream_t stream) {
    FusedPlan none;
    memset(&none, 0, sizeof(none));
    return detection_impl(ctx, nullptr, 0, 0, cls_dev, boxes_dev, batch, num_boxes, num_classes, params,
                          det_dev, keep_dev, counts_dev, m_dev, (cudaStream_t)stream,
                          "mlp_detection_proposal", none);
}

extern "C" int mlp_detect_from_heads(mlp_ctx* ctx, const mlp_prior_config* prior, const float* loc_dev,
                                     const float* cls_dev, int batch, int height, int width,
                                     int num_classes, const mlp_detection_params* params,
                                     float* det_dev, int32_t* keep_dev, int32_t* counts_dev,
                                     int32_t* m_dev, mlp_stream_t stream) {
    if (!prior) {
        mlp_set_error("mlp_detect_from_heads: prior config is NULL");
        return MLP_EINVAL;
    }
    int64_t N = mlp_prior_count(prior, height, width);
    if (N < 0) return (int)N;
    FusedPlan none;
    memset(&none, 0, sizeof(none));
    return detection_impl(ctx, prior, height, width, cls_dev, loc_dev, batch, N, num_classes, params,
                          det_dev, keep_dev, counts_dev, m_dev, (cudaStream_t)stream,
                          "mlp_detect_from_heads", none);
}

// Fused first half of the path (a2-a10): detection with MaskDistribute and the RoIAlign plan folded
// into the cross-class NMS epilogue (mlp_detect_plan, 3 kernels), then the RoIAlign run.
extern "C" int mlp_detect_plan(mlp_ctx* ctx, const mlp_prior_config* prior, const float* loc_dev,
                               const float* cls_dev, int batch, int height, int width, int num_classes,
                               const mlp_detection_params* params, int max_k, float base_size,
                               float* det_dev, int32_t* keep_dev, int32_t* counts_dev, int32_t* m_dev,
                               float* dist_dev, int32_t* level_counts_dev, int32_t* level_m_dev,
                               mlp_stream_t stream) {
    if (!ctx || !prior || !params || !dist_dev || !level_counts_dev || !level_m_dev) {
        mlp_set_error("mlp_detect_plan: NULL argument");
        return MLP_EINVAL;
    }
    MLP_CHECK_ARG(max_k >= 0 && max_k + 1 <= MLP_MAX_LEVELS, "mlp_detect_plan: max_k=%d out of range", max_k);
    MLP_CHECK_ARG(params->nms_max_output_size >= 1 && params->nms_max_output_size <= MLP_MAX_KEEP,
                  "mlp_detect_plan: nms_max_output_size out of range");
    int64_t N = mlp_prior_count(prior, height, width);
    if (N < 0) return (int)N;
    const int L = max_k + 1, K = params->nms_max_output_size;
    {
        DeviceGuard g(ctx->device);
        int rc = mlp_ensure_scratch(ctx, MLP_ARENA_ROI, (int64_t)L * batch * K * (int64_t)sizeof(RoiRec));
        if (rc) return rc;
    }
    FusedPlan fp;
    fp.enabled = 1;
    fp.L = L;
    fp.base_eps = (float)((double)base_size + 1e-7);
    fp.max_k = (float)max_k;
    fp.dist = dist_dev;
    fp.roi_rec = static_cast<RoiRec*>(ctx->arena[MLP_ARENA_ROI]);
    fp.level_counts = level_counts_dev;
    fp.level_m = level_m_dev;
    return detection_impl(ctx, prior, height, width, cls_dev, loc_dev, batch, N, num_classes, params,
                          det_dev, keep_dev, counts_dev, m_dev, (cudaStream_t)stream, "mlp_detect_plan", fp);
}

extern "C" int mlp_detect_align(mlp_ctx* ctx, const mlp_prior_config* prior, const float* loc_dev,
                                const float* cls_dev, int batch, int height, int width, int num_classes,
                                const mlp_detection_params* params, int max_k, float base_size,
                                const float* const* fmaps_dev, const int32_t* fh, const int32_t* fw,
                                int channels, int crop_h, int crop_w, float* det_dev, int32_t* keep_dev,
                                int32_t* counts_dev, int32_t* m_dev, float* dist_dev,
                                int32_t* level_counts_dev, int32_t* level_m_dev,
                                float* const* crops_dev, float* roi_boxes_dev, mlp_stream_t stream) {
    int rc = mlp_detect_plan(ctx, prior, loc_dev, cls_dev, batch, height, width, num_classes, params, max_k,
                             base_size, det_dev, keep_dev, counts_dev, m_dev, dist_dev, level_counts_dev,
                             level_m_dev, stream);
    if (rc) return rc;
    const int L = max_k + 1, K = params->nms_max_output_size;
    return mlp_roi_align_run(ctx, fmaps_dev, fh, fw, L, channels, dist_dev, batch, K, K, m_dev,
                             (float)height, (float)width, crop_h, crop_w, level_counts_dev, level_m_dev,
                             crops_dev, roi_boxes_dev, stream);
}
