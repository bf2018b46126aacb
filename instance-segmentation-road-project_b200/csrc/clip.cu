// clip.cu — the pasted binary masks in box-clipped form: CropAndPadMask (+ the consumers' > 0.5) without the
// zeros.  97 % of the [B,M,PH,PW] tensor of /root/reference/engine/layers/misc.py:358-401 is padding that the
// reference adds with tf.pad (:393-394) around each resized mask; a client on the far side of PCIe (or NCCL) only
// needs the resized part and where it goes.  Per instance of the [B,K] capacity grid this writes a 32-byte record
//     { xmin, ymin, w, h, offset_lo, offset_hi, class, conf }        (int32 x 8)
// and h rows of ceil(w / 8) bytes at `offset` in a byte pool: bit k of byte i of row r = pixel
// (ymin + r, xmin + 8 i + k) of the instance's frame-sized mask, (paste value > 0.5).  w = h = 0: nothing is
// pasted (MoldBatch padding row, filtered by the confidence rule, box clipped away).  The dense tensor is
// recovered exactly by zero-filling and OR-ing the rows in (masklab_b200.expand_clipped, tests/test_gpu_clip.py).
// cfg-2: 2.4 MB instead of 1.68 GB (uint8) / 210 MB (bit-packed frames).
//
// Runs behind mlp_trim_paste(.., MLP_PASTE_NONE, ..), which prepares the fused tail (int boxes, slot -> RoI table,
// bit tiles, M and the row-filter threshold).
#include "paste_common.cuh"

namespace {

constexpr int kClipThreads = 256;
constexpr int kClipBytes = 2048;           // bytes (= 16 K pixels) per work item

struct ClipArgs {
    const int32_t* det;        // [B,K,6]
    PasteSrc src;
    int B, K, mh, mw, PH, PW;
    int32_t* geom;             // [B,K,8]
    uint8_t* pool;
    long long pool_cap;
    unsigned long long* used;  // [2]: bytes, work items
    uint2* items;              // x = b*K + j, y = (first row inside the box << 16) | rows
    int item_cap;
};

// One CTA per image, one thread per slot (rounds of 256): geometry, byte and item counts, a block-wide exclusive scan
// (order j), one atomicAdd per round on the pool / item counters.
__global__ void __launch_bounds__(kClipThreads)
clip_plan_kernel(const ClipArgs A) {
    __shared__ unsigned long long s_bytes[kClipThreads / 32];
    __shared__ int s_items[kClipThreads / 32];
    __shared__ unsigned long long s_base_b;
    __shared__ int s_base_i;
    int M, thr;
    paste_scalars(A.src, A.B, A.K, M, thr);
    const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int j0 = 0; j0 < A.K; j0 += kClipThreads) {
        const int j = j0 + tid;
        int row[6] = {0, 0, 0, 0, -1, -100};
        PasteGeom g;
        g.xmin = g.xmax = g.ymin = g.ymax = 0; g.sx = g.sy = 0.f; g.active = false;
        if (j < M) {
#pragma unroll
            for (int q = 0; q < 6; ++q) row[q] = A.det[((int64_t)b * A.K + j) * 6 + q];
            g = paste_geometry(row, thr, A.mh, A.mw, A.PH, A.PW);
        }
        const int w = g.active ? g.xmax - g.xmin : 0, h = g.active ? g.ymax - g.ymin : 0;
        const int rb = (w + 7) >> 3;
        const unsigned long long bytes = (unsigned long long)rb * h;
        const int rows_per = rb > 0 ? max(1, kClipBytes / rb) : 1;
        const int nitems = h > 0 ? (h + rows_per - 1) / rows_per : 0;
        // exclusive scan over the CTA (warp shuffles + 8 warp totals)
        unsigned long long ib = bytes;
        int ii = nitems;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long tb = __shfl_up_sync(0xffffffffu, ib, d);
            const int ti = __shfl_up_sync(0xffffffffu, ii, d);
            if (lane >= d) { ib += tb; ii += ti; }
        }
        if (lane == 31) { s_bytes[warp] = ib; s_items[warp] = ii; }
        __syncthreads();
        unsigned long long pre_b = 0, tot_b = 0;
        int pre_i = 0, tot_i = 0;
#pragma unroll
        for (int q = 0; q < kClipThreads / 32; ++q) {
            if (q < warp) { pre_b += s_bytes[q]; pre_i += s_items[q]; }
            tot_b += s_bytes[q]; tot_i += s_items[q];
        }
        if (tid == 0) {
            s_base_b = atomicAdd(A.used, tot_b);
            s_base_i = (int)atomicAdd(A.used + 1, (unsigned long long)tot_i);
        }
        __syncthreads();
        const unsigned long long off = s_base_b + pre_b + ib - bytes;
        int it = s_base_i + pre_i + ii - nitems;
        if (j < A.K) {
            int32_t* o = A.geom + ((int64_t)b * A.K + j) * 8;
            o[0] = g.xmin; o[1] = g.ymin; o[2] = w; o[3] = h;
            o[4] = (int32_t)(uint32_t)off; o[5] = (int32_t)(uint32_t)(off >> 32);
            o[6] = row[4]; o[7] = row[5];
            for (int r0 = 0; r0 < h; r0 += rows_per, ++it) {
                MLP_BOUND(it, A.item_cap);
                if (it < A.item_cap)
                    A.items[it] = make_uint2((uint32_t)(b * A.K + j), ((uint32_t)r0 << 16) | (uint32_t)min(rows_per, h - r0));
            }
        }
        __syncthreads();
    }
}

// Grid-stride over the work items: the instance's tile and column terms in shared memory, one thread per output byte.
__global__ void __launch_bounds__(kClipThreads)
clip_rows_kernel(const ClipArgs A) {
    extern __shared__ __align__(16) unsigned char s_dyn[];
    float* s_tile = reinterpret_cast<float*>(s_dyn);
    uint2* s_col = reinterpret_cast<uint2*>(s_dyn + ((A.mh * A.mw * 4 + 15) & ~15));
    int M, thr;
    paste_scalars(A.src, A.B, A.K, M, thr);
    const unsigned long long n64 = A.used[1];
    const int n = (int)(n64 < (unsigned long long)A.item_cap ? n64 : (unsigned long long)A.item_cap);
    const int tid = threadIdx.x;
    for (int it = blockIdx.x; it < n; it += gridDim.x) {
        const uint2 q = __ldg(A.items + it);
        const int b = (int)q.x / A.K, j = (int)q.x - b * A.K;
        const int r0 = (int)(q.y >> 16), nrows = (int)(q.y & 0xffffu);
        const int32_t* gm = A.geom + (int64_t)q.x * 8;
        const int xmin = gm[0], ymin = gm[1], w = gm[2], h = gm[3];
        const unsigned long long off = (unsigned long long)(uint32_t)gm[4] | ((unsigned long long)(uint32_t)gm[5] << 32);
        int row[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) row[k] = A.det[(int64_t)q.x * 6 + k];
        const PasteGeom g = paste_geometry(row, thr, A.mh, A.mw, A.PH, A.PW);   // same box: the resize scales
        const TileRef tref = tile_ref(A.src, b, j, A.K, A.mh * A.mw, row[4], A.mh, A.mw);
        const int rb = (w + 7) >> 3;
        __syncthreads();                                   // previous item done with shared memory
        tref.fill(s_tile, A.mh, tid, kClipThreads);
        const bool cols = w <= kMaxCols;
        if (cols)                                          // box-aligned column terms (paste_value's x half), order c
            for (int c = tid; c < w; c += kClipThreads) {
                const float p = __fmul_rn((float)c, g.sx);
                const float fl = floorf(p);
                const int xlo = max((int)fl, 0), xhi = min((int)ceilf(p), A.mw - 1);
                s_col[c] = make_uint2((uint32_t)(xlo * 4) | ((uint32_t)(xhi * 4) << 16), __float_as_uint(__fsub_rn(p, fl)));
            }
        __syncthreads();
        (void)xmin; (void)ymin; (void)h;
        const int nb = nrows * rb;
        for (int i = tid; i < nb; i += kClipThreads) {
            const int r = i / rb, cb = i - r * rb;
            const int oy = g.ymin + r0 + r;
            const float p = __fmul_rn((float)(oy - g.ymin), g.sy);
            const float fl = floorf(p);
            const int ylo = max((int)fl, 0), yhi = min((int)ceilf(p), A.mh - 1);
            const float ly = __fsub_rn(p, fl);
            const unsigned char* row_lo = reinterpret_cast<const unsigned char*>(s_tile + ylo * A.mw);
            const unsigned char* row_hi = reinterpret_cast<const unsigned char*>(s_tile + yhi * A.mw);
            unsigned byte = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int c = cb * 8 + k;
                if (c < w) {
                    const float v = cols ? paste_value_cols(row_lo, row_hi, ly, s_col[c])
                                         : paste_value(s_tile, A.mh, A.mw, ylo, yhi, ly, c, g.sx);
                    byte |= (unsigned)(v > 0.5f) << k;
                }
            }
            const unsigned long long at = off + (unsigned long long)(r0 + r) * rb + cb;
            if ((long long)at < A.pool_cap) A.pool[at] = (uint8_t)byte;
        }
    }
}

}  // namespace

extern "C" int64_t mlp_clip_pool_bound(int batch, int k_rows, int frame_h, int frame_w) {
    return (int64_t)batch * k_rows * frame_h * ((frame_w + 7) / 8);
}

extern "C" int mlp_clip_masks(mlp_ctx* ctx, const int32_t* det_i32_dev, const float* roi_masks_dev, int r_rows,
                              const int32_t* r_dev, int num_classes, const int32_t* counts_dev, int batch, int k_rows,
                              int mask_h, int mask_w, int frame_h, int frame_w, int32_t* geom_dev, uint8_t* pool_dev,
                              int64_t pool_capacity, int64_t* used_dev, mlp_stream_t stream) {
    MLP_CHECK_ARG(ctx && det_i32_dev && roi_masks_dev && counts_dev && geom_dev && pool_dev && used_dev,
                  "mlp_clip_masks: NULL argument");
    MLP_CHECK_ARG(batch >= 1 && k_rows >= 1 && num_classes >= 1 && r_rows >= 1, "mlp_clip_masks: bad shape");
    MLP_CHECK_ARG(mask_h >= 1 && mask_w >= 1 && mask_h * mask_w <= kMaxTile, "mlp_clip_masks: mask tile %dx%d", mask_h,
                  mask_w);
    MLP_CHECK_ARG(frame_h >= 1 && frame_w >= 1 && frame_h < 65536, "mlp_clip_masks: frame %dx%d", frame_h, frame_w);
    MLP_CHECK_ARG(pool_capacity >= 0, "mlp_clip_masks: negative pool capacity");
    MLP_CHECK_ARG(mlp_aligned16(used_dev) && mlp_aligned16(geom_dev), "mlp_clip_masks: geom_dev / used_dev must be 16-byte aligned");
    const FusedTail need = fused_tail_layout(nullptr, batch, k_rows, mask_h, mask_w);
    MLP_CHECK_ARG(ctx->arena[MLP_ARENA_FUSED] && ctx->arena_bytes[MLP_ARENA_FUSED] >= need.bytes,
                  "mlp_clip_masks: call mlp_trim_paste (MLP_PASTE_NONE) with the same shapes first");
    DeviceGuard g(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    const FusedTail ft = fused_tail_layout(ctx->arena[MLP_ARENA_FUSED], batch, k_rows, mask_h, mask_w);
    // work items: a chunk holds at least max(1, kClipBytes / bytes per frame row) rows
    const int rbmax = (frame_w + 7) / 8;
    const int rows_min = kClipBytes / rbmax > 1 ? kClipBytes / rbmax : 1;
    const int64_t item_cap = (int64_t)batch * k_rows * ((frame_h + rows_min - 1) / rows_min);
    MLP_CHECK_ARG(item_cap < (1ll << 28), "mlp_clip_masks: work-item list too large");
    int rc = mlp_ensure_scratch(ctx, MLP_ARENA_PASTE, 256 + item_cap * (int64_t)sizeof(uint2));
    if (rc) return rc;
    ClipArgs A;
    memset(&A, 0, sizeof(A));
    A.det = det_i32_dev;
    A.src.fused = 1;
    A.src.roi_masks = roi_masks_dev;
    A.src.tail_src = ft.tail_src;
    A.src.tail_bits = ft.tail_bits;
    A.src.r_dev = r_dev;
    A.src.r_rows = r_rows;
    A.src.C = num_classes;
    A.src.planar = ctx->tail_planar;
    A.src.counts = counts_dev;
    A.src.confmax = ft.confmax;
    A.src.scalars = ft.scalars;
    A.B = batch; A.K = k_rows; A.mh = mask_h; A.mw = mask_w; A.PH = frame_h; A.PW = frame_w;
    A.geom = geom_dev;
    A.pool = pool_dev;
    A.pool_cap = pool_capacity;
    A.used = reinterpret_cast<unsigned long long*>(used_dev);
    A.items = reinterpret_cast<uint2*>(static_cast<char*>(ctx->arena[MLP_ARENA_PASTE]) + 256);
    A.item_cap = (int)item_cap;
    ProfScope prof(ctx, MLP_ST_PASTE, st);
    MLP_CUDA(cudaMemsetAsync(used_dev, 0, 16, st));
    clip_plan_kernel<<<batch, kClipThreads, 0, st>>>(A);
    MLP_LAUNCH_CHECK(ctx);
    const size_t smem = (size_t)((mask_h * mask_w * 4 + 15) & ~15) + (size_t)kMaxCols * sizeof(uint2);
    clip_rows_kernel<<<ctx->sm_count * 8, kClipThreads, smem, st>>>(A);
    MLP_LAUNCH_CHECK(ctx);
    return MLP_OK;
}
