// jpeg.cu — the JPEG encode at the end of the serving graph ("overlay + JPEG encode"):
// EncodeImageContent.call (/root/reference/engine/layers/misc.py:343-351) = tf.io.encode_jpeg(image) with every
// attribute at its default, wired in /root/reference/road_project/setup/serving.py:41.  TensorFlow's kernel drives
// libjpeg(-turbo): quality 95 (force_baseline), RGB -> YCbCr 4:2:0, JDCT_ISLOW, the Annex K Huffman tables,
// baseline sequential, JFIF 300x300 dpi.  All of it is integer work, so the byte stream is reproduced EXACTLY
// (oracle/jpeg_oracle.py restates it and is pinned byte for byte against libjpeg-turbo's own output).
//
// Six launches per batch of frames, everything stays on the device:
//   jpeg_dct_kernel     one WARP per pair of MCUs, four pairs in a row per warp (the next strip is in flight while this
//                       one is transformed; no CTA barrier inside the loop): colour conversion (jccolor.c fixed point), h2v2 chroma
//                       down-sampling with the alternating 1,2 bias (jcsample.c), edge replication and dummy blocks
//                       (jcprepct.c / jccoefct.c), jpeg_fdct_islow in registers with conflict-free shared-memory
//                       transposes (jfdctint.c), quantisation by exact reciprocal multiplication (jcdctmgr.c); writes
//                       zigzag int16 coefficients and the DC of every block
//   jpeg_enc_kernel     one THREAD per 8x8 block builds the 64-bit non-zero mask of its coefficients (two per
//                       comparison) and walks its set bits (jchuff.c encode_one_block): DC
//                       difference, ZRL / run-size codes and value bits go, left-aligned, into the block's private
//                       slot; the bit lengths are scanned inside the CTA (32 MCUs), one total per CTA
//   jpeg_offsets_kernel per frame: scan of the CTA totals -> bit offset of every CTA, and zero-fill of exactly the
//                       words the scan will occupy (spread over 16 CTAs per frame), 1-padding of the last byte
//   jpeg_place_kernel   one thread per block funnel-shifts its slot to the block's bit offset in the scan (only the
//                       first and last word of a block are atomics - neighbours share them)
//   jpeg_ffcount_kernel 0xFF bytes per 4 KB chunk of the scan (byte stuffing moves every later byte)
//   jpeg_stuff_kernel   header (jcmarker.c), scan bytes with a 0x00 after every 0xFF, EOI, length
// The coefficient passes are integer-issue bound, not HBM bound (3 B/px in, 3 B/px of coefficients out and back).
#include <stdarg.h>

#include "common.cuh"

namespace {

// ------------------------------------------------------------------------------------------------- tables
const uint8_t kStdLumaQ[64] = {16, 11, 10, 16, 24,  40,  51,  61,  12, 12, 14, 19, 26,  58,  60,  55,
                               14, 13, 16, 24, 40,  57,  69,  56,  14, 17, 22, 29, 51,  87,  80,  62,
                               18, 22, 37, 56, 68,  109, 103, 77,  24, 35, 55, 64, 81,  104, 113, 92,
                               49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99};
const uint8_t kStdChromaQ[64] = {17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99,
                                 99, 99, 47, 66, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99,
                                 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99};
const uint8_t kDcLumaBits[16] = {0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0};
const uint8_t kDcChromaBits[16] = {0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0};
const uint8_t kDcVals[12] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11};
const uint8_t kAcLumaBits[16] = {0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d};
const uint8_t kAcLumaVals[162] = {
    0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71,
    0x14, 0x32, 0x81, 0x91, 0xa1, 0x08, 0x23, 0x42, 0xb1, 0xc1, 0x15, 0x52, 0xd1, 0xf0, 0x24, 0x33, 0x62, 0x72,
    0x82, 0x09, 0x0a, 0x16, 0x17, 0x18, 0x19, 0x1a, 0x25, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x34, 0x35, 0x36, 0x37,
    0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59,
    0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x83,
    0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3,
    0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3,
    0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe1, 0xe2,
    0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa};
const uint8_t kAcChromaBits[16] = {0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77};
const uint8_t kAcChromaVals[162] = {
    0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22,
    0x32, 0x81, 0x08, 0x14, 0x42, 0x91, 0xa1, 0xb1, 0xc1, 0x09, 0x23, 0x33, 0x52, 0xf0, 0x15, 0x62, 0x72, 0xd1,
    0x0a, 0x16, 0x24, 0x34, 0xe1, 0x25, 0xf1, 0x17, 0x18, 0x19, 0x1a, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x35, 0x36,
    0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58,
    0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a,
    0x82, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a,
    0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba,
    0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda,
    0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa};

constexpr int kHeaderBytes = 623;
constexpr int kMaxBlockWords = 52;            // 20 (DC) + 63 * 26 (AC) = 1658 bits
constexpr int kChunkBytes = 4096;             // byte-stuffing chunk of the scan
constexpr int kPartBlocks = 192;              // blocks per CTA of the entropy pass: 32 whole MCUs

// Everything a launch needs that depends only on (H, W, quality): passed by value (kernel parameter space).
struct JpegTables {
    uint16_t div[2][64];       // natural order, q << 3 (jcdctmgr.c start_pass_fdctmgr)
    uint32_t rcp[2][64];       // ceil(2^32 / div): n / div == umulhi(n, rcp) for every n < 2^16
    uint8_t izz[64];           // natural index -> zigzag position
    uint32_t ac[2][256];       // symbol -> length << 16 | code (jchuff.c jpeg_make_c_derived_tbl)
    uint32_t dc[2][12];
};
struct JpegHeader {
    uint8_t bytes[kHeaderBytes + 1];
};

struct JpegGeom {
    int H, W;
    int mcu_rows, mcu_cols, n_mcu, n_blk;
    int y_blk_rows, y_blk_cols;     // real 8x8 blocks of the luma component
    int c_real_rows;                // chroma rows that come from real pixel rows: ceil(H / 2)
    int64_t words_cap;              // 32-bit words reserved for one frame's unstuffed scan
    int chunks_cap;                 // 4 KB chunks of that reservation
    int parts;                      // groups of kPartBlocks blocks per frame (two-level scan of the bit lengths)
};

void zigzag_natural(uint8_t* nat_of_zz) {     // jutils.c jpeg_natural_order
    int n = 0;
    for (int s = 0; s < 15; ++s) {
        for (int i = 0; i < 8; ++i) {
            // odd diagonals run top-right -> bottom-left (row ascending), even ones the other way
            const int r = (s & 1) ? i : 7 - i;
            const int c = s - r;
            if (c < 0 || c > 7) continue;
            nat_of_zz[n++] = (uint8_t)(r * 8 + c);
        }
    }
}

void derive_huff(const uint8_t* bits, const uint8_t* vals, uint32_t* table) {
    uint32_t code = 0;
    int k = 0;
    for (int len = 1; len <= 16; ++len) {
        for (int i = 0; i < bits[len - 1]; ++i) table[vals[k++]] = ((uint32_t)len << 16) | code++;
        code <<= 1;
    }
}

void quality_table(const uint8_t* std_tbl, int quality, int* out) {   // jcparam.c jpeg_set_quality(force_baseline)
    quality = quality < 1 ? 1 : (quality > 100 ? 100 : quality);
    const int scale = quality < 50 ? 5000 / quality : 200 - quality * 2;
    for (int i = 0; i < 64; ++i) {
        int v = (std_tbl[i] * scale + 50) / 100;
        out[i] = v < 1 ? 1 : (v > 255 ? 255 : v);
    }
}

void build_tables(int H, int W, int quality, JpegTables* T, JpegHeader* hdr) {
    uint8_t nat[64];
    zigzag_natural(nat);
    int q[2][64];
    quality_table(kStdLumaQ, quality, q[0]);
    quality_table(kStdChromaQ, quality, q[1]);
    memset(T, 0, sizeof(*T));
    for (int z = 0; z < 64; ++z) T->izz[nat[z]] = (uint8_t)z;
    for (int t = 0; t < 2; ++t)
        for (int i = 0; i < 64; ++i) {
            T->div[t][i] = (uint16_t)(q[t][i] << 3);
            T->rcp[t][i] = (uint32_t)((((uint64_t)1 << 32) + T->div[t][i] - 1) / T->div[t][i]);
        }
    derive_huff(kDcLumaBits, kDcVals, T->dc[0]);
    derive_huff(kDcChromaBits, kDcVals, T->dc[1]);
    derive_huff(kAcLumaBits, kAcLumaVals, T->ac[0]);
    derive_huff(kAcChromaBits, kAcChromaVals, T->ac[1]);

    // jcmarker.c: SOI, APP0 (JFIF 1.01, density unit 1 = inch, 300 x 300: tf.io.encode_jpeg's defaults), DQT x2,
    // SOF0 (8 bit, 3 components, 2x2 / 1x1 / 1x1), DHT x4, SOS
    uint8_t* p = hdr->bytes;
    auto put = [&](int v) { *p++ = (uint8_t)v; };
    auto put16 = [&](int v) { put(v >> 8); put(v & 255); };
    put(0xFF); put(0xD8);
    put(0xFF); put(0xE0); put16(16); put('J'); put('F'); put('I'); put('F'); put(0); put(1); put(1); put(1);
    put16(300); put16(300); put(0); put(0);
    for (int t = 0; t < 2; ++t) {
        put(0xFF); put(0xDB); put16(67); put(t);
        for (int z = 0; z < 64; ++z) put(q[t][nat[z]]);
    }
    put(0xFF); put(0xC0); put16(17); put(8); put16(H); put16(W); put(3);
    put(1); put(0x22); put(0); put(2); put(0x11); put(1); put(3); put(0x11); put(1);
    const struct { int id; const uint8_t* bits; const uint8_t* vals; int n; } dht[4] = {
        {0x00, kDcLumaBits, kDcVals, 12}, {0x10, kAcLumaBits, kAcLumaVals, 162},
        {0x01, kDcChromaBits, kDcVals, 12}, {0x11, kAcChromaBits, kAcChromaVals, 162}};
    for (const auto& d : dht) {
        put(0xFF); put(0xC4); put16(2 + 1 + 16 + d.n); put(d.id);
        for (int i = 0; i < 16; ++i) put(d.bits[i]);
        for (int i = 0; i < d.n; ++i) put(d.vals[i]);
    }
    put(0xFF); put(0xDA); put16(12); put(3); put(1); put(0x00); put(2); put(0x11); put(3); put(0x11);
    put(0); put(63); put(0);
}

JpegGeom make_geom(int H, int W) {
    JpegGeom g;
    g.H = H; g.W = W;
    g.mcu_rows = (H + 15) / 16; g.mcu_cols = (W + 15) / 16;
    g.n_mcu = g.mcu_rows * g.mcu_cols; g.n_blk = g.n_mcu * 6;
    g.y_blk_rows = (H + 7) / 8; g.y_blk_cols = (W + 7) / 8;
    g.c_real_rows = (H + 1) / 2;
    g.words_cap = (int64_t)g.n_blk * kMaxBlockWords + 4;
    g.words_cap = (g.words_cap + 3) & ~(int64_t)3;
    g.chunks_cap = (int)((g.words_cap * 4 + kChunkBytes - 1) / kChunkBytes);
    g.parts = (g.n_blk + kPartBlocks - 1) / kPartBlocks;
    return g;
}

// ------------------------------------------------------------------------------------------- colour + DCT
__device__ __forceinline__ void rgb_to_ycc(int r, int g, int b, int& y, int& cb, int& cr) {
    // jccolor.c: FIX(x) = (int)(x * 65536 + 0.5); ONE_HALF = 32768; CBCR_OFFSET = 128 << 16
    y = (19595 * r + 38470 * g + 7471 * b + 32768) >> 16;
    cb = (-11059 * r - 21709 * g + 32768 * b + (128 << 16) + 32767) >> 16;
    cr = (32768 * r - 27439 * g - 5329 * b + (128 << 16) + 32767) >> 16;
}

__device__ __forceinline__ int descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }

// jfdctint.c jpeg_fdct_islow, one 1-D pass over eight values (CONST_BITS 13, PASS1_BITS 2)
template <bool kFirst>
__device__ __forceinline__ void fdct8(int (&d)[8]) {
    const int t0 = d[0] + d[7], t7 = d[0] - d[7];
    const int t1 = d[1] + d[6], t6 = d[1] - d[6];
    const int t2 = d[2] + d[5], t5 = d[2] - d[5];
    const int t3 = d[3] + d[4], t4 = d[3] - d[4];
    const int t10 = t0 + t3, t13 = t0 - t3, t11 = t1 + t2, t12 = t1 - t2;
    constexpr int sh = kFirst ? 13 - 2 : 13 + 2;
    if (kFirst) {
        d[0] = (t10 + t11) << 2;
        d[4] = (t10 - t11) << 2;
    } else {
        d[0] = descale(t10 + t11, 2);
        d[4] = descale(t10 - t11, 2);
    }
    int z1 = (t12 + t13) * 4433;
    d[2] = descale(z1 + t13 * 6270, sh);
    d[6] = descale(z1 + t12 * (-15137), sh);
    z1 = t4 + t7;
    int z2 = t5 + t6, z3 = t4 + t6, z4 = t5 + t7;
    const int z5 = (z3 + z4) * 9633;
    const int a4 = t4 * 2446, a5 = t5 * 16819, a6 = t6 * 25172, a7 = t7 * 12299;
    z1 *= -7373; z2 *= -20995; z3 *= -16069; z4 *= -3196;
    z3 += z5; z4 += z5;
    d[7] = descale(a4 + z1 + z3, sh);
    d[5] = descale(a5 + z2 + z4, sh);
    d[3] = descale(a6 + z2 + z3, sh);
    d[1] = descale(a7 + z1 + z4, sh);
}

__device__ __forceinline__ int nbits_of(int v) { return 32 - __clz(abs(v)); }

// One WARP owns two neighbouring MCUs (a 16 x 32 pixel strip = 12 blocks = 96 (block, row) pairs = three full rounds
// of the warp) from the pixel loads to the coefficient stores, so the passes are separated by __syncwarp only and the
// warps of an SM drift apart instead of meeting at CTA barriers.
#ifndef JPEG_DCT_MINB
#define JPEG_DCT_MINB 6            // resident CTAs per SM the register allocation aims at
#endif
constexpr int kDctWarps = 4;
constexpr int kDctThreads = kDctWarps * 32;
constexpr int kPairsPerWarp = 4;              // MCU pairs a warp walks along the row (the next one is in flight)

struct __align__(16) DctWarpSmem {
    uint8_t raw[16][96];                       // 16 rows x 32 pixels x RGB
    int ws[12][72];                            // 8 rows of 9 words: both passes are bank-conflict free
    int16_t outc[12][64];
};

__global__ void __launch_bounds__(kDctThreads, JPEG_DCT_MINB)
jpeg_dct_kernel(const uint8_t* __restrict__ frames, JpegGeom G, const __grid_constant__ JpegTables T,
                int16_t* __restrict__ coefs, int16_t* __restrict__ dcs) {
    __shared__ DctWarpSmem s_warp[kDctWarps];
    // quantisation tables transposed per column of a block: the eight values a column-pass item needs are two 128-bit
    // loads (rows of 12 words: the eight columns of a quarter-warp land in eight different bank groups)
    __shared__ __align__(16) uint32_t s_rcp[2][8][12], s_hz[2][8][12];

    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int b = blockIdx.z, my = blockIdx.y;
    const uint8_t* img = frames + (size_t)b * G.H * G.W * 3;
    // quantisation tables, staged once per CTA: exact reciprocal, and div / 2 | zigzag position << 16
    {
        const int tb = t >> 6, n = t & 63;
        s_rcp[tb][n & 7][n >> 3] = T.rcp[tb][n];
        s_hz[tb][n & 7][n >> 3] = (uint32_t)(T.div[tb][n] >> 1) | ((uint32_t)T.izz[n] << 16);
    }
    __syncthreads();
    DctWarpSmem& S = s_warp[warp];

    const int pairs = (G.mcu_cols + 1) / 2;
    const int p0 = (blockIdx.x * kDctWarps + warp) * kPairsPerWarp, p1 = min(p0 + kPairsPerWarp, pairs);
    if (p0 >= pairs) return;

    // the strip loader: 16 rows x 6 vectors of 16 bytes, three per lane; only when rows are aligned and inside the frame
    const bool can_fast = (G.W % 16 == 0) && (my * 16 + 16 <= G.H) && ((reinterpret_cast<uintptr_t>(frames) & 15u) == 0);
    const uint8_t* rowbase = img + (size_t)(my * 16) * G.W * 3;
    uint4 v[3];
    auto fetch = [&](int p) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int idx = lane + 32 * k, row = idx / 6, seg = idx - row * 6;
            v[k] = __ldg(reinterpret_cast<const uint4*>(rowbase + (size_t)row * G.W * 3 + (size_t)p * 96) + seg);
        }
    };
    bool fast = can_fast && (G.mcu_cols - p0 * 2 >= 2);
    if (fast) fetch(p0);

    for (int p = p0; p < p1; ++p) {
        const int mx0 = p * 2;
        const int n_here = min(2, G.mcu_cols - mx0);
        if (fast) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int idx = lane + 32 * k, row = idx / 6, seg = idx - row * 6;
                *reinterpret_cast<uint4*>(&S.raw[row][seg * 16]) = v[k];
            }
        }
        __syncwarp();
        // the next strip travels while this one is transformed
        const bool fast_next = can_fast && (p + 1 < p1) && (G.mcu_cols - (mx0 + 2) >= 2);
        if (fast_next) fetch(p + 1);

        // ---- phase 1: a 4 x 2 pixel patch per lane and MCU: eight luma samples, two Cb and two Cr samples
        const int qy = lane >> 2, qp = lane & 3;
        for (int m = 0; m < n_here; ++m) {
            const int gy = my * 16 + 2 * qy, gx = (mx0 + m) * 16 + 4 * qp;
            int cbs[2] = {0, 0}, crs[2] = {0, 0};
#pragma unroll
            for (int dy = 0; dy < 2; ++dy) {
                uint32_t w[3];
                if (fast) {
                    const uint32_t* q = reinterpret_cast<const uint32_t*>(&S.raw[2 * qy + dy][m * 48 + qp * 12]);
                    w[0] = q[0]; w[1] = q[1]; w[2] = q[2];
                } else {
                    const int yy = min(gy + dy, G.H - 1);
                    uint8_t px[12];
#pragma unroll
                    for (int dx = 0; dx < 4; ++dx) {
                        const uint8_t* q = img + ((size_t)yy * G.W + min(gx + dx, G.W - 1)) * 3;
                        px[dx * 3] = __ldg(q); px[dx * 3 + 1] = __ldg(q + 1); px[dx * 3 + 2] = __ldg(q + 2);
                    }
#pragma unroll
                    for (int k = 0; k < 3; ++k)
                        w[k] = px[4 * k] | (px[4 * k + 1] << 8) | (px[4 * k + 2] << 16) | ((uint32_t)px[4 * k + 3] << 24);
                }
#pragma unroll
                for (int dx = 0; dx < 4; ++dx) {
                    const int o = dx * 3;
                    const int r = (w[o >> 2] >> (8 * (o & 3))) & 255;
                    const int gch = (w[(o + 1) >> 2] >> (8 * ((o + 1) & 3))) & 255;
                    const int bl = (w[(o + 2) >> 2] >> (8 * ((o + 2) & 3))) & 255;
                    int y, cb, cr;
                    rgb_to_ycc(r, gch, bl, y, cb, cr);
                    const int yy = 2 * qy + dy, xx = 4 * qp + dx;
                    S.ws[m * 6 + (yy >> 3) * 2 + (xx >> 3)][(yy & 7) * 9 + (xx & 7)] = y - 128;
                    cbs[dx >> 1] += cb; crs[dx >> 1] += cr;
                }
            }
            if (gy >= G.H) {
                // chroma rows below the last real row group replicate the last DOWN-SAMPLED row (jcprepct.c pads the
                // down-sampler's output), which is not what the clamped luma rows give when H is even
                const int r0 = 2 * (G.c_real_rows - 1), r1 = min(r0 + 1, G.H - 1);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    cbs[h] = 0; crs[h] = 0;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int yy = (k >> 1) ? r1 : r0, xx = min(gx + 2 * h + (k & 1), G.W - 1);
                        const uint8_t* q = img + ((size_t)yy * G.W + xx) * 3;
                        int y, cb, cr;
                        rgb_to_ycc(__ldg(q), __ldg(q + 1), __ldg(q + 2), y, cb, cr);
                        cbs[h] += cb; crs[h] += cr;
                    }
                }
            }
            // h2v2_downsample: bias 1, 2, 1, 2, ... along the row; this lane owns an even and an odd column
            S.ws[m * 6 + 4][qy * 9 + 2 * qp] = ((cbs[0] + 1) >> 2) - 128;
            S.ws[m * 6 + 4][qy * 9 + 2 * qp + 1] = ((cbs[1] + 2) >> 2) - 128;
            S.ws[m * 6 + 5][qy * 9 + 2 * qp] = ((crs[0] + 1) >> 2) - 128;
            S.ws[m * 6 + 5][qy * 9 + 2 * qp + 1] = ((crs[1] + 2) >> 2) - 128;
        }
        __syncwarp();

        // ---- phase 2: row pass, item = block * 8 + row (address 9 * item + j: no bank conflicts)
        const int rounds = n_here == 2 ? 3 : 2;              // one MCU: 48 items, the second round half empty
#pragma unroll
        for (int rd = 0; rd < 3; ++rd) {
            const int item = rd * 32 + lane;
            if (rd < rounds && item < n_here * 48) {
                int d[8];
                int* row = &S.ws[0][0] + item * 9;
#pragma unroll
                for (int j = 0; j < 8; ++j) d[j] = row[j];
                fdct8<true>(d);
#pragma unroll
                for (int j = 0; j < 8; ++j) row[j] = d[j];
            }
        }
        __syncwarp();

        // ---- phase 3: column pass + quantisation, zigzag order
#pragma unroll
        for (int rd = 0; rd < 3; ++rd) {
            const int item = rd * 32 + lane;
            if (rd < rounds && item < n_here * 48) {
                const int blk = item >> 3, col = item & 7;
                const int tb = (blk == 4) | (blk == 5) | (blk >= 10);
                const uint4 ra = *reinterpret_cast<const uint4*>(&s_rcp[tb][col][0]);
                const uint4 rb = *reinterpret_cast<const uint4*>(&s_rcp[tb][col][4]);
                const uint4 ha = *reinterpret_cast<const uint4*>(&s_hz[tb][col][0]);
                const uint4 hb = *reinterpret_cast<const uint4*>(&s_hz[tb][col][4]);
                const uint32_t rcp[8] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
                const uint32_t hz[8] = {ha.x, ha.y, ha.z, ha.w, hb.x, hb.y, hb.z, hb.w};
                int d[8];
#pragma unroll
                for (int r = 0; r < 8; ++r) d[r] = S.ws[blk][r * 9 + col];
                fdct8<false>(d);
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    // (|d| + div / 2) / div, rounding half away from zero; the dividend is < 2^16: the reciprocal is exact
                    const uint32_t hzr = hz[r];
                    const int mag = (int)__umulhi((uint32_t)abs(d[r]) + (hzr & 0xffffu), rcp[r]);
                    S.outc[blk][hzr >> 16] = (int16_t)(d[r] < 0 ? -mag : mag);
                }
            }
        }
        __syncwarp();
        // dummy blocks right of / below the frame: zero AC, the DC of the previous block of the MCU (jccoefct.c)
        if ((mx0 + n_here) * 2 > G.y_blk_cols || my * 2 + 2 > G.y_blk_rows) {
            if (lane < n_here) {
                int prev = 0;
                for (int k = 0; k < 4; ++k) {
                    const int by = my * 2 + (k >> 1), bx = (mx0 + lane) * 2 + (k & 1);
                    if (by >= G.y_blk_rows || bx >= G.y_blk_cols) {
                        for (int z = 1; z < 64; ++z) S.outc[lane * 6 + k][z] = 0;
                        S.outc[lane * 6 + k][0] = (int16_t)prev;
                    }
                    prev = S.outc[lane * 6 + k][0];
                }
            }
            __syncwarp();
        }

        // ---- phase 4: coefficients out (contiguous: the MCUs of a pair are neighbours in scan order) and DCs
        const size_t mcu0 = (size_t)b * G.n_mcu + (size_t)my * G.mcu_cols + mx0;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int idx = lane + 32 * k;
            if (idx < n_here * 48)
                reinterpret_cast<uint4*>(coefs + mcu0 * 384)[idx] = reinterpret_cast<const uint4*>(&S.outc[0][0])[idx];
        }
        if (lane < n_here * 6) dcs[mcu0 * 6 + lane] = S.outc[lane][0];
        __syncwarp();
        fast = fast_next;
    }
}

// --------------------------------------------------------------------------------------- lengths and offsets
__device__ __forceinline__ int pred_block(int i) {          // previous block of the same component, -1: none
    const int k = i % 6;
    if (k >= 1 && k <= 3) return i - 1;
    if (i < 6) return -1;
    return k == 0 ? i - 3 : i - 6;
}

// block-wide exclusive scan of one value per thread; returns the exclusive prefix, *total = the sum
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* s_warp, uint32_t* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += n;
    }
    __syncthreads();                                        // s_warp may still be read from a previous call
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < nwarps ? s_warp[lane] : 0, wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t n = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += n;
        }
        s_warp[lane] = wi - w;
        if (lane == 31) s_warp[32] = wi;
    }
    __syncthreads();
    *total = s_warp[32];
    return s_warp[warp] + inc - v;
}

// One THREAD per 8x8 block encodes it (jchuff.c encode_one_block) into a private slot of kMaxBlockWords words,
// left-aligned from bit 0.  Warps 0-3 of a CTA take the 128 luma blocks of 32 MCUs, warps 4-5 their 64 chroma blocks:
// a warp then walks blocks of one kind (similar numbers of coefficients, one Huffman table).  The bit lengths are
// scanned in scan order inside the CTA; blk_meta = exclusive offset << 11 | length.
__global__ void __launch_bounds__(kPartBlocks)
jpeg_enc_kernel(const int16_t* __restrict__ coefs, const int16_t* __restrict__ dcs,
                JpegGeom G, const __grid_constant__ JpegTables T, uint32_t* __restrict__ slots,
                uint32_t* __restrict__ blk_meta, uint32_t* __restrict__ part_bits) {
    __shared__ uint32_t s_ac[2][256];
    __shared__ uint32_t s_dc[2][12];
    __shared__ uint32_t s_len[kPartBlocks];
    __shared__ uint32_t s_warp[33];
    __shared__ int16_t s_cf[64][kPartBlocks];              // the thread's block, transposed: column t
    const int t = threadIdx.x, b = blockIdx.y;
    for (int i = t; i < 512; i += kPartBlocks) s_ac[i >> 8][i & 255] = T.ac[i >> 8][i & 255];
    if (t < 24) s_dc[t / 12][t % 12] = T.dc[t / 12][t % 12];
    __syncthreads();
    const int chroma = t >= 128;
    const int li = chroma ? ((t - 128) >> 1) * 6 + 4 + (t & 1) : (t >> 2) * 6 + (t & 3);
    const int i = blockIdx.x * kPartBlocks + li;            // block index inside the frame, scan order
    uint32_t nbits = 0;
    if (i < G.n_blk) {
        const int64_t gb = (int64_t)b * G.n_blk + i;
        // slot words of the 32 blocks of a warp are interleaved (word j of lane l at (warp * 52 + j) * 32 + l):
        // lanes that are equally far along share a 128-byte line, here and in the placement pass
        uint32_t* w = slots + (((int64_t)b * G.parts + blockIdx.x) * (kPartBlocks / 32) + (t >> 5)) * (kMaxBlockWords * 32) + (t & 31);
        uint32_t cur = 0;
        int fill = 0;
        auto emit = [&](uint32_t bits, int len) {           // len <= 27
            nbits += len;
            const int room = 32 - fill;
            if (len < room) {
                cur |= bits << (room - len);
                fill += len;
            } else {
                *w = cur | (bits >> (len - room));
                w += 32;
                fill = len - room;
                cur = __funnelshift_lc(0u, bits, 32 - fill);    // bits << (32 - fill), 0 when fill == 0
            }
        };
        {   // DC difference
            const int p = pred_block(i);
            const int diff = (int)__ldg(dcs + gb) - (p >= 0 ? (int)__ldg(dcs + gb - i + p) : 0);
            const int nb = nbits_of(diff);
            const uint32_t cl = s_dc[chroma][nb];
            emit(((cl & 0xffffu) << nb) | (uint32_t)((diff < 0 ? diff - 1 : diff) & ((1 << nb) - 1)), (int)(cl >> 16) + nb);
        }
        const uint32_t* ac = s_ac[chroma];
        const int16_t* cf = coefs + gb * 64;
        // non-zero mask of the 63 AC positions: the block's 128 bytes pass through registers once, on their way into this
        // thread's column of shared memory (the walk below picks single coefficients from there); two per comparison
        uint64_t mask = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const uint4 x = __ldg(reinterpret_cast<const uint4*>(cf) + q);
            const uint32_t xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                s_cf[8 * q + 2 * j][t] = (int16_t)(xs[j] & 0xffffu);
                s_cf[8 * q + 2 * j + 1][t] = (int16_t)(xs[j] >> 16);
            }
            uint32_t byte = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t ne = __vsetne2(xs[j], 0u) & 0x00010001u;       // bit 0 / bit 16: halfword != 0
                byte |= ((ne | (ne >> 15)) & 3u) << (2 * j);
            }
            mask |= (uint64_t)byte << (8 * q);
        }
        mask &= ~1ull;                                        // position 0 is the DC slot
        int prev = 0;
        // the set bits of one 32-bit half (32-bit ffs / clear-lowest are a quarter of the 64-bit instruction count)
        auto walk = [&](uint32_t m32, int base) {
            if (!m32) return;
            auto coef = [&](int kk) { return (int)s_cf[kk][t]; };
            int k = base + __ffs((int)m32) - 1;
            int v = coef(k);
            for (;;) {
                m32 &= m32 - 1;
                int kn = 0, vn = 0;
                if (m32) {                                   // the next coefficient is in flight while this one is coded
                    kn = base + __ffs((int)m32) - 1;
                    vn = coef(kn);
                }
                const int nb = nbits_of(v);
                const int run = k - prev - 1;
                for (int zr = run >> 4; zr > 0; --zr) emit(ac[0xF0] & 0xffffu, (int)(ac[0xF0] >> 16));
                const uint32_t cl = ac[((run & 15) << 4) | nb];
                emit(((cl & 0xffffu) << nb) | (uint32_t)((v < 0 ? v - 1 : v) & ((1 << nb) - 1)), (int)(cl >> 16) + nb);
                prev = k;
                if (!m32) break;
                k = kn; v = vn;
            }
        };
        walk((uint32_t)mask, 0);
        walk((uint32_t)(mask >> 32), 32);
        if (prev != 63) emit(ac[0x00] & 0xffffu, (int)(ac[0x00] >> 16));     // EOB
        if (fill) *w = cur;
    }
    s_len[li] = nbits;
    __syncthreads();
    const uint32_t mine = s_len[t];
    uint32_t total;
    const uint32_t exc = block_exclusive_scan(mine, s_warp, &total);
    const int i2 = blockIdx.x * kPartBlocks + t;
    if (i2 < G.n_blk) blk_meta[(int64_t)b * G.n_blk + i2] = (exc << 11) | mine;
    if (t == 0) part_bits[(int64_t)b * G.parts + blockIdx.x] = total;
}

constexpr int kZeroCtas = 16;
constexpr int kOffThreads = 256;

__global__ void __launch_bounds__(kOffThreads)
jpeg_offsets_kernel(const uint32_t* __restrict__ part_bits, JpegGeom G, uint32_t* __restrict__ part_off,
                    uint32_t* __restrict__ frame_bits, uint32_t* __restrict__ stream) {
    __shared__ uint32_t s_warp[33];
    const int t = threadIdx.x, z = blockIdx.x, b = blockIdx.y;
    const uint32_t* pb = part_bits + (int64_t)b * G.parts;
    const int per = (G.parts + kOffThreads - 1) / kOffThreads;
    const int i0 = min(t * per, G.parts), i1 = min(i0 + per, G.parts);
    uint32_t sum = 0;
    for (int i = i0; i < i1; ++i) sum += __ldg(pb + i);
    uint32_t total;
    uint32_t run = block_exclusive_scan(sum, s_warp, &total);
    if (z == 0) {
        for (int i = i0; i < i1; ++i) {
            part_off[(int64_t)b * G.parts + i] = run;
            run += __ldg(pb + i);
        }
        if (t == 0) frame_bits[b] = total;
    }
    // the scan of this frame occupies `total` bits: clear exactly those words (the Huffman pass ORs into its
    // boundary words) and pad the last byte with 1-bits (jchuff.c flush_bits); kZeroCtas CTAs share the frame
    uint32_t* st = stream + (int64_t)b * G.words_cap;
    const uint32_t nvec = ((total + 31) / 32 + 1 + 3) / 4;
    const uint32_t slice = (nvec + kZeroCtas - 1) / kZeroCtas;
    const uint32_t v0 = z * slice, v1 = min(v0 + slice, nvec);
    uint4* st4 = reinterpret_cast<uint4*>(st);
    for (uint32_t i = v0 + t; i < v1; i += kOffThreads) st4[i] = make_uint4(0, 0, 0, 0);
    const uint32_t pad_word = total >> 5, pad = (8 - (total & 7)) & 7;
    if (pad && (pad_word >> 2) >= v0 && (pad_word >> 2) < v1) {
        __syncthreads();
        if (t == 0) st[pad_word] = ((1u << pad) - 1u) << (32 - (total & 31) - pad);
    }
}

// ------------------------------------------------------------------------------------------------ placement
// Moves every block's bits from its slot to its bit offset in the frame's scan: a funnel shift per word; the first
// and the last word of a block are shared with its neighbours (atomic OR into the zero-filled scan), the others
// are plain stores.
__global__ void __launch_bounds__(kPartBlocks)
jpeg_place_kernel(const uint32_t* __restrict__ slots, const uint32_t* __restrict__ blk_meta,
                  const uint32_t* __restrict__ part_off, JpegGeom G, uint32_t* __restrict__ stream) {
    const int t = threadIdx.x, b = blockIdx.y;
    // same thread <-> block mapping as the entropy pass, so that a warp reads its interleaved slots line by line
    const int li = t >= 128 ? ((t - 128) >> 1) * 6 + 4 + (t & 1) : (t >> 2) * 6 + (t & 3);
    const int i = blockIdx.x * kPartBlocks + li;
    if (i >= G.n_blk) return;
    const uint32_t m = __ldg(blk_meta + (int64_t)b * G.n_blk + i);
    const uint32_t len = m & 2047u;
    const uint32_t start = __ldg(part_off + (int64_t)b * G.parts + blockIdx.x) + (m >> 11);
    const uint32_t* src = slots + (((int64_t)b * G.parts + blockIdx.x) * (kPartBlocks / 32) + (t >> 5)) * (kMaxBlockWords * 32) + (t & 31);
    uint32_t* dst = stream + (int64_t)b * G.words_cap + (start >> 5);
    const uint32_t sh = start & 31u;
    const int nsrc = (int)((len + 31) >> 5), nd = (int)((sh + len + 31) >> 5);
    uint32_t prev = 0;
    for (int j = 0; j < nd; ++j) {
        const uint32_t cur = j < nsrc ? __ldg(src + j * 32) : 0u;
        const uint32_t w = __funnelshift_r(cur, prev, sh);  // (prev : cur) >> sh
        if (j == 0 || j == nd - 1) atomicOr(dst + j, w);
        else dst[j] = w;
        prev = cur;
    }
}

// ------------------------------------------------------------------------------------------ byte stuffing
__device__ __forceinline__ int count_ff(uint32_t w) { return __popc(__vcmpeq4(w, 0xffffffffu)) >> 3; }

constexpr int kStuffThreads = 256;

__global__ void __launch_bounds__(kStuffThreads)
jpeg_ffcount_kernel(const uint32_t* __restrict__ stream, const uint32_t* __restrict__ frame_bits, JpegGeom G,
                    uint32_t* __restrict__ chunk_ff) {
    __shared__ uint32_t s_n;
    const int b = blockIdx.y, t = threadIdx.x;
    const uint32_t nbytes = (frame_bits[b] + 7) >> 3;
    const int n_chunks = (int)((nbytes + kChunkBytes - 1) / kChunkBytes);
    const uint4* st = reinterpret_cast<const uint4*>(stream + (int64_t)b * G.words_cap);
    // bytes past nbytes inside the last vector are zero (cleared by the offsets kernel), never 0xFF
    for (int c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        if (t == 0) s_n = 0;
        __syncthreads();
        int n = 0;
        if ((uint32_t)c * kChunkBytes + 16u * t < nbytes) {
            const uint4 x = st[(int64_t)c * (kChunkBytes / 16) + t];
            n = count_ff(x.x) + count_ff(x.y) + count_ff(x.z) + count_ff(x.w);
        }
        n = __reduce_add_sync(0xffffffffu, n);
        if ((t & 31) == 0 && n) atomicAdd(&s_n, (uint32_t)n);
        __syncthreads();
        if (t == 0) chunk_ff[(int64_t)b * G.chunks_cap + c] = s_n;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kStuffThreads)
jpeg_stuff_kernel(const uint32_t* __restrict__ stream, const uint32_t* __restrict__ frame_bits,
                  const uint32_t* __restrict__ chunk_ff, JpegGeom G, const __grid_constant__ JpegHeader hdr,
                  uint8_t* __restrict__ out, int64_t out_stride, int32_t* __restrict__ len_out) {
    __shared__ uint8_t s_out[2 * kChunkBytes];
    __shared__ uint32_t s_warp[33];
    __shared__ uint32_t s_sum[2];
    const int b = blockIdx.y, t = threadIdx.x;
    const uint32_t nbytes = (frame_bits[b] + 7) >> 3;
    const int n_chunks = (int)((nbytes + kChunkBytes - 1) / kChunkBytes);
    const uint32_t* cff = chunk_ff + (int64_t)b * G.chunks_cap;
    uint8_t* dst = out + (int64_t)b * out_stride;
    const uint4* st = reinterpret_cast<const uint4*>(stream + (int64_t)b * G.words_cap);
    for (int c = blockIdx.x; c < n_chunks || c == 0; c += gridDim.x) {
        // 0xFF bytes in the chunks before this one, and in the whole scan (decides whether the file fits)
        if (t < 2) s_sum[t] = 0;
        __syncthreads();
        uint32_t pre = 0, all = 0;
        for (int i = t; i < n_chunks; i += kStuffThreads) {
            const uint32_t v = __ldg(cff + i);
            all += v;
            if (i < c) pre += v;
        }
        pre = __reduce_add_sync(0xffffffffu, pre);
        all = __reduce_add_sync(0xffffffffu, all);
        if ((t & 31) == 0) { atomicAdd(&s_sum[0], pre); atomicAdd(&s_sum[1], all); }
        __syncthreads();
        pre = s_sum[0]; all = s_sum[1];
        const int64_t need = (int64_t)kHeaderBytes + nbytes + all + 2;
        if (need > out_stride) {                               // does not fit: nothing is written but the size needed
            if (c == 0 && t == 0) len_out[b] = (int32_t)-need;
            return;
        }
        if (c == 0) {
            for (int i = t; i < kHeaderBytes; i += kStuffThreads) dst[i] = hdr.bytes[i];
            if (t == 0) { dst[need - 2] = 0xFF; dst[need - 1] = 0xD9; len_out[b] = (int32_t)need; }
        }
        if (c >= n_chunks) break;
        const uint32_t base = (uint32_t)c * kChunkBytes;
        const uint32_t mine = base + 16u * t;                   // first byte of this thread
        uint4 x = make_uint4(0, 0, 0, 0);
        int valid = 0;
        if (mine < nbytes) {
            x = st[(int64_t)c * (kChunkBytes / 16) + t];
            valid = (int)min(16u, nbytes - mine);
        }
        const uint32_t w[4] = {x.x, x.y, x.z, x.w};
        const int ff = count_ff(x.x) + count_ff(x.y) + count_ff(x.z) + count_ff(x.w);
        uint32_t total;
        const uint32_t before = block_exclusive_scan((uint32_t)ff, s_warp, &total);
        uint32_t o = 16u * t + before;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            if (k < valid) {
                const uint8_t v = (uint8_t)(w[k >> 2] >> (24 - 8 * (k & 3)));   // the scan is big-endian in its words
                s_out[o++] = v;
                if (v == 0xFF) s_out[o++] = 0;
            }
        }
        __syncthreads();
        const uint32_t n_out = min((uint32_t)kChunkBytes, nbytes - base) + total;
        uint8_t* d = dst + kHeaderBytes + base + pre;
        for (uint32_t i = t; i < n_out; i += kStuffThreads) d[i] = s_out[i];
        __syncthreads();
    }
}

}  // namespace

// ----------------------------------------------------------------------------------------------- C ABI
extern "C" int64_t mlp_jpeg_max_bytes(int frame_h, int frame_w) {
    if (frame_h <= 0 || frame_w <= 0) return 0;
    const JpegGeom g = make_geom(frame_h, frame_w);
    return (int64_t)kHeaderBytes + 2 * ((int64_t)g.n_blk * kMaxBlockWords * 4) + 2;
}

extern "C" int mlp_jpeg_header(int frame_h, int frame_w, int quality, uint8_t* out_host, int capacity) {
    MLP_CHECK_ARG(out_host != nullptr && capacity >= kHeaderBytes, "mlp_jpeg_header: need %d bytes", kHeaderBytes);
    MLP_CHECK_ARG(frame_h > 0 && frame_w > 0 && frame_h <= 65535 && frame_w <= 65535,
                  "mlp_jpeg_header: frame %dx%d outside 1..65535", frame_h, frame_w);
    JpegTables T;
    JpegHeader H;
    build_tables(frame_h, frame_w, quality, &T, &H);
    memcpy(out_host, H.bytes, kHeaderBytes);
    return kHeaderBytes;
}

extern "C" int mlp_jpeg_encode(mlp_ctx* ctx, const uint8_t* images_dev, int batch, int frame_h, int frame_w, int quality,
                               uint8_t* out_dev, int64_t out_stride, int32_t* len_dev, mlp_stream_t stream_) {
    MLP_CHECK_ARG(ctx != nullptr, "mlp_jpeg_encode: null context");
    MLP_CHECK_ARG(images_dev && out_dev && len_dev, "mlp_jpeg_encode: null pointer");
    MLP_CHECK_ARG(batch > 0, "mlp_jpeg_encode: batch %d", batch);
    MLP_CHECK_ARG(frame_h > 0 && frame_w > 0 && frame_h <= 65535 && frame_w <= 65535,
                  "mlp_jpeg_encode: frame %dx%d outside 1..65535 (SOF0 holds 16-bit sizes)", frame_h, frame_w);
    MLP_CHECK_ARG(quality >= 1 && quality <= 100, "mlp_jpeg_encode: quality %d outside 1..100", quality);
    MLP_CHECK_ARG(out_stride >= kHeaderBytes + 2 + 4, "mlp_jpeg_encode: out_stride %lld too small", (long long)out_stride);
    const JpegGeom G = make_geom(frame_h, frame_w);
    MLP_CHECK_ARG((int64_t)G.n_blk * 1664 < ((int64_t)1 << 32), "mlp_jpeg_encode: frame %dx%d too large (bit offsets are 32-bit)",
                  frame_h, frame_w);
    MLP_CHECK_ARG(batch <= 65535, "mlp_jpeg_encode: batch %d > 65535", batch);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DeviceGuard guard(ctx->device);

    const int64_t nb = (int64_t)batch * G.n_blk;
    auto up = [](int64_t v) { return (v + 255) & ~(int64_t)255; };
    const int64_t o_coef = 0;
    const int64_t o_dc = o_coef + up(nb * 64 * 2);
    const int64_t o_loc = o_dc + up(nb * 2);
    const int64_t o_slots = o_loc + up(nb * 4);
    const int64_t o_pbits = o_slots + up((int64_t)batch * G.parts * kPartBlocks * kMaxBlockWords * 4);
    const int64_t o_poff = o_pbits + up((int64_t)batch * G.parts * 4);
    const int64_t o_bits = o_poff + up((int64_t)batch * G.parts * 4);
    const int64_t o_cff = o_bits + up((int64_t)batch * 4);
    const int64_t o_stream = o_cff + up((int64_t)batch * G.chunks_cap * 4);
    const int64_t bytes = o_stream + up((int64_t)batch * G.words_cap * 4);
    int rc = mlp_ensure_scratch(ctx, MLP_ARENA_JPEG, bytes);
    if (rc != MLP_OK) return rc;
    char* base = static_cast<char*>(ctx->arena[MLP_ARENA_JPEG]);
    int16_t* coefs = reinterpret_cast<int16_t*>(base + o_coef);
    int16_t* dcs = reinterpret_cast<int16_t*>(base + o_dc);
    uint32_t* blk_meta = reinterpret_cast<uint32_t*>(base + o_loc);
    uint32_t* slots = reinterpret_cast<uint32_t*>(base + o_slots);
    uint32_t* part_bits = reinterpret_cast<uint32_t*>(base + o_pbits);
    uint32_t* part_off = reinterpret_cast<uint32_t*>(base + o_poff);
    uint32_t* frame_bits = reinterpret_cast<uint32_t*>(base + o_bits);
    uint32_t* chunk_ff = reinterpret_cast<uint32_t*>(base + o_cff);
    uint32_t* scan = reinterpret_cast<uint32_t*>(base + o_stream);

    JpegTables T;
    JpegHeader H;
    build_tables(frame_h, frame_w, quality, &T, &H);

    ProfScope prof(ctx, MLP_ST_JPEG, stream);
    // frames taller than 65535 MCU rows cannot exist (H <= 65535), so the grid's y extent is safe
    const int dpairs = (G.mcu_cols + 1) / 2;
    dim3 dgrid((dpairs + kDctWarps * kPairsPerWarp - 1) / (kDctWarps * kPairsPerWarp), G.mcu_rows, batch);
    jpeg_dct_kernel<<<dgrid, kDctThreads, 0, stream>>>(images_dev, G, T, coefs, dcs);
    MLP_LAUNCH_CHECK(ctx);
    jpeg_enc_kernel<<<dim3(G.parts, batch), kPartBlocks, 0, stream>>>(coefs, dcs, G, T, slots, blk_meta, part_bits);
    MLP_LAUNCH_CHECK(ctx);
    jpeg_offsets_kernel<<<dim3(kZeroCtas, batch), kOffThreads, 0, stream>>>(part_bits, G, part_off, frame_bits, scan);
    MLP_LAUNCH_CHECK(ctx);
    jpeg_place_kernel<<<dim3(G.parts, batch), kPartBlocks, 0, stream>>>(slots, blk_meta, part_off, G, scan);
    MLP_LAUNCH_CHECK(ctx);
    // typical scans take 0.3-1 byte per pixel; the chunk loops cover the rest
    int sgrid = (int)(((int64_t)frame_h * frame_w + kChunkBytes - 1) / kChunkBytes);
    sgrid = sgrid < 1 ? 1 : (sgrid > G.chunks_cap ? G.chunks_cap : sgrid);
    sgrid = sgrid > 65535 ? 65535 : sgrid;
    jpeg_ffcount_kernel<<<dim3(sgrid, batch), kStuffThreads, 0, stream>>>(scan, frame_bits, G, chunk_ff);
    MLP_LAUNCH_CHECK(ctx);
    jpeg_stuff_kernel<<<dim3(sgrid, batch), kStuffThreads, 0, stream>>>(scan, frame_bits, chunk_ff, G, H, out_dev, out_stride,
                                                                        len_dev);
    MLP_LAUNCH_CHECK(ctx);
    return MLP_OK;
}
