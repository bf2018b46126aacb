// jpeg.cu — the JPEG encode at the end of the serving graph ("overlay + JPEG encode"):
// EncodeImageContent.call (/root/reference/engine/layers/misc.py:343-351) = tf.io.encode_jpeg(image) with every
// attribute at its default, wired in /root/reference/road_project/setup/serving.py:41.  TensorFlow's kernel drives
// libjpeg(-turbo): quality 95 (force_baseline), RGB -> YCbCr 4:2:0, JDCT_ISLOW, the Annex K Huffman tables,
// baseline sequential, JFIF 300x300 dpi.  All of it is integer work, so the byte stream is reproduced EXACTLY
// (oracle/jpeg_oracle.py restates it and is pinned byte for byte against libjpeg-turbo's own output).
//
// Five launches per batch of frames, everything stays on the device:
//   jpeg_dct_kernel     one CTA per four MCUs of a row: colour conversion (jccolor.c fixed point), h2v2 chroma
//                       down-sampling with the alternating 1,2 bias (jcsample.c), edge replication and dummy blocks
//                       (jcprepct.c / jccoefct.c), jpeg_fdct_islow in registers with conflict-free shared-memory
//                       transposes (jfdctint.c), quantisation (jcdctmgr.c); writes zigzag int16 coefficients and per
//                       block (AC bit length, DC)                       HBM: 3 B/px in, 3 B/px out
//   jpeg_scan_kernel    one CTA per frame: DC differences -> bit length of every block -> exclusive scan (bit offset
//                       of every block in the frame's scan), zero-fills exactly the words the scan will occupy
//   jpeg_huff_kernel    one warp per 8x8 block: every lane composes the bit field of two coefficients (ZRL run, run/size
//                       code, value bits; jchuff.c encode_one_block), a warp scan places them, the block is assembled
//                       in shared memory and stored at its bit offset (only the two boundary words are atomics)
//   jpeg_ffscan_kernel  one CTA per frame: 0xFF bytes per 4 KB chunk of the scan + exclusive scan (byte stuffing moves
//                       every later byte) and the final length
//   jpeg_stuff_kernel   header (jcmarker.c), scan bytes with a 0x00 after every 0xFF, EOI
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

// Everything a launch needs that depends only on (H, W, quality): passed by value (kernel parameter space).
struct JpegTables {
    uint16_t div[2][64];       // natural order, q << 3 (jcdctmgr.c start_pass_fdctmgr)
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
        for (int i = 0; i < 64; ++i) T->div[t][i] = (uint16_t)(q[t][i] << 3);
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

// Bit length of the AC part of one block held as (position lane, position lane + 32) per lane.
// ac_len: smem table of code lengths for this block's component.
__device__ __forceinline__ int ac_bit_length(int c0, int c1, int lane, const uint8_t* ac_len) {
    const uint32_t lo = __ballot_sync(0xffffffffu, c0 != 0) & ~1u;          // position 0 is the DC slot
    const uint32_t hi = __ballot_sync(0xffffffffu, c1 != 0);
    const uint64_t nz = ((uint64_t)hi << 32) | lo;
    const int zrl = ac_len[0xF0];
    int bits = 0;
    if (lane > 0 && c0 != 0) {
        const uint32_t below = lo & ((1u << lane) - 1u);
        const int prev = below ? 31 - __clz(below) : 0;
        const int run = lane - prev - 1, nb = nbits_of(c0);
        bits += (run >> 4) * zrl + ac_len[((run & 15) << 4) | nb] + nb;
    }
    if (c1 != 0) {
        const uint64_t below = nz & ((1ull << (lane + 32)) - 1ull);
        const int prev = below ? 63 - __clzll(below) : 0;
        const int run = lane + 32 - prev - 1, nb = nbits_of(c1);
        bits += (run >> 4) * zrl + ac_len[((run & 15) << 4) | nb] + nb;
    } else if (lane == 31) {
        bits += ac_len[0x00];                                               // EOB
    }
    return __reduce_add_sync(0xffffffffu, bits);
}

constexpr int kDctThreads = 256;
constexpr int kMcuPerCta = 4;

__global__ void __launch_bounds__(kDctThreads)
jpeg_dct_kernel(const uint8_t* __restrict__ frames, JpegGeom G, const __grid_constant__ JpegTables T,
                int16_t* __restrict__ coefs, uint32_t* __restrict__ meta) {
    __shared__ __align__(16) uint8_t raw[16][kMcuPerCta * 48];
    __shared__ int ws[kMcuPerCta * 6][72];                 // 8 rows of 9: both passes are bank-conflict free
    __shared__ __align__(16) int16_t outc[kMcuPerCta * 6][64];
    __shared__ uint8_t s_aclen[2][256];
    __shared__ uint8_t s_izz[64];
    __shared__ uint16_t s_div[2][64];

    const int t = threadIdx.x;
    const int b = blockIdx.z, my = blockIdx.y, mx0 = blockIdx.x * kMcuPerCta;
    const int n_here = min(kMcuPerCta, G.mcu_cols - mx0);
    const uint8_t* img = frames + (int64_t)b * G.H * G.W * 3;

    for (int i = t; i < 512; i += kDctThreads) s_aclen[i >> 8][i & 255] = (uint8_t)(T.ac[i >> 8][i & 255] >> 16);
    if (t < 64) s_izz[t] = T.izz[t];
    if (t < 128) s_div[t >> 6][t & 63] = T.div[t >> 6][t & 63];

    // ---- phase 0: the 16 x 64 pixel strip of this CTA, 16-byte loads when rows are aligned and inside the frame
    const bool fast = (G.W % 16 == 0) && (my * 16 + 16 <= G.H) && (n_here == kMcuPerCta) &&
                      ((reinterpret_cast<uintptr_t>(frames) & 15u) == 0);
    if (fast) {
        if (t < 16 * 12) {
            const int row = t / 12, seg = t - row * 12;
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(img + ((int64_t)(my * 16 + row) * G.W + mx0 * 16) * 3) + seg);
            *reinterpret_cast<uint4*>(&raw[row][seg * 16]) = v;
        }
    }
    __syncthreads();

    // ---- phase 1: one 2x2 quad per thread: four luma samples, one Cb and one Cr sample
    const int m = t >> 6, q = t & 63, qy = q >> 3, qx = q & 7;
    if (m < n_here) {
        int cbs = 0, crs = 0;
        const int gy = my * 16 + 2 * qy, gx = (mx0 + m) * 16 + 2 * qx;
#pragma unroll
        for (int dy = 0; dy < 2; ++dy) {
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
                int r, g, bl;
                if (fast) {
                    const uint8_t* p = &raw[2 * qy + dy][m * 48 + (2 * qx + dx) * 3];
                    r = p[0]; g = p[1]; bl = p[2];
                } else {
                    const int yy = min(gy + dy, G.H - 1), xx = min(gx + dx, G.W - 1);
                    const uint8_t* p = img + ((int64_t)yy * G.W + xx) * 3;
                    r = __ldg(p); g = __ldg(p + 1); bl = __ldg(p + 2);
                }
                int y, cb, cr;
                rgb_to_ycc(r, g, bl, y, cb, cr);
                const int yy = 2 * qy + dy, xx = 2 * qx + dx;
                ws[m * 6 + (yy >> 3) * 2 + (xx >> 3)][(yy & 7) * 9 + (xx & 7)] = y - 128;
                cbs += cb; crs += cr;
            }
        }
        if (gy >= G.H) {
            // chroma rows below the last real row group replicate the last DOWN-SAMPLED row (jcprepct.c pads the
            // down-sampler's output), which is not what the clamped luma rows give when H is even
            const int r0 = 2 * (G.c_real_rows - 1), r1 = min(r0 + 1, G.H - 1);
            cbs = 0; crs = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int yy = (k >> 1) ? r1 : r0, xx = min(gx + (k & 1), G.W - 1);
                const uint8_t* p = img + ((int64_t)yy * G.W + xx) * 3;
                int y, cb, cr;
                rgb_to_ycc(__ldg(p), __ldg(p + 1), __ldg(p + 2), y, cb, cr);
                cbs += cb; crs += cr;
            }
        }
        const int bias = 1 + (qx & 1);                       // h2v2_downsample: 1, 2, 1, 2, ...
        ws[m * 6 + 4][qy * 9 + qx] = ((cbs + bias) >> 2) - 128;
        ws[m * 6 + 5][qy * 9 + qx] = ((crs + bias) >> 2) - 128;
    }
    __syncthreads();

    // ---- phase 2: row pass (thread = block * 8 + row; address 9 * t + j: no bank conflicts)
    if (t < kMcuPerCta * 48) {
        int d[8];
        int* row = &ws[t >> 3][(t & 7) * 9];
#pragma unroll
        for (int j = 0; j < 8; ++j) d[j] = row[j];
        fdct8<true>(d);
#pragma unroll
        for (int j = 0; j < 8; ++j) row[j] = d[j];
    }
    __syncthreads();

    // ---- phase 3: column pass + quantisation, zigzag order
    if (t < kMcuPerCta * 48) {
        const int blk = t >> 3, c = t & 7, k = blk % 6, mm = blk / 6;
        bool dummy = false;
        if (k < 4) {
            const int by = my * 2 + (k >> 1), bx = (mx0 + mm) * 2 + (k & 1);
            dummy = by >= G.y_blk_rows || bx >= G.y_blk_cols;
        }
        int d[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) d[r] = ws[blk][r * 9 + c];
        fdct8<false>(d);
        const uint16_t* dv = s_div[k >= 4];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int n = r * 8 + c;
            const int qv = dv[n];
            const int mag = (int)((unsigned)(abs(d[r]) + (qv >> 1)) / (unsigned)qv);
            outc[blk][s_izz[n]] = dummy ? (int16_t)0 : (int16_t)(d[r] < 0 ? -mag : mag);
        }
    }
    __syncthreads();
    // dummy blocks right of / below the frame carry the DC of the previous block of the MCU (jccoefct.c)
    if (t < n_here) {
        int prev = 0;
        for (int k = 0; k < 4; ++k) {
            const int by = my * 2 + (k >> 1), bx = (mx0 + t) * 2 + (k & 1);
            if (by >= G.y_blk_rows || bx >= G.y_blk_cols) outc[t * 6 + k][0] = (int16_t)prev;
            prev = outc[t * 6 + k][0];
        }
    }
    __syncthreads();

    // ---- phase 4: coefficients out (contiguous: the MCUs of a CTA are neighbours in scan order), AC bit lengths
    const int64_t mcu0 = (int64_t)b * G.n_mcu + (int64_t)my * G.mcu_cols + mx0;
    if (t < n_here * 48)
        reinterpret_cast<uint4*>(coefs + mcu0 * 384)[t] = reinterpret_cast<const uint4*>(&outc[0][0])[t];
    const int warp = t >> 5, lane = t & 31;
    for (int blk = warp; blk < n_here * 6; blk += kDctThreads / 32) {
        const int c0 = outc[blk][lane], c1 = outc[blk][lane + 32];
        const int bits = ac_bit_length(c0, c1, lane, s_aclen[(blk % 6) >= 4]);
        if (lane == 0) meta[mcu0 * 6 + blk] = ((uint32_t)bits << 16) | (uint32_t)(uint16_t)(int16_t)c0;
    }
}

// ------------------------------------------------------------------------------------------------- scan
__device__ __forceinline__ int pred_block(int i) {          // previous block of the same component, -1: none
    const int k = i % 6;
    if (k >= 1 && k <= 3) return i - 1;
    if (i < 6) return -1;
    return k == 0 ? i - 3 : i - 6;
}
__device__ __forceinline__ int dc_of(uint32_t m) { return (int)(int16_t)(m & 0xffffu); }

constexpr int kScanThreads = 1024;

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

__global__ void __launch_bounds__(kScanThreads)
jpeg_scan_kernel(const uint32_t* __restrict__ meta, JpegGeom G, const __grid_constant__ JpegTables T,
                 uint32_t* __restrict__ blk_off, uint32_t* __restrict__ frame_bits, uint32_t* __restrict__ stream) {
    __shared__ uint32_t s_warp[33];
    __shared__ uint8_t s_dclen[2][12];
    const int b = blockIdx.x, t = threadIdx.x;
    if (t < 24) s_dclen[t / 12][t % 12] = (uint8_t)(T.dc[t / 12][t % 12] >> 16);
    __syncthreads();
    const uint32_t* mt = meta + (int64_t)b * G.n_blk;
    uint32_t* off = blk_off + (int64_t)b * G.n_blk;
    const int per = (G.n_blk + kScanThreads - 1) / kScanThreads;
    const int i0 = min(t * per, G.n_blk), i1 = min(i0 + per, G.n_blk);
    auto length_of = [&](int i) -> uint32_t {
        const uint32_t mi = __ldg(mt + i);
        const int p = pred_block(i);
        const int diff = dc_of(mi) - (p >= 0 ? dc_of(__ldg(mt + p)) : 0);
        const int nb = nbits_of(diff);
        return (mi >> 16) + s_dclen[(i % 6) >= 4][nb] + nb;
    };
    uint32_t sum = 0;
    for (int i = i0; i < i1; ++i) sum += length_of(i);
    uint32_t total;
    uint32_t run = block_exclusive_scan(sum, s_warp, &total);
    for (int i = i0; i < i1; ++i) {
        off[i] = run;
        run += length_of(i);
    }
    // the scan of this frame occupies `total` bits: clear exactly those words (the Huffman pass ORs into them) and
    // pad the last byte with 1-bits (jchuff.c flush_bits)
    uint32_t* st = stream + (int64_t)b * G.words_cap;
    const uint32_t nwords = (total + 31) / 32 + 1;
    uint4* st4 = reinterpret_cast<uint4*>(st);
    for (uint32_t i = t; i < (nwords + 3) / 4; i += kScanThreads) st4[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    if (t == 0) {
        frame_bits[b] = total;
        const uint32_t pad = (8 - (total & 7)) & 7;
        if (pad) st[total >> 5] = ((1u << pad) - 1u) << (32 - (total & 31) - pad);
    }
}

// ---------------------------------------------------------------------------------------------- Huffman
constexpr int kHuffWarps = 8;
constexpr int kStageWords = 56;

__device__ __forceinline__ void field_append(uint64_t& f, int& n, uint32_t code_len) {
    const int len = (int)(code_len >> 16);
    f = (f << len) | (code_len & 0xffffu);
    n += len;
}

// ORs the `n` low bits of `f` (MSB first) into the word array at bit position q
__device__ __forceinline__ void stage_put(uint32_t* words, uint32_t q, uint64_t f, int n) {
    if (n == 0) return;
    const uint64_t F = f << (64 - n);
    const uint32_t hi = (uint32_t)(F >> 32), lo = (uint32_t)F;
    const uint32_t sh = q & 31, w = q >> 5;
    atomicOr(&words[w], hi >> sh);
    if (sh + n > 32) atomicOr(&words[w + 1], __funnelshift_r(lo, hi, sh));
    if (sh + n > 64) atomicOr(&words[w + 2], __funnelshift_r(0u, lo, sh));
}

__global__ void __launch_bounds__(kHuffWarps * 32)
jpeg_huff_kernel(const int16_t* __restrict__ coefs, const uint32_t* __restrict__ meta,
                 const uint32_t* __restrict__ blk_off, JpegGeom G, const __grid_constant__ JpegTables T, int batch,
                 uint32_t* __restrict__ stream) {
    __shared__ uint32_t s_ac[2][256];
    __shared__ uint32_t s_dc[2][12];
    __shared__ uint32_t s_stage[kHuffWarps][kStageWords];
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    for (int i = t; i < 512; i += kHuffWarps * 32) s_ac[i >> 8][i & 255] = T.ac[i >> 8][i & 255];
    if (t < 24) s_dc[t / 12][t % 12] = T.dc[t / 12][t % 12];
    __syncthreads();

    const int64_t total_blk = (int64_t)batch * G.n_blk;
    uint32_t* stage = s_stage[warp];
    for (int64_t gb = (int64_t)blockIdx.x * kHuffWarps + warp; gb < total_blk; gb += (int64_t)gridDim.x * kHuffWarps) {
        const int b = (int)(gb / G.n_blk), i = (int)(gb - (int64_t)b * G.n_blk);
        const int chroma = (i % 6) >= 4;
        const int16_t* cf = coefs + gb * 64;
        int c0 = cf[lane];
        const int c1 = cf[lane + 32];
        const uint32_t* ac = s_ac[chroma];
        stage[lane] = 0;
        if (lane < kStageWords - 32) stage[lane + 32] = 0;

        const uint32_t lo = __ballot_sync(0xffffffffu, c0 != 0) & ~1u;
        const uint32_t hi = __ballot_sync(0xffffffffu, c1 != 0);
        const uint64_t nz = ((uint64_t)hi << 32) | lo;

        uint64_t f0 = 0, f1 = 0;
        int n0 = 0, n1 = 0;
        if (lane == 0) {                                      // DC difference (jchuff.c encode_one_block)
            const int p = pred_block(i);
            const int diff = c0 - (p >= 0 ? dc_of(__ldg(meta + (int64_t)b * G.n_blk + p)) : 0);
            const int nb = nbits_of(diff);
            field_append(f0, n0, s_dc[chroma][nb]);
            f0 = (f0 << nb) | (uint32_t)((diff < 0 ? diff - 1 : diff) & ((1 << nb) - 1));
            n0 += nb;
        } else if (c0 != 0) {
            const uint32_t below = lo & ((1u << lane) - 1u);
            const int prev = below ? 31 - __clz(below) : 0;
            const int run = lane - prev - 1, nb = nbits_of(c0);
            for (int z = 0; z < (run >> 4); ++z) field_append(f0, n0, ac[0xF0]);
            field_append(f0, n0, ac[((run & 15) << 4) | nb]);
            f0 = (f0 << nb) | (uint32_t)((c0 < 0 ? c0 - 1 : c0) & ((1 << nb) - 1));
            n0 += nb;
        }
        if (c1 != 0) {
            const uint64_t below = nz & ((1ull << (lane + 32)) - 1ull);
            const int prev = below ? 63 - __clzll(below) : 0;
            const int run = lane + 32 - prev - 1, nb = nbits_of(c1);
            for (int z = 0; z < (run >> 4); ++z) field_append(f1, n1, ac[0xF0]);
            field_append(f1, n1, ac[((run & 15) << 4) | nb]);
            f1 = (f1 << nb) | (uint32_t)((c1 < 0 ? c1 - 1 : c1) & ((1 << nb) - 1));
            n1 += nb;
        } else if (lane == 31) {
            field_append(f1, n1, ac[0x00]);                   // EOB
        }
        // positions 0..31 come before 32..63: one warp scan over the packed pair of lengths
        uint32_t inc = (uint32_t)n0 | ((uint32_t)n1 << 16);
        const uint32_t own = inc;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t n = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += n;
        }
        const uint32_t tot = __shfl_sync(0xffffffffu, inc, 31);
        const uint32_t exc = inc - own;
        const uint32_t len0 = tot & 0xffffu, blk_len = len0 + (tot >> 16);
        const uint32_t start = __ldg(blk_off + gb);
        const uint32_t sh0 = start & 31;
        __syncwarp();
        stage_put(stage, sh0 + (exc & 0xffffu), f0, n0);
        stage_put(stage, sh0 + len0 + (exc >> 16), f1, n1);
        __syncwarp();
        uint32_t* dst = stream + (int64_t)b * G.words_cap + (start >> 5);
        const int nwords = (int)((sh0 + blk_len + 31) >> 5);
        for (int j = lane; j < nwords; j += 32) {
            const uint32_t v = stage[j];
            if (j == 0 || j == nwords - 1) atomicOr(dst + j, v);
            else dst[j] = v;
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------ byte stuffing
__device__ __forceinline__ int count_ff(uint32_t w) { return __popc(__vcmpeq4(w, 0xffffffffu)) >> 3; }

__global__ void __launch_bounds__(kScanThreads)
jpeg_ffscan_kernel(const uint32_t* __restrict__ stream, const uint32_t* __restrict__ frame_bits, JpegGeom G,
                   uint32_t* __restrict__ chunk_ff, int64_t out_stride, int32_t* __restrict__ len_out) {
    __shared__ uint32_t s_warp[33];
    const int b = blockIdx.x, t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const uint32_t nbytes = (frame_bits[b] + 7) >> 3;
    const int n_chunks = (int)((nbytes + kChunkBytes - 1) / kChunkBytes);
    const uint4* st = reinterpret_cast<const uint4*>(stream + (int64_t)b * G.words_cap);
    uint32_t* cff = chunk_ff + (int64_t)b * G.chunks_cap;
    // bytes past nbytes inside the last word are zero (cleared by the scan kernel), never 0xFF
    for (int c = warp; c < n_chunks; c += kScanThreads / 32) {
        const uint32_t vecs = min((uint32_t)(kChunkBytes / 16), (nbytes - (uint32_t)c * kChunkBytes + 15) / 16);
        int n = 0;
        for (uint32_t v = lane; v < vecs; v += 32) {
            const uint4 x = st[(int64_t)c * (kChunkBytes / 16) + v];
            n += count_ff(x.x) + count_ff(x.y) + count_ff(x.z) + count_ff(x.w);
        }
        n = __reduce_add_sync(0xffffffffu, n);
        if (lane == 0) cff[c] = (uint32_t)n;
    }
    __syncthreads();
    const int per = (n_chunks + kScanThreads - 1) / kScanThreads;
    const int i0 = min(t * per, n_chunks), i1 = min(i0 + per, n_chunks);
    uint32_t sum = 0;
    for (int i = i0; i < i1; ++i) sum += cff[i];
    uint32_t total;
    uint32_t run = block_exclusive_scan(sum, s_warp, &total);
    for (int i = i0; i < i1; ++i) {
        const uint32_t n = cff[i];
        cff[i] = run;
        run += n;
    }
    if (t == 0) {
        const int64_t need = (int64_t)kHeaderBytes + nbytes + total + 2;
        len_out[b] = need <= out_stride ? (int32_t)need : (int32_t)-need;
    }
}

constexpr int kStuffThreads = 256;

__global__ void __launch_bounds__(kStuffThreads)
jpeg_stuff_kernel(const uint32_t* __restrict__ stream, const uint32_t* __restrict__ frame_bits,
                  const uint32_t* __restrict__ chunk_ff, JpegGeom G, const __grid_constant__ JpegHeader hdr,
                  const int32_t* __restrict__ len_out, uint8_t* __restrict__ out, int64_t out_stride) {
    __shared__ uint8_t s_out[2 * kChunkBytes];
    __shared__ uint32_t s_warp[33];
    const int b = blockIdx.y, t = threadIdx.x;
    const int32_t len = len_out[b];
    if (len < 0) return;                                        // does not fit out_stride: nothing is written
    const uint32_t nbytes = (frame_bits[b] + 7) >> 3;
    const int n_chunks = (int)((nbytes + kChunkBytes - 1) / kChunkBytes);
    uint8_t* dst = out + (int64_t)b * out_stride;
    if (blockIdx.x == 0) {
        for (int i = t; i < kHeaderBytes; i += kStuffThreads) dst[i] = hdr.bytes[i];
        if (t == 0) { dst[len - 2] = 0xFF; dst[len - 1] = 0xD9; }
    }
    const uint4* st = reinterpret_cast<const uint4*>(stream + (int64_t)b * G.words_cap);
    for (int c = blockIdx.x; c < n_chunks; c += gridDim.x) {
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
        uint8_t* d = dst + kHeaderBytes + base + chunk_ff[(int64_t)b * G.chunks_cap + c];
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
    MLP_CHECK_ARG(G.mcu_rows <= 65535, "mlp_jpeg_encode: frame too tall");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DeviceGuard guard(ctx->device);

    const int64_t nb = (int64_t)batch * G.n_blk;
    auto up = [](int64_t v) { return (v + 255) & ~(int64_t)255; };
    const int64_t o_coef = 0;
    const int64_t o_meta = o_coef + up(nb * 64 * 2);
    const int64_t o_off = o_meta + up(nb * 4);
    const int64_t o_bits = o_off + up(nb * 4);
    const int64_t o_cff = o_bits + up((int64_t)batch * 4);
    const int64_t o_stream = o_cff + up((int64_t)batch * G.chunks_cap * 4);
    const int64_t bytes = o_stream + up((int64_t)batch * G.words_cap * 4);
    int rc = mlp_ensure_scratch(ctx, MLP_ARENA_JPEG, bytes);
    if (rc != MLP_OK) return rc;
    char* base = static_cast<char*>(ctx->arena[MLP_ARENA_JPEG]);
    int16_t* coefs = reinterpret_cast<int16_t*>(base + o_coef);
    uint32_t* meta = reinterpret_cast<uint32_t*>(base + o_meta);
    uint32_t* blk_off = reinterpret_cast<uint32_t*>(base + o_off);
    uint32_t* frame_bits = reinterpret_cast<uint32_t*>(base + o_bits);
    uint32_t* chunk_ff = reinterpret_cast<uint32_t*>(base + o_cff);
    uint32_t* scan = reinterpret_cast<uint32_t*>(base + o_stream);

    JpegTables T;
    JpegHeader H;
    build_tables(frame_h, frame_w, quality, &T, &H);

    ProfScope prof(ctx, MLP_ST_JPEG, stream);
    dim3 dgrid((G.mcu_cols + kMcuPerCta - 1) / kMcuPerCta, G.mcu_rows, batch);
    jpeg_dct_kernel<<<dgrid, kDctThreads, 0, stream>>>(images_dev, G, T, coefs, meta);
    MLP_LAUNCH_CHECK(ctx);
    jpeg_scan_kernel<<<batch, kScanThreads, 0, stream>>>(meta, G, T, blk_off, frame_bits, scan);
    MLP_LAUNCH_CHECK(ctx);
    const int64_t hgrid = (nb + kHuffWarps - 1) / kHuffWarps;
    jpeg_huff_kernel<<<(unsigned)(hgrid < (1 << 30) ? hgrid : (1 << 30)), kHuffWarps * 32, 0, stream>>>(
        coefs, meta, blk_off, G, T, batch, scan);
    MLP_LAUNCH_CHECK(ctx);
    jpeg_ffscan_kernel<<<batch, kScanThreads, 0, stream>>>(scan, frame_bits, G, chunk_ff, out_stride, len_dev);
    MLP_LAUNCH_CHECK(ctx);
    // typical scans take 0.3-1 byte per pixel; the chunk loop covers the rest
    int sgrid = (int)(((int64_t)frame_h * frame_w + kChunkBytes - 1) / kChunkBytes);
    sgrid = sgrid < 1 ? 1 : (sgrid > G.chunks_cap ? G.chunks_cap : sgrid);
    jpeg_stuff_kernel<<<dim3(sgrid, batch), kStuffThreads, 0, stream>>>(scan, frame_bits, chunk_ff, G, H, len_dev, out_dev,
                                                                        out_stride);
    MLP_LAUNCH_CHECK(ctx);
    return MLP_OK;
}
