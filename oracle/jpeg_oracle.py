"""NumPy restatement of the JPEG encode at the end of the serving graph ("overlay +
JPEG encode").  TEST INFRASTRUCTURE — see oracle/__init__.py.

Reference: /root/reference/engine/layers/misc.py:343-351 (EncodeImageContent.call =
`tf.io.encode_jpeg(inputs[0])` with every attribute at its default), wired in
/root/reference/road_project/setup/serving.py:41.  The arithmetic lives in a third-party dependency that is not
under /root/reference: TensorFlow's `EncodeJpeg` kernel (core/kernels/encode_jpeg_op.cc -> core/lib/jpeg/jpeg_mem.cc)
drives libjpeg(-turbo) with `jpeg_set_defaults`, `jpeg_set_quality(95, force_baseline=TRUE)`, in_color_space RGB,
default 4:2:0 chroma subsampling, default JDCT_ISLOW, the standard (Annex K) Huffman tables (optimize_size=False),
baseline sequential (progressive=False), JFIF density unit 1 ("in") 300x300, no comment / XMP.  This file restates
libjpeg's baseline compressor for exactly that parameter set:

  jccolor.c  rgb_ycc_convert   16-bit fixed-point RGB -> YCbCr
  jcprepct.c / jcsample.c      edge replication, fullsize (Y) and h2v2 (Cb, Cr; alternating 1,2 bias) down-sampling
  jcdctmgr.c / jfdctint.c      sample - 128, jpeg_fdct_islow (CONST_BITS 13, PASS1_BITS 2), divide by 8*q rounding
                               half away from zero
  jccoefct.c compress_data     MCU order, dummy blocks right / below the image (zero AC, DC of the previous block)
  jchuff.c   encode_one_block  DC differences, run/size symbols with ZRL and EOB, 0xFF byte stuffing, 1-padding
  jcmarker.c                   SOI, APP0 JFIF 1.01, two DQT, SOF0, four DHT, SOS, EOI

PARITY PINNED (unlike the rest of the oracle): Pillow in this container links libjpeg-turbo, the library family
TensorFlow itself links, and `Image.save(format="JPEG", quality=95, dpi=(300, 300))` sets the same parameters.  The
byte streams it produces are committed under tests/golden/jpeg_golden.npz (made by tests/golden/make_jpeg_golden.py)
and this restatement reproduces every one of them byte for byte (tests/test_jpeg_oracle.py).
"""
import numpy as np

# --- Annex K tables (jcparam.c std_luminance_quant_tbl / std_chrominance_quant_tbl, std_huff_tables) ----------------
STD_LUMA_Q = np.array([
    16, 11, 10, 16, 24, 40, 51, 61,
    12, 12, 14, 19, 26, 58, 60, 55,
    14, 13, 16, 24, 40, 57, 69, 56,
    14, 17, 22, 29, 51, 87, 80, 62,
    18, 22, 37, 56, 68, 109, 103, 77,
    24, 35, 55, 64, 81, 104, 113, 92,
    49, 64, 78, 87, 103, 121, 120, 101,
    72, 92, 95, 98, 112, 100, 103, 99], dtype=np.int64)
STD_CHROMA_Q = np.array([
    17, 18, 24, 47, 99, 99, 99, 99,
    18, 21, 26, 66, 99, 99, 99, 99,
    24, 26, 56, 99, 99, 99, 99, 99,
    47, 66, 99, 99, 99, 99, 99, 99,
    99, 99, 99, 99, 99, 99, 99, 99,
    99, 99, 99, 99, 99, 99, 99, 99,
    99, 99, 99, 99, 99, 99, 99, 99,
    99, 99, 99, 99, 99, 99, 99, 99], dtype=np.int64)

DC_LUMA_BITS = [0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0]
DC_CHROMA_BITS = [0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0]
DC_VALS = list(range(12))
AC_LUMA_BITS = [0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d]
AC_LUMA_VALS = [
    0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07,
    0x22, 0x71, 0x14, 0x32, 0x81, 0x91, 0xa1, 0x08, 0x23, 0x42, 0xb1, 0xc1, 0x15, 0x52, 0xd1, 0xf0,
    0x24, 0x33, 0x62, 0x72, 0x82, 0x09, 0x0a, 0x16, 0x17, 0x18, 0x19, 0x1a, 0x25, 0x26, 0x27, 0x28,
    0x29, 0x2a, 0x34, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49,
    0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69,
    0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89,
    0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7,
    0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5,
    0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe1, 0xe2,
    0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8,
    0xf9, 0xfa]
AC_CHROMA_BITS = [0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77]
AC_CHROMA_VALS = [
    0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71,
    0x13, 0x22, 0x32, 0x81, 0x08, 0x14, 0x42, 0x91, 0xa1, 0xb1, 0xc1, 0x09, 0x23, 0x33, 0x52, 0xf0,
    0x15, 0x62, 0x72, 0xd1, 0x0a, 0x16, 0x24, 0x34, 0xe1, 0x25, 0xf1, 0x17, 0x18, 0x19, 0x1a, 0x26,
    0x27, 0x28, 0x29, 0x2a, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48,
    0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68,
    0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x82, 0x83, 0x84, 0x85, 0x86, 0x87,
    0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5,
    0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3,
    0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda,
    0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8,
    0xf9, 0xfa]


def zigzag_order():
    """jutils.c jpeg_natural_order: zigzag position -> natural (row-major) index."""
    order = []
    for s in range(15):
        diag = [(i, s - i) for i in range(8) if 0 <= s - i < 8]
        if s % 2 == 0:
            diag.reverse()                     # even diagonals run bottom-left -> top-right
        order += [r * 8 + c for r, c in diag]
    return np.array(order, dtype=np.int64)


ZIGZAG = zigzag_order()


def quant_table(std, quality):
    """jcparam.c jpeg_quality_scaling + jpeg_add_quant_table(force_baseline=TRUE)."""
    quality = min(max(int(quality), 1), 100)
    scale = 5000 // quality if quality < 50 else 200 - quality * 2
    return np.clip((std * scale + 50) // 100, 1, 255)


def huff_table(bits, vals):
    """jchuff.c jpeg_make_c_derived_tbl: symbol -> (code, length)."""
    code, k = 0, 0
    ehufco = np.zeros(256, dtype=np.int64)
    ehufsi = np.zeros(256, dtype=np.int64)
    for length in range(1, 17):
        for _ in range(bits[length - 1]):
            ehufco[vals[k]] = code
            ehufsi[vals[k]] = length
            code += 1
            k += 1
        code <<= 1
    return ehufco, ehufsi


def rgb_to_ycc(rgb):
    """jccolor.c rgb_ycc_start/rgb_ycc_convert: SCALEBITS 16, FIX(x) = int(x * 65536 + 0.5)."""
    def fix(x):
        return int(x * 65536 + 0.5)
    r = rgb[..., 0].astype(np.int64)
    g = rgb[..., 1].astype(np.int64)
    b = rgb[..., 2].astype(np.int64)
    half = 1 << 15
    off = 128 << 16
    y = (fix(0.29900) * r + fix(0.58700) * g + fix(0.11400) * b + half) >> 16
    cb = (-fix(0.16874) * r - fix(0.33126) * g + fix(0.50000) * b + off + half - 1) >> 16
    cr = (fix(0.50000) * r - fix(0.41869) * g - fix(0.08131) * b + off + half - 1) >> 16
    return y, cb, cr


def _pad_edge(plane, rows, cols):
    """expand_right_edge / expand_bottom_edge: replicate the last column / row."""
    h, w = plane.shape
    return np.pad(plane, ((0, rows - h), (0, cols - w)), mode="edge")


def h2v2_downsample(plane, out_rows, out_cols):
    """jcsample.c h2v2_downsample on the edge-replicated plane; rows below the last real row group replicate the
    last DOWN-SAMPLED row (jcprepct.c pads the output buffer, not the input)."""
    h, w = plane.shape
    real_rows = (h + 1) // 2
    p = _pad_edge(plane, real_rows * 2, out_cols * 2)
    bias = np.tile(np.array([1, 2], dtype=np.int64), out_cols)[:out_cols]
    s = p[0::2, 0::2] + p[0::2, 1::2] + p[1::2, 0::2] + p[1::2, 1::2] + bias[None, :]
    return _pad_edge(s >> 2, out_rows, out_cols)


def fdct_islow(block):
    """jfdctint.c jpeg_fdct_islow on an [..., 8, 8] int64 array of centred samples; output scaled by 8."""
    C_BITS, P1 = 13, 2
    F_0_298, F_0_390, F_0_541, F_0_765 = 2446, 3196, 4433, 6270
    F_0_899, F_1_175, F_1_501, F_1_847 = 7373, 9633, 12299, 15137
    F_1_961, F_2_053, F_2_562, F_3_072 = 16069, 16819, 20995, 25172

    def descale(x, n):
        return (x + (1 << (n - 1))) >> n

    def one_pass(d, first):
        # d[..., k] is the k-th element along the transformed axis
        t0, t7 = d[..., 0] + d[..., 7], d[..., 0] - d[..., 7]
        t1, t6 = d[..., 1] + d[..., 6], d[..., 1] - d[..., 6]
        t2, t5 = d[..., 2] + d[..., 5], d[..., 2] - d[..., 5]
        t3, t4 = d[..., 3] + d[..., 4], d[..., 3] - d[..., 4]
        t10, t13 = t0 + t3, t0 - t3
        t11, t12 = t1 + t2, t1 - t2
        out = [None] * 8
        if first:
            out[0] = (t10 + t11) << P1
            out[4] = (t10 - t11) << P1
            sh = C_BITS - P1
        else:
            out[0] = descale(t10 + t11, P1)
            out[4] = descale(t10 - t11, P1)
            sh = C_BITS + P1
        z1 = (t12 + t13) * F_0_541
        out[2] = descale(z1 + t13 * F_0_765, sh)
        out[6] = descale(z1 + t12 * (-F_1_847), sh)
        z1, z2, z3, z4 = t4 + t7, t5 + t6, t4 + t6, t5 + t7
        z5 = (z3 + z4) * F_1_175
        t4, t5, t6, t7 = t4 * F_0_298, t5 * F_2_053, t6 * F_3_072, t7 * F_1_501
        z1, z2, z3, z4 = z1 * (-F_0_899), z2 * (-F_2_562), z3 * (-F_1_961), z4 * (-F_0_390)
        z3, z4 = z3 + z5, z4 + z5
        out[7] = descale(t4 + z1 + z3, sh)
        out[5] = descale(t5 + z2 + z4, sh)
        out[3] = descale(t6 + z2 + z3, sh)
        out[1] = descale(t7 + z1 + z4, sh)
        return np.stack(out, axis=-1)

    rows = one_pass(block, True)                                        # pass 1: along each row
    cols = one_pass(np.swapaxes(rows, -1, -2), False)                   # pass 2: along each column
    return np.swapaxes(cols, -1, -2)


def quantize(work, qtbl):
    """jcdctmgr.c forward_DCT: divisor = q << 3, round half away from zero."""
    div = (qtbl << 3).reshape(8, 8)
    mag = (np.abs(work) + (div >> 1)) // div
    return np.where(work < 0, -mag, mag)


def component_blocks(plane, qtbl):
    """[rows, cols] samples (multiples of 8) -> quantised coefficients [rows/8, cols/8, 64] in zigzag order."""
    r, c = plane.shape
    blocks = plane.reshape(r // 8, 8, c // 8, 8).transpose(0, 2, 1, 3) - 128
    q = quantize(fdct_islow(blocks), qtbl)
    return q.reshape(r // 8, c // 8, 64)[..., ZIGZAG]


def mcu_coefficients(rgb, quality=95):
    """uint8 [H, W, 3] -> int64 [n_mcu, 6, 64] zigzag coefficients in scan order (Y00 Y01 Y10 Y11 Cb Cr per MCU),
    dummy blocks included (jccoefct.c compress_data)."""
    H, W, _ = rgb.shape
    y, cb, cr = rgb_to_ycc(rgb)
    mcu_rows, mcu_cols = -(-H // 16), -(-W // 16)
    yb_rows, yb_cols = -(-H // 8), -(-W // 8)                            # real block rows / cols of Y
    cw, ch = -(-W // 2), -(-H // 2)
    cb_rows, cb_cols = -(-ch // 8), -(-cw // 8)
    assert cb_rows == mcu_rows and cb_cols == mcu_cols
    qy, qc = quant_table(STD_LUMA_Q, quality), quant_table(STD_CHROMA_Q, quality)
    Y = component_blocks(_pad_edge(y, yb_rows * 8, yb_cols * 8), qy)
    CB = component_blocks(h2v2_downsample(cb, cb_rows * 8, cb_cols * 8), qc)
    CR = component_blocks(h2v2_downsample(cr, cb_rows * 8, cb_cols * 8), qc)
    out = np.zeros((mcu_rows, mcu_cols, 6, 64), dtype=np.int64)
    for my in range(mcu_rows):
        for mx in range(mcu_cols):
            prev_dc = 0
            for yi in range(2):
                for xi in range(2):
                    by, bx = my * 2 + yi, mx * 2 + xi
                    k = yi * 2 + xi
                    if by < yb_rows and bx < yb_cols:
                        out[my, mx, k] = Y[by, bx]
                    else:
                        out[my, mx, k, 0] = prev_dc                    # dummy block: DC of the previous block
                    prev_dc = out[my, mx, k, 0]
            out[my, mx, 4] = CB[my, mx]
            out[my, mx, 5] = CR[my, mx]
    return out.reshape(mcu_rows * mcu_cols, 6, 64)


def _nbits(v):
    a = np.abs(v)
    n = np.zeros(a.shape, dtype=np.int64)
    for k in range(16):
        n += (a >> k) > 0
    return n


def entropy_items(coefs):
    """[n_mcu, 6, 64] -> (codes, lengths) of every bit field of the scan, in order (jchuff.c encode_one_block)."""
    n_mcu = coefs.shape[0]
    dc_l, ac_l = huff_table(DC_LUMA_BITS, DC_VALS), huff_table(AC_LUMA_BITS, AC_LUMA_VALS)
    dc_c, ac_c = huff_table(DC_CHROMA_BITS, DC_VALS), huff_table(AC_CHROMA_BITS, AC_CHROMA_VALS)
    blocks = coefs.reshape(n_mcu * 6, 64).copy()
    comp = np.tile(np.array([0, 0, 0, 0, 1, 2]), n_mcu)                 # component of every block
    # DC differences per component, in scan order
    for c in range(3):
        idx = np.nonzero(comp == c)[0]
        dc = blocks[idx, 0]
        blocks[idx, 0] = dc - np.concatenate([[0], dc[:-1]])
    chroma = comp > 0
    nb = _nbits(blocks)
    val_bits = np.where(blocks < 0, blocks - 1, blocks) & ((1 << nb) - 1)
    # item list: sort key (block, position, sub) ; position 0 = DC, 1..63 = AC, 64 = EOB
    keys, codes, lens = [], [], []

    def add(block_idx, pos, sub, code, length):
        keys.append(block_idx * 65 * 8 + pos * 8 + sub)
        codes.append(code)
        lens.append(length)

    b_all = np.arange(blocks.shape[0])
    dsz = nb[:, 0]
    add(b_all, 0, 0, np.where(chroma, dc_c[0][dsz], dc_l[0][dsz]), np.where(chroma, dc_c[1][dsz], dc_l[1][dsz]))
    add(b_all, 0, 1, val_bits[:, 0], dsz)
    bi, pos = np.nonzero(blocks[:, 1:])
    pos = pos + 1
    # previous non-zero position inside the block (0 = the DC slot)
    prev = np.zeros_like(pos)
    if pos.size:
        same = np.concatenate([[False], bi[1:] == bi[:-1]])
        prev[same] = pos[:-1][same[1:]]
    run = pos - prev - 1
    ch = chroma[bi]
    for z in range(3):                                                   # up to three ZRL (0xF0) symbols
        m = (run >> 4) > z
        add(bi[m], pos[m], z, np.where(ch[m], ac_c[0][0xF0], ac_l[0][0xF0]),
            np.where(ch[m], ac_c[1][0xF0], ac_l[1][0xF0]))
    sym = ((run & 15) << 4) | nb[bi, pos]
    add(bi, pos, 3, np.where(ch, ac_c[0][sym], ac_l[0][sym]), np.where(ch, ac_c[1][sym], ac_l[1][sym]))
    add(bi, pos, 4, val_bits[bi, pos], nb[bi, pos])
    eob = np.nonzero(blocks[:, 63] == 0)[0]
    add(eob, 64, 0, np.where(chroma[eob], ac_c[0][0], ac_l[0][0]), np.where(chroma[eob], ac_c[1][0], ac_l[1][0]))
    keys = np.concatenate(keys)
    codes = np.concatenate(codes)
    lens = np.concatenate(lens)
    order = np.argsort(keys, kind="stable")
    return codes[order], lens[order]


def pack_bits(codes, lens):
    """emit_bits + flush_bits: MSB first, padded with 1-bits to a byte, 0x00 stuffed after every 0xFF."""
    keep = lens > 0
    codes, lens = codes[keep], lens[keep]
    total = int(lens.sum())
    start = np.cumsum(lens) - lens
    item = np.repeat(np.arange(lens.size), lens)
    j = np.arange(total) - start[item]
    bits = ((codes[item] >> (lens[item] - 1 - j)) & 1).astype(np.uint8)
    pad = (-total) % 8
    bits = np.concatenate([bits, np.ones(pad, dtype=np.uint8)])
    data = np.packbits(bits)
    ff = np.nonzero(data == 0xFF)[0]
    return np.insert(data, ff + 1, 0).astype(np.uint8)


def headers(H, W, quality=95, density_unit=1, x_density=300, y_density=300):
    """jcmarker.c write_file_header + write_frame_header + write_scan_header for this parameter set."""
    def be16(v):
        return [(v >> 8) & 0xFF, v & 0xFF]
    out = [0xFF, 0xD8]
    out += [0xFF, 0xE0] + be16(16) + [0x4A, 0x46, 0x49, 0x46, 0x00, 1, 1, density_unit] + be16(x_density) \
        + be16(y_density) + [0, 0]
    for idx, std in enumerate((STD_LUMA_Q, STD_CHROMA_Q)):
        out += [0xFF, 0xDB] + be16(67) + [idx] + quant_table(std, quality)[ZIGZAG].tolist()
    out += [0xFF, 0xC0] + be16(17) + [8] + be16(H) + be16(W) + [3, 1, 0x22, 0, 2, 0x11, 1, 3, 0x11, 1]
    for tc_th, bits, vals in ((0x00, DC_LUMA_BITS, DC_VALS), (0x10, AC_LUMA_BITS, AC_LUMA_VALS),
                              (0x01, DC_CHROMA_BITS, DC_VALS), (0x11, AC_CHROMA_BITS, AC_CHROMA_VALS)):
        out += [0xFF, 0xC4] + be16(2 + 1 + 16 + len(vals)) + [tc_th] + list(bits) + list(vals)
    out += [0xFF, 0xDA] + be16(12) + [3, 1, 0x00, 2, 0x11, 3, 0x11, 0, 63, 0]
    return np.array(out, dtype=np.uint8)


def encode_jpeg(rgb, quality=95):
    """tf.io.encode_jpeg(image) with default attributes: uint8 [H, W, 3] -> bytes."""
    rgb = np.asarray(rgb)
    assert rgb.dtype == np.uint8 and rgb.ndim == 3 and rgb.shape[2] == 3
    H, W, _ = rgb.shape
    coefs = mcu_coefficients(rgb, quality)
    scan = pack_bits(*entropy_items(coefs))
    return np.concatenate([headers(H, W, quality), scan, np.array([0xFF, 0xD9], dtype=np.uint8)]).tobytes()


def encode_image_content(images, quality=95):
    """misc.py:347-351: encodes images[0] only and returns it as a one-element list (string tensor of shape [1])."""
    return [encode_jpeg(np.asarray(images)[0], quality)]
