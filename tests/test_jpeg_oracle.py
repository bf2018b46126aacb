"""CPU pins of the JPEG restatement (oracle/jpeg_oracle.py): byte-for-byte against libjpeg-turbo's own output
(tests/golden/jpeg_golden.npz, made with Pillow by tests/golden/make_jpeg_golden.py) and, when Pillow or OpenCV
is importable where the tests run, against fresh encodes of random frames."""
import io
import os

import numpy as np
import pytest

from oracle import jpeg_oracle as jo

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "jpeg_golden.npz")


def golden_cases():
    z = np.load(GOLDEN)
    names = sorted({k.split("/")[0] for k in z.files})
    return [(n, z[n + "/rgb"], z[n + "/jpeg"].tobytes()) for n in names]


@pytest.mark.parametrize("name,rgb,want", golden_cases(), ids=[c[0] for c in golden_cases()])
def test_oracle_reproduces_libjpeg_bytes(name, rgb, want):
    assert jo.encode_jpeg(rgb) == want


def test_header_layout():
    h = jo.headers(1080, 1920).tobytes()
    assert h[:4] == b"\xff\xd8\xff\xe0" and h[6:11] == b"JFIF\0"
    assert h[13:18] == bytes([1, 1, 0x2C, 1, 0x2C])                    # density unit "in", 300 x 300
    sof = h.index(b"\xff\xc0")
    assert h[sof + 5:sof + 9] == bytes([1080 >> 8, 1080 & 255, 1920 >> 8, 1920 & 255])
    assert h[sof + 10:sof + 19] == bytes([1, 0x22, 0, 2, 0x11, 1, 3, 0x11, 1])   # 4:2:0
    assert len(h) == 623
    q = jo.quant_table(jo.STD_LUMA_Q, 95)
    assert q[:8].tolist() == [2, 1, 1, 2, 2, 4, 5, 6]


def test_zigzag_and_huffman_tables():
    assert jo.ZIGZAG[:10].tolist() == [0, 1, 8, 16, 9, 2, 3, 10, 17, 24] and jo.ZIGZAG[-1] == 63
    assert sorted(jo.ZIGZAG.tolist()) == list(range(64))
    co, si = jo.huff_table(jo.AC_LUMA_BITS, jo.AC_LUMA_VALS)
    assert (co[0x00], si[0x00]) == (0b1010, 4) and (co[0xF0], si[0xF0]) == (0b11111111001, 11)
    co, si = jo.huff_table(jo.AC_CHROMA_BITS, jo.AC_CHROMA_VALS)
    assert (co[0x00], si[0x00]) == (0b00, 2) and (co[0xF0], si[0xF0]) == (0b1111111010, 10)


def test_dc_is_the_sample_sum():
    """jfdctint's two passes leave exactly the sum of the centred samples in the DC slot (scale 8)."""
    rng = np.random.default_rng(1)
    blk = rng.integers(-128, 128, (5, 8, 8)).astype(np.int64)
    out = jo.fdct_islow(blk)
    assert np.array_equal(out[:, 0, 0], blk.sum(axis=(1, 2)))


def test_encode_image_content_takes_the_first_frame():
    rng = np.random.default_rng(2)
    frames = rng.integers(0, 256, (3, 16, 24, 3), dtype=np.uint8)
    out = jo.encode_image_content(frames)
    assert len(out) == 1 and out[0] == jo.encode_jpeg(frames[0])


def test_against_a_live_libjpeg_when_present():
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(3)
    for H, W in ((40, 56), (33, 47), (16, 130)):
        yy, xx = np.mgrid[0:H, 0:W]
        img = np.clip(np.stack([xx * 3 + yy, 255 - xx * 2, yy * 5], -1) + rng.normal(0, 20, (H, W, 3)), 0, 255).astype(np.uint8)
        buf = io.BytesIO()
        Image.fromarray(img).save(buf, format="JPEG", quality=95, dpi=(300, 300))
        assert jo.encode_jpeg(img) == buf.getvalue()
        back = np.asarray(Image.open(io.BytesIO(jo.encode_jpeg(img))).convert("RGB")).astype(int)
        assert np.abs(back - img).mean() < 20                           # it decodes to the frame it came from


def test_against_opencv_when_present():
    """A second consumer of libjpeg-turbo (OpenCV's bundled build, BGR input): the same file except for the JFIF
    density field (OpenCV writes unit 0, 1 x 1; tf.io.encode_jpeg writes unit 1 = inch, 300 x 300)."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(4)
    for H, W in ((48, 80), (35, 51)):
        yy, xx = np.mgrid[0:H, 0:W]
        img = np.clip(np.stack([xx * 3 + yy, 255 - xx * 2, yy * 5], -1) + rng.normal(0, 10, (H, W, 3)), 0, 255).astype(np.uint8)
        ok, buf = cv2.imencode(".jpg", np.ascontiguousarray(img[..., ::-1]), [cv2.IMWRITE_JPEG_QUALITY, 95])
        assert ok
        theirs, mine = bytes(buf), jo.encode_jpeg(img)
        assert len(theirs) == len(mine)
        assert theirs[:13] == mine[:13] and theirs[18:] == mine[18:]          # all but density unit / X / Y
        assert mine[13:18] == bytes([1, 0x01, 0x2C, 0x01, 0x2C]) and theirs[13:18] == bytes([0, 0, 1, 0, 1])
