"""Golden JPEG byte streams for oracle/jpeg_oracle.py and the CUDA encoder.

Run in the build container (Pillow 12.2, bundled libjpeg-turbo):  python tests/golden/make_jpeg_golden.py
`Image.save(format="JPEG", quality=95, dpi=(300, 300))` configures libjpeg exactly like TensorFlow's
`tf.io.encode_jpeg` defaults do (jpeg_set_defaults, quality 95 baseline, 4:2:0, ISLOW DCT, standard Huffman tables,
JFIF density 300x300 dpi) - the call behind /root/reference/engine/layers/misc.py:348.  Inputs and outputs are stored
together so that nothing has to be regenerated on the test side.
"""
import io
import os

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = [  # (name, H, W, kind)
    ("mcu_aligned_smooth", 32, 48, "smooth"),
    ("mcu_aligned_noise", 32, 48, "noise"),
    ("odd_smooth", 37, 29, "smooth"),          # dummy Y blocks right and below, replicated chroma edges
    ("odd_noise", 37, 29, "noise"),
    ("one_pixel", 1, 1, "noise"),
    ("one_block", 8, 8, "noise"),
    ("wide_strip", 9, 250, "smooth"),
    ("tall_strip", 120, 7, "overlay"),
    ("flat_black", 24, 40, "black"),           # every AC zero: EOB-only blocks
    ("flat_white", 16, 16, "white"),
    ("saturated_checker", 48, 64, "checker"),  # largest coefficient magnitudes, long codes, 0xFF stuffing
    ("overlay_like", 96, 160, "overlay"),      # smooth frame + flat colour regions + one-pixel white lines
    ("half_1080_rows", 72, 64, "smooth"),      # H = 4.5 MCU rows, like 1080 = 67.5
]


def make_image(kind, H, W, rng):
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float64)
    if kind == "noise":
        return rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    if kind == "black":
        return np.zeros((H, W, 3), dtype=np.uint8)
    if kind == "white":
        return np.full((H, W, 3), 255, dtype=np.uint8)
    if kind == "checker":
        c = (((yy.astype(int) // 1) + (xx.astype(int) // 1)) % 2 * 255).astype(np.uint8)
        return np.stack([c, 255 - c, c], -1)
    img = np.stack([128 + 100 * np.sin(xx / 9.0 + yy / 17.0), 128 + 90 * np.cos(xx / 5.0), 100 + yy * 0.7 + xx * 0.3], -1)
    img += rng.normal(0, 6, img.shape)
    img = np.clip(img, 0, 255).astype(np.uint8)
    if kind == "overlay":
        img[H // 4:H // 2, W // 5:W // 2] = (img[H // 4:H // 2, W // 5:W // 2] * 0.7 + np.array([57, 9, 38])).astype(np.uint8)
        img[H // 3, :] = 255
        img[:, W // 3] = 255
    return img


def main():
    rng = np.random.default_rng(20261018)
    out = {}
    for name, H, W, kind in CASES:
        img = make_image(kind, H, W, rng)
        buf = io.BytesIO()
        Image.fromarray(img).save(buf, format="JPEG", quality=95, dpi=(300, 300))
        out[name + "/rgb"] = img
        out[name + "/jpeg"] = np.frombuffer(buf.getvalue(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "jpeg_golden.npz"), **out)
    print("wrote", len(CASES), "cases,", os.path.getsize(os.path.join(HERE, "jpeg_golden.npz")), "bytes")


if __name__ == "__main__":
    main()
