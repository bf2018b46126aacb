"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/masklab_b200.h declares, the ctypes table covers them all, and the product package
never touches oracle/ (no compute call is made here - there is no GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "masklab_b200.h")
PKG = os.path.join(ROOT, "instance-segmentation-road-project_b200")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mlp_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    import masklab_b200
    return masklab_b200.load_library()


def test_header_declares_the_whole_path():
    syms = declared_symbols()
    for must in ("mlp_prior_layer", "mlp_restore_boxes", "mlp_normalize_boxes", "mlp_detection_proposal",
                 "mlp_detect_from_heads", "mlp_mask_distribute", "mlp_roi_align_plan", "mlp_roi_align_run",
                 "mlp_trim_plan", "mlp_trim_run", "mlp_upsample_output", "mlp_crop_and_pad_mask",
                 "mlp_mold_batch_plan", "mlp_mold_batch_run", "mlp_dlpack_view", "mlp_ctx_create"):
        assert must in syms


def test_library_exports_every_declared_symbol(lib):
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in masklab_b200.h but not exported"


def test_ctypes_table_matches_header(lib):
    from masklab_b200 import runtime as rt
    assert sorted(rt.SIGNATURES) == declared_symbols()
    assert lib.mlp_version() == 100
    assert isinstance(lib.mlp_last_error(), bytes)
    assert ctypes.sizeof(rt.PriorConfigC) == 4 * (2 + 8 + 8 + 2 * 8 * 32)
    assert ctypes.sizeof(rt.DetectionParamsC) == 20


def test_prior_count_is_host_side(lib):
    import masklab_b200 as ml
    import synth
    pc = ml.PriorBoxes(**synth.prior_config()).to_c("same")
    assert lib.mlp_prior_count(ctypes.byref(pc), 512, 1024) == 163680
    assert lib.mlp_prior_count(ctypes.byref(pc), 540, 960) == 163275
    pv = ml.PriorBoxes(**synth.prior_config()).to_c("valid")
    assert lib.mlp_prior_count(ctypes.byref(pv), 20, 20) == 5 * 15
    bad = ml.PriorBoxes(**synth.prior_config()).to_c("same")
    bad.num_levels = 0
    assert lib.mlp_prior_count(ctypes.byref(bad), 64, 64) < 0
    assert b"num_levels" in lib.mlp_last_error()


def test_no_gpu_means_loud_failure():
    import torch
    import masklab_b200 as ml
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError):
        ml.Context.get()
    with pytest.raises((RuntimeError, ValueError)):
        ml.DetectionProposal()([torch.zeros((1, 4, 2)), torch.zeros((1, 4, 4)), None])


def test_product_never_imports_oracle():
    offenders = []
    for base, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(base, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M) or "oracle/" in text and f.endswith(".py") and "import" in text and re.search(r"import.*oracle", text):
                    offenders.append(f)
    assert not offenders, offenders


def test_prior_boxes_mirror_matches_golden():
    import json
    import masklab_b200 as ml
    with open(os.path.join(ROOT, "tests", "golden", "prior_tables.json")) as f:
        golden = json.load(f)
    for name, g in golden.items():
        pb = ml.PriorBoxes(g["strides"], g["sizes"], g["pr_scales"], g["pr_ratios"])
        assert pb.table.tolist() == g["rows"], name
        assert len(pb) == g["num_anchors"]
        assert pb.get_config() == {"strides": g["strides"], "sizes": g["sizes"],
                                   "pr_scales": g["pr_scales"], "pr_ratios": g["pr_ratios"]}
        assert list(pb.boxes.columns) == ["stride", "w", "h"] and pb.boxes.index[0] == 1


def test_layer_configs_roundtrip_without_gpu():
    import masklab_b200 as ml
    import synth
    objs = ml.get_custom_objects()
    for name in ("PriorLayer", "RestoreBoxes", "NormalizeBoxes", "DetectionProposal", "MoldBatch",
                 "MaskDistribute", "PyramidRoiAlign", "TrimInstances", "UpSampleOutput", "CropAndPadMask"):
        assert name in objs
    layers = [ml.PriorLayer(synth.prior_config(), padding="valid"), ml.DetectionProposal(0.3, 0.5, 0.7, 50, 8),
              ml.MaskDistribute(3, 40), ml.PyramidRoiAlign((7, 7), 16), ml.TrimInstances(False, 4),
              ml.MoldBatch(12), ml.RestoreBoxes(), ml.NormalizeBoxes(), ml.UpSampleOutput(),
              ml.CropAndPadMask(output="uint8")]
    # the consumers and the steps either side of the path
    for name in ("DownSampleInput", "ResizeLike", "CrackToInstance", "SummaryOutput", "IncludeMyRoad", "CalculateInstanceSize",
                 "DrawBoxes", "DrawSegmentation", "DrawInstance", "SemanticSmoothing", "CalculateIOU", "AssignBoxes",
                 "AssignMasks", "DetectionIOUMetric", "EncodeImageContent"):
        assert name in objs
    colors = [[192, 32, 128], [160, 96, 0]]
    layers += [ml.DownSampleInput((270, 480)), ml.CrackToInstance(4), ml.SummaryOutput(2.5), ml.IncludeMyRoad(0.2),
               ml.CalculateInstanceSize(3.0), ml.DrawBoxes(), ml.DrawSegmentation(colors, 0.4),
               ml.DrawInstance(colors, 0.5), ml.UpSampleOutput(semantic=False), ml.SemanticSmoothing(6, 0.5),
               ml.CalculateIOU(), ml.AssignBoxes(7), ml.AssignMasks(0.6), ml.DetectionIOUMetric(), ml.ResizeLike(False),
               ml.EncodeImageContent(), ml.EncodeImageContent(quality=80)]
    for layer in layers:
        clone = type(layer).from_config(layer.get_config())
        assert clone.get_config() == layer.get_config()
    assert layers[0].get_config()["trainable"] is False        # detection.py:264-266


def test_entry_points_reject_bad_arguments_before_touching_the_gpu(lib):
    """Argument validation happens on the host: NULL pointers and bad shapes come back as MLP_EINVAL
    with a message, with or without a CUDA device."""
    from masklab_b200 import runtime as rt
    null = ctypes.c_void_p(None)
    cases = [
        lib.mlp_road_scan(null, null, 1, 8, 8, 3, 1, 2, 3.25, null, null, null, null, null),
        lib.mlp_summary_output(null, null, null, rt.MLP_F32, null, null, null, 1, 1, 1, null, 8, 8, 0.1, null, null, null),
        lib.mlp_tile_summary(null, null, null, null, 0, null, 0, null, 1, 1, 1, null, 28, 28, null, null, null, 8, 8,
                             0.1, null, null, null, null),
        lib.mlp_draw_boxes(null, null, rt.MLP_U8, null, 1, 1, 1, null, 8, 8, null, null),
        lib.mlp_draw_segmentation(null, null, rt.MLP_U8, null, rt.MLP_I32, 1, 8, 8, None, null, null),
        lib.mlp_draw_instance(null, null, rt.MLP_U8, null, null, rt.MLP_U8, 1, 1, 1, null, 8, 8, None, null, null),
        lib.mlp_draw_tiles(null, null, rt.MLP_U8, null, null, null, 0, null, 0, null, 1, 1, 1, null, 28, 28, 8, 8,
                           None, null, rt.MLP_I32, None, null, null),
        lib.mlp_resize_bilinear(null, null, rt.MLP_F32, 1, 4, 4, 3, 8, 8, 0, null, null),
        lib.mlp_trim_paste(null, null, null, 1, 1, null, 28, 28, 1, 1.0, 1.0, 1, 8, 8, rt.MLP_PASTE_U8, null, null,
                           null, null, null),
    ]
    assert all(rc == rt.MLP_EINVAL for rc in cases), cases
    assert b"NULL" in lib.mlp_last_error()
    with pytest.raises(rt.InvalidArgumentError):
        rt.DrawColorsC.make([[1, 2]], 0.3)                      # not an RGB triple
    with pytest.raises(rt.InvalidArgumentError):
        rt.DrawColorsC.make([[0, 0, 0]] * 17, 0.3)              # more than MLP_MAX_DRAW_CLASSES
    c = rt.DrawColorsC.make([[1, 2, 3], [4, 5, 6]], 0.25)
    assert c.num_classes == 2 and c.alpha == 0.25 and list(c.rgb[1]) == [4.0, 5.0, 6.0]
    assert ctypes.sizeof(rt.DrawColorsC) == 8 + 16 * 3 * 4


def test_jpeg_header_and_bounds_are_host_side(lib):
    """mlp_jpeg_header / mlp_jpeg_max_bytes run on the host: the 623 header bytes (SOI, JFIF 300 dpi, DQT at the
    requested quality, SOF0 4:2:0, the four Annex K DHT segments, SOS) equal the oracle's, which equals libjpeg's."""
    import ctypes
    from oracle import jpeg_oracle as jo
    buf = (ctypes.c_uint8 * 640)()
    for H, W, q in ((1080, 1920, 95), (512, 1024, 95), (37, 29, 30), (1, 1, 100), (65535, 65535, 1)):
        n = lib.mlp_jpeg_header(H, W, q, buf, 640)
        assert n == 623 and bytes(buf[:n]) == jo.headers(H, W, q).tobytes()
    assert lib.mlp_jpeg_header(0, 10, 95, buf, 640) < 0 and lib.mlp_jpeg_header(10, 70000, 95, buf, 640) < 0
    assert lib.mlp_jpeg_header(16, 16, 95, buf, 100) < 0
    assert lib.mlp_jpeg_max_bytes(16, 16) == 623 + 2 * 6 * 208 + 2 and lib.mlp_jpeg_max_bytes(0, 5) == 0
    # argument validation of the encoder happens before any CUDA call
    assert lib.mlp_jpeg_encode(None, None, 1, 16, 16, 95, None, 4096, None, None) < 0
