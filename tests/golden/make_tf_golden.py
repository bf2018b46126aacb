"""Generate golden vectors for the WHOLE hot path from the reference's own Keras layers under TensorFlow 1.x.

TensorFlow 1.14 / 1.15 (the version the reference's API use implies) cannot be installed in the
build container (Python 3.12, no network), so this script could NOT be run there: the fixtures it writes are
absent from the repository and `tests/test_tf_golden.py` SKIPS (it does not pass) until they exist.  Run it where
TF 1.x is available - e.g. a python3.6/3.7 virtualenv with `pip install tensorflow==1.15.5 numpy pandas` - from
a checkout of this repository next to a checkout of the reference:

    python tests/golden/make_tf_golden.py --reference /path/to/instance-segmentation-road-project

It feeds the seeded synthetic inputs of tests/synth.py (the same the parity tests use) through the reference's
layers in the reference's wiring -

    PriorLayer -> RestoreBoxes -> DetectionProposal -> MaskDistribute -> PyramidRoiAlign
                                                  (engine/retinamasklab.py:431-469)
    [synthetic mask-head output, tests/synth.mask_probs]
    TrimInstances(mold=True) -> UpSampleOutput      (engine/retinamasklab.py:615-616, 635-636)
    CropAndPadMask                                  (road_project/setup/serving.py:30)
    SummaryOutput                                   (road_project/setup/serving.py:47-48)

- and stores every intermediate in tests/golden/tf_golden_<case>.npz.  Large tensors are stored losslessly but
small: RoI features as float32 for the first 8 RoIs plus a SHA-256 of the whole tensor's bytes, the pasted
[B,M,PH,PW] masks as packed bits of (value > 0.5) plus the SHA-256 of the float32 bytes (np.savez_compressed).

Only `engine.layers` and `engine.prior` of the reference are imported (through a stub `engine` package, so that
engine/__init__.py - which drags in every backbone and its third-party dependencies - does not run); nothing of
the reference is copied here.
"""
import argparse
import hashlib
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
TESTS = os.path.dirname(HERE)
sys.path.insert(0, TESTS)

# name -> shapes and hyper-parameters; inputs come from tests/synth.py with these seeds
CASES = {
    # small enough for the NumPy oracle in seconds; ragged counts, one image without detections
    "tiny": dict(B=3, H=128, W=256, PH=256, PW=512, C=3, Cf=16, mu=-4.0, seed=11, empty_image=1,
                 min_confidence=0.05, nms_iou_threshold=0.4, post_iou_threshold=0.65, nms_max_output_size=50,
                 max_k=2, base_size=36),
    # the default DetectionProposal ctor values (detection.py:469-473) on stress-like scores
    "ctor_defaults": dict(B=2, H=96, W=160, PH=96, PW=160, C=6, Cf=8, mu=-2.5, seed=21, empty_image=None,
                          min_confidence=0.05, nms_iou_threshold=0.4, post_iou_threshold=0.65,
                          nms_max_output_size=1000, max_k=2, base_size=64),
    # one frame of BASELINE.json configs[1] (cfg-2 shapes, ResNeXt default ModelConfiguration)
    "cfg2_frame": dict(B=1, H=512, W=1024, PH=512, PW=1024, C=5, Cf=128, mu=-5.8, seed=31, empty_image=None,
                       min_confidence=0.05, nms_iou_threshold=0.4, post_iou_threshold=0.6, nms_max_output_size=100,
                       max_k=2, base_size=36),
    # the serving shape: model at 540x960, frame 1080x1920 (BASELINE.json configs[0] / [4])
    "serving_frame": dict(B=1, H=540, W=960, PH=1080, PW=1920, C=5, Cf=32, mu=-5.8, seed=41, empty_image=None,
                          min_confidence=0.05, nms_iou_threshold=0.4, post_iou_threshold=0.6,
                          nms_max_output_size=100, max_k=2, base_size=36),
}


def case_inputs(c):
    """The seeded inputs of one case (NumPy, shared with tests/test_tf_golden.py)."""
    import synth
    cfgp = synth.prior_config()
    N = synth.num_anchors(cfgp, c["H"], c["W"])
    loc, cls = synth.head_tensors(c["B"], N, c["C"], mu=c["mu"], seed=c["seed"])
    if c["empty_image"] is not None:
        cls[c["empty_image"]] = 0
    fmaps = synth.fpn_maps(c["B"], c["H"], c["W"], c["Cf"], seed=c["seed"] + 1)
    seg = synth.semantic_map(c["B"], c["PH"], c["PW"], seed=c["seed"] + 3)
    return cfgp, N, loc, cls, fmaps, seg


def case_masks(c, R):
    import synth
    return synth.mask_probs(c["B"], R, c["C"], seed=c["seed"] + 2)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def import_reference_layers(ref_root):
    """engine.layers / engine.prior of the reference without running engine/__init__.py."""
    if not hasattr(np, "int"):
        np.int = int                     # engine/prior.py:60-66 uses the alias NumPy 1.24 removed
    pkg = types.ModuleType("engine")
    pkg.__path__ = [os.path.join(ref_root, "engine")]
    sys.modules["engine"] = pkg
    import engine.layers as L            # noqa: E402  (detection, semantic, instance, misc)
    from engine.prior import PriorBoxes  # noqa: E402
    return L, PriorBoxes


def run_case(name, c, L, PriorBoxes, tf, out_dir):
    K = tf.keras.backend
    K.clear_session()
    sess = tf.compat.v1.Session()
    K.set_session(sess)
    ph = tf.compat.v1.placeholder
    cfgp, N, loc, cls, fmaps, seg = case_inputs(c)
    B, H, W, PH, PW, C = c["B"], c["H"], c["W"], c["PH"], c["PW"], c["C"]

    # ---- first half: retinamasklab.py:431-469
    images = ph(tf.uint8, [None, H, W, 3])
    loc_ph, cls_ph = ph(tf.float32, [None, N, 4]), ph(tf.float32, [None, N, C])
    fmap_ph = [ph(tf.float32, [None] + list(f.shape[1:])) for f in fmaps]
    prior = PriorBoxes(**cfgp)
    pr_boxes = L.PriorLayer(prior)(images)
    restored = L.RestoreBoxes()([loc_ph, pr_boxes])
    proposed = L.DetectionProposal(min_confidence=c["min_confidence"], nms_iou_threshold=c["nms_iou_threshold"],
                                   post_iou_threshold=c["post_iou_threshold"],
                                   nms_max_output_size=c["nms_max_output_size"])([cls_ph, restored, images])
    dist = L.MaskDistribute(max_k=c["max_k"], base_size=c["base_size"])(proposed)
    roi_fmaps, roi_boxes = L.PyramidRoiAlign()([fmap_ph[:c["max_k"] + 1], dist, images])
    feed = {images: np.zeros((B, H, W, 3), np.uint8), loc_ph: loc, cls_ph: cls}
    feed.update({p: f for p, f in zip(fmap_ph, fmaps)})
    v_prior, v_restored, v_prop, v_dist, v_rf, v_rb = sess.run(
        [pr_boxes, restored, proposed, dist, roi_fmaps, roi_boxes], feed)

    # ---- second half: retinamasklab.py:615-616, 635-636; serving.py:30, 47-48
    R = v_rb.shape[1]
    masks = case_masks(c, R)
    frames = ph(tf.uint8, [None, PH, PW, 3])
    rb_ph, rm_ph = ph(tf.float32, [None, None, 6]), ph(tf.float32, [None, None, 28, 28, C])
    # UpSampleOutput resizes its third input to the frame; feed the semantic map at model resolution so that
    # ratio = frame / model size as in the serving graph
    sem_ph = ph(tf.float32, [None, H, W, seg.shape[-1]])
    det, ins = L.TrimInstances(mold=True)([rb_ph, rm_ph])
    det_i, ins_i, sem_i = L.UpSampleOutput()([det, ins, sem_ph], target=frames)
    pasted = L.CropAndPadMask()([frames, det_i, ins_i, sem_i])
    seg_ph = ph(tf.int32, [None, PH, PW, seg.shape[-1]])
    summary = L.SummaryOutput()([det_i, seg_ph, pasted])
    sem_small = np.zeros((B, H, W, seg.shape[-1]), np.float32)
    feed2 = {frames: np.zeros((B, PH, PW, 3), np.uint8), rb_ph: v_rb, rm_ph: masks, sem_ph: sem_small, seg_ph: seg}
    v_det, v_ins, v_det_i, v_ins_i, v_pasted, v_summary = sess.run([det, ins, det_i, ins_i, pasted, summary], feed2)

    out = dict(
        case=name, tf_version=tf.__version__, numpy_version=np.__version__,
        prior=v_prior[0].astype(np.int32), restored_sha=sha(v_restored.astype(np.float32)),
        restored_head=v_restored[:, :4096].astype(np.float32),
        proposed=v_prop.astype(np.float32), dist=v_dist.astype(np.float32), roi_boxes=v_rb.astype(np.float32),
        trim_boxes=v_det.astype(np.float32), trim_masks_sha=sha(v_ins.astype(np.float32)),
        det_i=v_det_i.astype(np.int32), ins_i_bits=np.packbits(v_ins_i.astype(bool), axis=-1, bitorder="little"),
        pasted_bits=np.packbits(v_pasted > 0.5, axis=-1, bitorder="little"),
        pasted_sha=sha(v_pasted.astype(np.float32)), pasted_sum=v_pasted.astype(np.float64).sum(axis=(2, 3)),
        summary=v_summary.astype(np.float32))
    for f, a in enumerate(v_rf):
        out[f"roi_fmaps{f}_head"] = a[:, :8].astype(np.float32)
        out[f"roi_fmaps{f}_sha"] = sha(a.astype(np.float32))
        out[f"roi_fmaps{f}_shape"] = np.array(a.shape, np.int64)
    path = os.path.join(out_dir, f"tf_golden_{name}.npz")
    np.savez_compressed(path, **out)
    print(f"{name}: M={v_prop.shape[1]} R={R} -> {path} ({os.path.getsize(path) / 1e6:.2f} MB)")


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--reference", default="/root/reference", help="checkout of the reference repository")
    ap.add_argument("--out", default=HERE)
    ap.add_argument("--cases", nargs="*", default=sorted(CASES))
    ap.add_argument("--describe", action="store_true", help="list the cases and exit (needs no TensorFlow)")
    args = ap.parse_args()
    if args.describe:
        for k in args.cases:
            print(k, CASES[k])
        return 0
    try:
        import tensorflow as tf
    except ImportError:
        print("TensorFlow is not installed here; run this script where TF 1.14/1.15 is (see the docstring).",
              file=sys.stderr)
        return 2
    if not tf.__version__.startswith("1."):
        print(f"TensorFlow {tf.__version__}: the reference needs 1.14/1.15 (tf.log, K.get_session, ...).",
              file=sys.stderr)
        return 2
    L, PriorBoxes = import_reference_layers(args.reference)
    for k in args.cases:
        run_case(k, CASES[k], L, PriorBoxes, tf, args.out)
    return 0


if __name__ == "__main__":
    sys.exit(main())
