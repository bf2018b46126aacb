"""CPU oracle for the MaskLab post-backbone hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU, the algorithm of the reference's Keras
layers (engine/layers/{detection,instance,misc}.py, engine/prior.py) and of the
TensorFlow kernels they delegate to.  It exists to CHECK the CUDA path; it is
never the thing shipped or measured.  Only `tests/`, `__graft_entry__.smoke()`
and the `cpu_baseline` / `--impl reference` legs of `bench.py` may import it.
The product package must never import from here (tests/test_abi.py::test_product_never_imports_oracle greps
for that).

PARITY UNPINNED.  The reference holds no tests, golden vectors or fixtures for
this path and its arithmetic lives in TensorFlow 1.14/1.15,
which is neither vendored under /root/reference nor installable here.  The TF
kernel semantics (NonMaxSuppressionV3, CropAndResize, legacy ResizeBilinear,
Where, Unique, DynamicPartition) are restated from their published algorithms;
the only reference code that executes here is engine/prior.py, whose outputs
are committed as tests/golden/prior_tables.json (made by
tests/golden/make_prior_golden.py).  Everything else is pinned by hand-derived
known-answer tests, by the known-answer vectors published with TensorFlow's
own kernel unit tests (non_max_suppression_op_test, crop_and_resize_op_test,
resize_bilinear_op_test; transcribed in tests/test_tf_published_vectors.py -
the closest published fixtures of the third-party arithmetic), by cross-checks
against independent implementations (torchvision.ops.nms, torch
interpolate/grid_sample with align_corners=True) and by a second,
independently written C restatement (oracle/c/).

One row IS pinned: oracle/jpeg_oracle.py (EncodeImageContent = tf.io.encode_jpeg
defaults, i.e. libjpeg baseline) reproduces byte for byte the files libjpeg-turbo
itself wrote for tests/golden/jpeg_golden.npz (made with Pillow by
tests/golden/make_jpeg_golden.py).

Two documented deviations from "whatever TF does", both fixed by
BASELINE.json's north_star or forced by bit-reproducibility:
  * NMS score ties pop the lower candidate index first (TF>=2.2 comparator;
    TF 1.x is heap-order dependent).
  * exp() in box decoding and log() in FPN-level assignment are the correctly
    rounded float32 values (computed in float64, rounded once).  TF's Eigen
    pexp/plog differ from that by <=1 ulp; so do glibc and CUDA libm.  Choosing
    the correctly rounded value makes CPU and GPU agree bit for bit.
"""
