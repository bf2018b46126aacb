#!/bin/bash
# Memory-safety check without compute-sanitizer (closed on this pool): the library built with -DMLP_DEBUG_BOUNDS
# (every checked scratch / shared-memory index traps with its source line) under the small- and odd-shape GPU tests.
# Run on a GPU box from the repo root:  bash tools/run_debug_bounds.sh  > profiles/debug_bounds_r02.txt
set -u
cd "$(dirname "$0")/.."
python -m masklab_b200.build --debug-bounds 2>/dev/null | tail -1
export MASKLAB_B200_LIB="$PWD/instance-segmentation-road-project_b200/libmasklab_b200_dbg.so"
python - <<'PY'
import masklab_b200.runtime as rt
print("library under test:", rt.library_path())
PY
LOG=$(mktemp)
timeout 1500 python -m pytest tests/test_gpu_random.py tests/test_gpu_layers.py tests/test_gpu_round2.py tests/test_gpu_stress.py \
    tests/test_gpu_tf_published_vectors.py -q -x > "$LOG" 2>&1
tail -5 "$LOG"
echo "MLP_BOUND violations reported by the kernels: $(grep -c 'MLP_BOUND violated' "$LOG")"
rm -f "$LOG"
