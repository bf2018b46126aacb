"""Build libmasklab_b200.so in-tree with nvcc for sm_100a.

`python -m masklab_b200.build` or `__graft_entry__.build()`.  The .so is git-ignored
but travels to the GPU box with the gpurun snapshot.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(ROOT, "include")
LIB_NAME = "libmasklab_b200.so"
LIB_PATH = os.path.join(HERE, LIB_NAME)

SOURCES = ["ctx.cu", "elementwise.cu", "detect.cu", "roi.cu", "paste.cu", "mold.cu", "summary.cu", "draw.cu", "resize.cu", "assign.cu", "jpeg.cu", "clip.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    # parity: multiply and add round separately, exactly like TF/NumPy
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true",
    "-Xcompiler", "-fPIC", "-shared",
]


def find_nvcc():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; libmasklab_b200.so cannot be built")
    return nvcc


def sources():
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def is_stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = sources() + [os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "paste_common.cuh"),
                        os.path.join(INCLUDE, "masklab_b200.h"),
                        os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build_library(force=False, verbose=False, out=None, extra=()):
    """Compile every CUDA source into one shared library.  Returns the path.  `out`/`extra`
    build a tuning variant (other -D flags) next to the in-tree library."""
    if out is None and not force and not is_stale():
        return LIB_PATH
    out = out or LIB_PATH
    cmd = [find_nvcc()] + NVCC_FLAGS + list(extra) + ["-I", INCLUDE, "-I", CSRC, "-o", out] + sources()
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd))
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return out


DEBUG_LIB_PATH = os.path.join(HERE, "libmasklab_b200_dbg.so")


def build_debug_bounds(verbose=False):
    """The same sources with -DMLP_DEBUG_BOUNDS (common.cuh): every checked scratch / shared-memory index traps
    with its source line when it leaves its extent.  Select it with MASKLAB_B200_LIB=<this path>."""
    return build_library(force=True, verbose=verbose, out=DEBUG_LIB_PATH, extra=["-DMLP_DEBUG_BOUNDS"])


if __name__ == "__main__":
    if "--debug-bounds" in sys.argv:
        print(build_debug_bounds(verbose="-v" in sys.argv))
    else:
        print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
