"""ctypes binding of libmasklab_b200.so (include/masklab_b200.h) and the DLPack handoff.

PyTorch is used for device memory, streams and (elsewhere) torch.distributed only;
every computation of the path is a kernel of the shared library.  There is NO CPU or
eager fallback: loading fails loudly when the library is missing, and every tensor must
live on the ctx's CUDA device.
"""
import ctypes
import os
import threading

import torch
import torch.utils.dlpack

from . import build as _build

MLP_OK, MLP_EINVAL, MLP_ECUDA, MLP_ENOMEM, MLP_EDLPACK, MLP_EBATCH, MLP_EFROZEN = 0, -1, -2, -3, -4, -5, -6
MLP_F32, MLP_I32, MLP_U8, MLP_I64 = 0, 1, 2, 3
MLP_MAX_LEVELS, MLP_MAX_ANCHORS, MLP_MAX_BATCH, MLP_MAX_KEEP = 8, 32, 32, 2048
MLP_PASTE_F32, MLP_PASTE_U8, MLP_PASTE_BITS, MLP_PASTE_NONE = 0, 1, 2, 3
MLP_PASTE_PREFILLED, MLP_MASKS_PLANAR = 0x100, 0x200

_ERRNAMES = {MLP_EINVAL: "MLP_EINVAL", MLP_ECUDA: "MLP_ECUDA", MLP_ENOMEM: "MLP_ENOMEM",
             MLP_EDLPACK: "MLP_EDLPACK", MLP_EBATCH: "MLP_EBATCH", MLP_EFROZEN: "MLP_EFROZEN"}


class MaskLabError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"{_ERRNAMES.get(code, code)}: {message}")
        self.code = code


class InvalidArgumentError(MaskLabError, ValueError):
    """Raised where the reference would raise tf.errors.InvalidArgumentError."""


class PriorConfigC(ctypes.Structure):
    _fields_ = [("num_levels", ctypes.c_int32), ("padding_same", ctypes.c_int32),
                ("stride", ctypes.c_int32 * MLP_MAX_LEVELS),
                ("num_anchors", ctypes.c_int32 * MLP_MAX_LEVELS),
                ("anchor_w", (ctypes.c_int32 * MLP_MAX_ANCHORS) * MLP_MAX_LEVELS),
                ("anchor_h", (ctypes.c_int32 * MLP_MAX_ANCHORS) * MLP_MAX_LEVELS)]


class DetectionParamsC(ctypes.Structure):
    _fields_ = [("min_confidence", ctypes.c_float), ("nms_iou_threshold", ctypes.c_float),
                ("post_iou_threshold", ctypes.c_float), ("nms_max_output_size", ctypes.c_int32),
                ("strict_batch", ctypes.c_int32)]


MLP_MAX_DRAW_CLASSES = 16
MLP_NUM_STAGES = 24


class DrawColorsC(ctypes.Structure):
    _fields_ = [("num_classes", ctypes.c_int32), ("alpha", ctypes.c_float),
                ("rgb", (ctypes.c_float * 3) * MLP_MAX_DRAW_CLASSES)]

    @classmethod
    def make(cls, colors, alpha):
        colors = [list(c) for c in colors]
        if not 1 <= len(colors) <= MLP_MAX_DRAW_CLASSES or any(len(c) != 3 for c in colors):
            raise InvalidArgumentError(MLP_EINVAL, f"colors must be 1..{MLP_MAX_DRAW_CLASSES} RGB triples")
        out = cls()
        out.num_classes = len(colors)
        out.alpha = float(alpha)
        for i, c in enumerate(colors):
            for k in range(3):
                out.rgb[i][k] = float(c[k])
        return out


class TensorViewC(ctypes.Structure):
    _fields_ = [("data", ctypes.c_void_p), ("device_id", ctypes.c_int32), ("ndim", ctypes.c_int32),
                ("dtype_code", ctypes.c_int32), ("dtype_bits", ctypes.c_int32),
                ("shape", ctypes.c_int64 * 8), ("numel", ctypes.c_int64)]


_P, _I, _L, _F = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float

# name -> (restype, argtypes); mirrors include/masklab_b200.h one to one
SIGNATURES = {
    "mlp_version": (_I, []),
    "mlp_last_error": (ctypes.c_char_p, []),
    "mlp_ctx_create": (_I, [_I, ctypes.POINTER(_P)]),
    "mlp_ctx_destroy": (None, [_P]),
    "mlp_ctx_device": (_I, [_P]),
    "mlp_ctx_sm_count": (_I, [_P]),
    "mlp_ctx_scratch_bytes": (_L, [_P]),
    "mlp_ctx_launch_count": (_L, [_P]),
    "mlp_ctx_freeze_scratch": (_I, [_P, _I]),
    "mlp_stage_name": (ctypes.c_char_p, [_I]),
    "mlp_ctx_profile_enable": (_I, [_P, _I]),
    "mlp_ctx_profile_read": (_I, [_P, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int64)]),
    "mlp_dlpack_view": (_I, [_P, _P, _I, ctypes.POINTER(TensorViewC)]),
    "mlp_prior_count": (_L, [ctypes.POINTER(PriorConfigC), _I, _I]),
    "mlp_prior_layer": (_I, [_P, ctypes.POINTER(PriorConfigC), _I, _I, _I, _P, _P]),
    "mlp_restore_boxes": (_I, [_P, _P, _P, _I, _L, _P, _P]),
    "mlp_restore_boxes_from_prior": (_I, [_P, ctypes.POINTER(PriorConfigC), _P, _I, _I, _I, _P, _P]),
    "mlp_normalize_boxes": (_I, [_P, _P, _L, _I, _F, _F, _P, _P]),
    "mlp_detection_proposal": (_I, [_P, _P, _P, _I, _L, _I, ctypes.POINTER(DetectionParamsC),
                                    _P, _P, _P, _P, _P]),
    "mlp_detect_from_heads": (_I, [_P, ctypes.POINTER(PriorConfigC), _P, _P, _I, _I, _I, _I,
                                   ctypes.POINTER(DetectionParamsC), _P, _P, _P, _P, _P]),
    "mlp_mask_distribute": (_I, [_P, _P, _L, _I, _F, _P, _P]),
    "mlp_roi_align_plan": (_I, [_P, _P, _I, _I, _I, _P, _I, _P, _P, _P]),
    "mlp_roi_align_run": (_I, [_P, ctypes.POINTER(_P), ctypes.POINTER(ctypes.c_int32),
                               ctypes.POINTER(ctypes.c_int32), _I, _I, _P, _I, _I, _I, _P, _F, _F,
                               _I, _I, _P, _P, ctypes.POINTER(_P), _P, _P]),
    "mlp_trim_plan": (_I, [_P, _P, _I, _I, _P, _P, _P, _P]),
    "mlp_trim_run": (_I, [_P, _P, _P, _I, _I, _P, _I, _I, _I, _P, _P, _P, _P]),
    "mlp_upsample_output": (_I, [_P, _P, _L, _F, _F, _P, _P, _L, _P, _P]),
    "mlp_crop_and_pad_mask": (_I, [_P, _P, _P, _I, _I, _I, _P, _I, _I, _I, _I, _I, _P, _P]),
    "mlp_detect_align": (_I, [_P, ctypes.POINTER(PriorConfigC), _P, _P, _I, _I, _I, _I,
                              ctypes.POINTER(DetectionParamsC), _I, _F, ctypes.POINTER(_P),
                              ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32), _I, _I, _I,
                              _P, _P, _P, _P, _P, _P, _P, ctypes.POINTER(_P), _P, _P]),
    "mlp_detect_plan": (_I, [_P, ctypes.POINTER(PriorConfigC), _P, _P, _I, _I, _I, _I,
                             ctypes.POINTER(DetectionParamsC), _I, _F, _P, _P, _P, _P, _P, _P, _P, _P]),
    "mlp_paste_prefill": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P]),
    "mlp_trim_paste": (_I, [_P, _P, _P, _I, _I, _P, _I, _I, _I, _F, _F, _I, _I, _I, _I, _P, _P, _P, _P, _P]),
    "mlp_clip_pool_bound": (_L, [_I, _I, _I, _I]),
    "mlp_clip_masks": (_I, [_P, _P, _P, _I, _P, _I, _P, _I, _I, _I, _I, _I, _I, _P, _P, _L, _P, _P]),
    "mlp_road_scan": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _F, _P, _P, _P, _P, _P]),
    "mlp_summary_output": (_I, [_P, _P, _P, _I, _P, _P, _P, _I, _I, _I, _P, _I, _I, _F, _P, _P, _P]),
    "mlp_tile_summary": (_I, [_P, _P, _P, _P, _I, _P, _I, _P, _I, _I, _I, _P, _I, _I, _P, _P, _P, _I, _I,
                              _F, _P, _P, _P, _P]),
    "mlp_resize_bilinear": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P]),
    "mlp_calculate_iou": (_I, [_P, _P, _I, _I, _P, _I, _I, _P, _P]),
    "mlp_assign_boxes": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _P]),
    "mlp_assign_masks": (_I, [_P, _P, _I, _P, _I, _P, _I, _I, _I, _I, _I, _I, _F, _P, _P]),
    "mlp_detection_iou_metric": (_I, [_P, _P, _I, _P, _I, _I, _P, _P]),
    "mlp_semantic_smoothing": (_I, [_P, _P, _I, _I, _I, _I, _I, _F, _P, _P]),
    "mlp_draw_boxes": (_I, [_P, _P, _I, _P, _I, _I, _I, _P, _I, _I, _P, _P]),
    "mlp_draw_segmentation": (_I, [_P, _P, _I, _P, _I, _I, _I, _I, ctypes.POINTER(DrawColorsC), _P, _P]),
    "mlp_draw_instance": (_I, [_P, _P, _I, _P, _P, _I, _I, _I, _I, _P, _I, _I, ctypes.POINTER(DrawColorsC), _P, _P]),
    "mlp_draw_tiles": (_I, [_P, _P, _I, _P, _P, _P, _I, _P, _I, _P, _I, _I, _I, _P, _I, _I, _I, _I,
                            ctypes.POINTER(DrawColorsC), _P, _I, ctypes.POINTER(DrawColorsC), _P, _P]),
    "mlp_draw_tiles_boxes": (_I, [_P, _P, _I, _P, _P, _P, _I, _P, _I, _P, _I, _I, _I, _P, _I, _I, _I, _I,
                                  ctypes.POINTER(DrawColorsC), _P, _I, ctypes.POINTER(DrawColorsC), _P, _P]),
    "mlp_jpeg_max_bytes": (_L, [_I, _I]),
    "mlp_jpeg_header": (_I, [_I, _I, _I, _P, _I]),
    "mlp_jpeg_encode": (_I, [_P, _P, _I, _I, _I, _I, _P, _L, _P, _P]),
    "mlp_mold_batch_plan": (_I, [_P, _P, _L, _I, _P, _P, _P]),
    "mlp_mold_batch_run": (_I, [_P, _P, _P, _L, _L, _I, _I, _P, _P, _P]),
}

_lib = None
_lib_lock = threading.Lock()


def library_path():
    """The in-tree library; MASKLAB_B200_LIB points at another BUILD of the same sources (A/B tuning)."""
    return os.environ.get("MASKLAB_B200_LIB") or _build.LIB_PATH


def load_library():
    """dlopen the in-tree shared library; raise (never fall back) if it is not there."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        path = library_path()
        if not os.path.exists(path):
            raise RuntimeError(
                f"{path} is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). masklab_b200 has no CPU or PyTorch fallback.")
        lib = ctypes.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name, None)
            if fn is None:
                continue            # optional symbols are checked by tests/test_abi.py
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


def last_error():
    return load_library().mlp_last_error().decode("utf-8", "replace")


def check(rc):
    if rc == MLP_OK:
        return
    msg = last_error()
    if rc in (MLP_EINVAL, MLP_EBATCH, MLP_EDLPACK):
        raise InvalidArgumentError(rc, msg)
    raise MaskLabError(rc, msg)


_PyCapsule_GetPointer = ctypes.pythonapi.PyCapsule_GetPointer
_PyCapsule_GetPointer.restype = ctypes.c_void_p
_PyCapsule_GetPointer.argtypes = [ctypes.py_object, ctypes.c_char_p]

_TORCH_TO_MLP = {torch.float32: MLP_F32, torch.int32: MLP_I32, torch.uint8: MLP_U8, torch.int64: MLP_I64}


class Context:
    """One mlp_ctx per (process, CUDA device).  Owns scratch; not thread-safe."""

    _instances = {}

    def __init__(self, device):
        self.lib = load_library()
        self.device = int(device)
        h = ctypes.c_void_p()
        check(self.lib.mlp_ctx_create(self.device, ctypes.byref(h)))
        self.handle = h
        self.shared = False          # True for the per-device singleton of Context.get()
        self.frozen = False

    def __del__(self):
        # private contexts (pipelines that own their scratch) release it with the last reference
        try:
            if getattr(self, "handle", None) and not getattr(self, "shared", True):
                self.lib.mlp_ctx_destroy(self.handle)
                self.handle = None
        except Exception:            # interpreter shutdown
            pass

    def freeze(self, on=True):
        """A captured CUDA graph bakes this ctx's scratch pointers in: once frozen, a call that would
        have to grow (= free and re-allocate) an arena raises MLP_EFROZEN instead."""
        check(self.lib.mlp_ctx_freeze_scratch(self.handle, 1 if on else 0))
        self.frozen = bool(on)

    @classmethod
    def get(cls, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("masklab_b200 needs a CUDA device (B200, sm_100a); there is no CPU path")
        if device is None:
            device = torch.cuda.current_device()
        if isinstance(device, torch.device):
            device = device.index if device.index is not None else torch.cuda.current_device()
        device = int(device)
        inst = cls._instances.get(device)
        if inst is None:
            inst = cls._instances[device] = cls(device)
            inst.shared = True
        return inst

    def close(self):
        if self.handle:
            self.lib.mlp_ctx_destroy(self.handle)
            self.handle = None
            Context._instances.pop(self.device, None)

    # ---- helpers ----
    @property
    def torch_device(self):
        return torch.device("cuda", self.device)

    def stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def launch_count(self):
        return int(self.lib.mlp_ctx_launch_count(self.handle))

    def scratch_bytes(self):
        return int(self.lib.mlp_ctx_scratch_bytes(self.handle))

    def sm_count(self):
        return int(self.lib.mlp_ctx_sm_count(self.handle))

    def profile(self, enable=True):
        check(self.lib.mlp_ctx_profile_enable(self.handle, 1 if enable else 0))

    def profile_read(self):
        """{stage name: (total ms, bracketed calls)} since profile(True); synchronises."""
        ms = (ctypes.c_double * MLP_NUM_STAGES)()
        cnt = (ctypes.c_int64 * MLP_NUM_STAGES)()
        check(self.lib.mlp_ctx_profile_read(self.handle, ms, cnt))
        out = {}
        for i in range(MLP_NUM_STAGES):
            name = self.lib.mlp_stage_name(i).decode()
            if name and cnt[i]:
                out[name] = (float(ms[i]), int(cnt[i]))
        return out

    def view(self, tensor, dtype=None):
        """DLPack handoff: validate `tensor` in C and return its device pointer."""
        if not isinstance(tensor, torch.Tensor):
            raise InvalidArgumentError(MLP_EINVAL, f"expected a torch.Tensor, got {type(tensor)}")
        if dtype is not None and tensor.dtype != dtype:
            raise InvalidArgumentError(MLP_EDLPACK, f"expected dtype {dtype}, got {tensor.dtype}")
        code = _TORCH_TO_MLP.get(tensor.dtype, -1)
        if tensor.numel() == 0:
            return ctypes.c_void_p(tensor.data_ptr())
        cap = torch.utils.dlpack.to_dlpack(tensor)
        ptr = _PyCapsule_GetPointer(cap, b"dltensor")
        v = TensorViewC()
        check(self.lib.mlp_dlpack_view(self.handle, ptr, code, ctypes.byref(v)))
        return ctypes.c_void_p(v.data)

    def empty(self, shape, dtype):
        return torch.empty(shape, dtype=dtype, device=self.torch_device)


def as_device_f32(ctx, x, what):
    """The layers accept whatever tensor the upstream graph produced; like tf.cast(.., f32)
    in the reference, non-f32 inputs are converted on the device (plumbing only)."""
    if not isinstance(x, torch.Tensor):
        raise InvalidArgumentError(MLP_EINVAL, f"{what}: expected a CUDA torch.Tensor, got {type(x)}")
    if not x.is_cuda:
        raise InvalidArgumentError(
            MLP_EDLPACK, f"{what}: tensor is on {x.device}; masklab_b200 has no CPU path")
    if x.dtype != torch.float32:
        x = x.to(torch.float32)
    return x.contiguous()
