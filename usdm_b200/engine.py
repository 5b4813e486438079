"""ctypes binding of libtvb.so (include/tvb.h, include/tvb_vocoder.h).  Raw pointers and sizes only; no torch types cross the ABI.

The library is the product: if it is missing or cannot be loaded this module raises, it never falls back to another
implementation.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

from . import build as _build

PRECISION_BF16 = 0
PRECISION_FP32_VALIDATE = 1
SOLVER_EULER = 0
SOLVER_HEUN = 1

ERR_INVALID, ERR_CUDA, ERR_STATE, ERR_NOMEM = -1, -2, -3, -4


class TvbConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "n_feats", "n_tokens", "embedding_dim", "hidden_size", "intermediate_size", "num_attention_heads",
        "num_hidden_layers", "convpos_width", "convpos_groups", "convpos_depth")] + [("sigma_min", C.c_float)]


class TvbVocConfig(C.Structure):          # include/tvb_vocoder.h: tvb_voc_config
    _fields_ = [("num_mels", C.c_int32), ("upsample_initial_channel", C.c_int32), ("num_upsamples", C.c_int32),
                ("upsample_rates", C.c_int32 * 8), ("upsample_kernel_sizes", C.c_int32 * 8), ("num_kernels", C.c_int32),
                ("resblock_kernel_sizes", C.c_int32 * 4), ("resblock_dilation_sizes", (C.c_int32 * 3) * 4),
                ("snake_logscale", C.c_int32), ("snakebeta", C.c_int32)]


_P = C.c_void_p
_SIGNATURES = {
    # name: (restype, argtypes)  -- must list every symbol include/tvb.h declares (tests check this)
    "tvb_create": (C.c_int, [C.POINTER(TvbConfig), C.c_int, C.POINTER(_P)]),
    "tvb_destroy": (C.c_int, [_P]),
    "tvb_load_weight": (C.c_int, [_P, C.c_char_p, _P, C.POINTER(C.c_int64), C.c_int32]),
    "tvb_finalize_weights": (C.c_int, [_P]),
    "tvb_set_precision": (C.c_int, [_P, C.c_int32]),
    "tvb_get_precision": (C.c_int, [_P]),
    "tvb_set_option": (C.c_int, [_P, C.c_char_p, C.c_int32]),
    "tvb_estimator_forward": (C.c_int, [_P, _P, _P, _P, C.POINTER(C.c_float), C.POINTER(C.c_int64), C.c_int32, C.c_int32, _P, _P]),
    "tvb_sample_cfg": (C.c_int, [_P, _P, _P, _P, C.POINTER(C.c_int64), C.c_int32, C.c_int32, C.c_float, C.c_float, C.c_int32, _P, _P]),
    "tvb_solve": (C.c_int, [_P, C.c_int32, _P, _P, _P, C.POINTER(C.c_int64), C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                            C.POINTER(C.c_float), C.c_float, C.c_int32, _P, _P, _P, _P]),
    "tvb_process_unit": (C.c_int, [C.c_int, _P, C.c_int64, C.c_int32, C.c_int32, _P, C.c_int64, _P]),
    "tvb_mel_spectrogram": (C.c_int, [C.c_int, _P, C.c_int64, C.c_int32, C.c_int32, C.c_int32, _P, C.c_float, C.c_float, _P, C.c_int64, _P]),
    "tvb_profile_read": (C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "tvb_async_status": (C.c_int, [_P]),
    "tvb_kernel_launches": (C.c_int64, [_P]),
    "tvb_device_bytes": (C.c_int64, [_P]),
    "tvb_debug_gemm_bf16": (C.c_int, [C.c_int, _P, _P, _P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P]),
    "tvb_debug_attention": (C.c_int, [C.c_int, _P, _P, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int32, C.c_int32, C.c_int32,
                                      C.c_int32, _P, C.c_int32, C.c_int32, _P]),
    "tvb_debug_set_trace": (C.c_int, [_P]),
    "tvb_debug_pair_gemm_bn": (C.c_int, [C.c_int32, C.c_int32, C.c_int32]),
    # include/tvb_vocoder.h
    "tvb_voc_create": (C.c_int, [C.POINTER(TvbVocConfig), C.c_int, C.POINTER(_P)]),
    "tvb_voc_destroy": (None, [_P]),
    "tvb_voc_load_weight": (C.c_int, [_P, C.c_char_p, _P, C.POINTER(C.c_int64), C.c_int32]),
    "tvb_voc_finalize_weights": (C.c_int, [_P]),
    "tvb_voc_set_precision": (C.c_int, [_P, C.c_int32]),
    "tvb_voc_set_option": (C.c_int, [_P, C.c_char_p, C.c_int32]),
    "tvb_voc_forward": (C.c_int, [_P, _P, C.c_int32, C.c_int32, _P, _P]),
    "tvb_voc_debug_activation": (C.c_int, [_P, C.c_char_p, _P, C.c_int32, C.c_int32, _P, _P]),
    "tvb_voc_kernel_launches": (C.c_int64, [_P]),
    "tvb_voc_set_profile": (C.c_int, [_P, C.c_int32]),
    "tvb_voc_profile_read": (C.c_int, [_P, C.POINTER(C.c_float), C.c_int32]),
    "tvb_voc_device_bytes": (C.c_int64, [_P]),
    "tvb_last_error": (C.c_char_p, []),
    "tvb_version": (C.c_char_p, []),
}

_lib: Optional[C.CDLL] = None


def library_path() -> str:
    """The in-tree build; TVB_LIBRARY overrides it (A/B runs of two builds on one box, tools/runs/)."""
    return os.environ.get("TVB_LIBRARY") or _build.LIB


def load_library() -> C.CDLL:
    """dlopen libtvb.so and bind every exported symbol.  Raises RuntimeError if the library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: build it with `python -m usdm_b200.build` "
                           "(usdm_b200 has no CPU or PyTorch fallback)")
    lib = C.CDLL(path)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int) -> None:
    """Map a tvb_status to the Python exception the reference's callers would see."""
    if rc == 0:
        return
    msg = load_library().tvb_last_error().decode("utf-8", "replace")
    if rc == ERR_INVALID:
        raise ValueError(msg)
    if rc == ERR_NOMEM:
        raise MemoryError(msg)
    raise RuntimeError(msg)
