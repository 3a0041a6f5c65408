"""ctypes binding of libserenc.so (the C ABI declared in include/serenc.h).

There is deliberately no fallback: if the library is missing or no sm_100 GPU is present, the product path
raises. PyTorch is used only for device memory (tensor handles) and streams.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG_DIR, "libserenc.so")

# every symbol include/serenc.h declares (tests/test_abi.py checks the header against this list)
EXPORTED_SYMBOLS = [
    "serenc_create", "serenc_destroy", "serenc_load_tensor", "serenc_finalize", "serenc_last_error",
    "serenc_version", "serenc_w2v_num_frames", "serenc_w2v_workspace_bytes", "serenc_wav_normalize",
    "serenc_encode_w2v", "serenc_unpack_frames", "serenc_logmel", "serenc_whisper_workspace_bytes",
    "serenc_encode_whisper", "serenc_op_gemm", "serenc_op_gemm_grouped", "serenc_op_layernorm",
    "serenc_op_attention", "serenc_wavlm_bucket", "serenc_launch_count", "serenc_set_profiling", "serenc_get_profile",
    "serenc_debug_gemm_trace", "serenc_sync", "serenc_is_poisoned", "serenc_encode_w2v_ex", "serenc_unpack_rows",
    "serenc_logmel_ex", "serenc_encode_whisper_ex", "serenc_text_workspace_bytes", "serenc_encode_text",
]

WAV_F32, WAV_I16 = 0, 1


class SerencConfig(C.Structure):
    _fields_ = [
        ("arch", C.c_int32), ("hidden", C.c_int32), ("layers", C.c_int32), ("heads", C.c_int32),
        ("ffn", C.c_int32), ("conv_dim", C.c_int32), ("conv_bias", C.c_int32), ("wavlm_rel_bias", C.c_int32),
        ("num_buckets", C.c_int32), ("max_distance", C.c_int32), ("pos_conv_kernel", C.c_int32),
        ("pos_conv_groups", C.c_int32), ("n_mels", C.c_int32), ("max_source_positions", C.c_int32),
        ("layer_norm_eps", C.c_float), ("conv_group_norm", C.c_int32), ("post_layer_norm", C.c_int32),
        ("no_feat_proj_ln", C.c_int32), ("vocab_size", C.c_int32), ("max_positions", C.c_int32),
        ("type_vocab_size", C.c_int32), ("pad_token_id", C.c_int32), ("reserved", C.c_int32 * 1),
    ]


class SerencW2VCall(C.Structure):
    """serenc_w2v_call (include/serenc.h)."""
    _fields_ = [
        ("wav_dev", C.c_void_p), ("wav_dtype", C.c_int32), ("batch", C.c_int32),
        ("sample_start", C.POINTER(C.c_int64)), ("sample_len", C.POINTER(C.c_int32)),
        ("normalize", C.c_int32), ("reduce", C.c_int32), ("layer_mask", C.c_uint64),
        ("layer_weights", C.POINTER(C.c_float)), ("frames_out_dev", C.c_void_p), ("pooled_out_dev", C.c_void_p),
        ("extract_features_out_dev", C.c_void_p), ("frame_offsets_out", C.POINTER(C.c_int64)),
        ("workspace_dev", C.c_void_p), ("workspace_bytes", C.c_size_t), ("stream", C.c_void_p),
    ]


class SerencError(RuntimeError):
    """Raised when a libserenc entry point returns a negative status."""

    def __init__(self, status: int, message: str):
        super().__init__(f"libserenc status {status}: {message}")
        self.status = status


_lib = None
_lock = threading.Lock()


def load_library(build_if_missing: bool = True) -> C.CDLL:
    """Load (building in-tree if needed) libserenc.so and declare the argument types."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if build_if_missing:
            # always go through build_library(): it compares the digest of csrc/ with the stamp of the built library
            # (cheap) and rebuilds when they differ, so an edited kernel is never measured through a stale .so; it
            # builds into a temporary file under a lock, so concurrent torchrun ranks never load a half-written file
            from .build import build_library
            build_library()
        elif not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -m interspeech_ser_b200.build` "
                               "(there is no CPU or PyTorch fallback for this path)")
        lib = C.CDLL(LIB_PATH)
        vp, i32, i64, u64, sz = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_size_t
        pi32, pi64, pf = C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_float)
        sig = {
            "serenc_create": (C.c_int, [C.POINTER(SerencConfig), C.c_int, C.POINTER(vp)]),
            "serenc_destroy": (C.c_int, [vp]),
            "serenc_load_tensor": (C.c_int, [vp, C.c_char_p, vp, pi64, C.c_int]),
            "serenc_finalize": (C.c_int, [vp]),
            "serenc_last_error": (C.c_char_p, []),
            "serenc_version": (C.c_char_p, []),
            "serenc_w2v_num_frames": (i64, [i64]),
            "serenc_wavlm_bucket": (C.c_int, [C.c_int, C.c_int, C.c_int]),
            "serenc_w2v_workspace_bytes": (C.c_int, [vp, pi32, C.c_int, C.POINTER(sz)]),
            "serenc_wav_normalize": (C.c_int, [vp, vp, pi64, pi32, C.c_int, vp, i64, i32, vp]),
            "serenc_encode_w2v": (C.c_int, [vp, vp, pi64, pi32, C.c_int, C.c_int, u64, C.c_int, vp, vp, pi64, vp, sz, vp]),
            "serenc_unpack_frames": (C.c_int, [vp, vp, pi64, C.c_int, i32, vp, vp]),
            "serenc_logmel": (C.c_int, [vp, vp, pi64, pi32, C.c_int, vp, vp, vp]),
            "serenc_whisper_workspace_bytes": (C.c_int, [vp, C.c_int, C.POINTER(sz)]),
            "serenc_encode_whisper": (C.c_int, [vp, vp, C.c_int, u64, C.c_int, pi32, vp, vp, vp, sz, vp]),
            "serenc_op_gemm": (C.c_int, [vp, vp, i64, i64, i64, vp, i64, vp, vp, C.c_int, vp, vp, vp]),
            "serenc_op_gemm_grouped": (C.c_int, [vp, vp, i64, C.c_int, C.c_int, C.c_int, vp, C.c_int, vp, C.c_int, vp, vp]),
            "serenc_op_layernorm": (C.c_int, [vp, vp, i64, C.c_int, vp, vp, C.c_float, C.c_int, vp, vp, vp]),
            "serenc_op_attention": (C.c_int, [vp, vp, pi64, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp]),
            "serenc_launch_count": (i64, [vp]),
            "serenc_debug_gemm_trace": (C.c_int, [vp, vp]),
            "serenc_set_profiling": (C.c_int, [vp, C.c_int]),
            "serenc_get_profile": (C.c_int, [vp, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), pi64]),
            "serenc_sync": (C.c_int, [vp, vp]),
            "serenc_is_poisoned": (C.c_int, [vp]),
            "serenc_encode_w2v_ex": (C.c_int, [vp, C.POINTER(SerencW2VCall)]),
            "serenc_unpack_rows": (C.c_int, [vp, vp, pi64, C.c_int, i32, i32, vp, vp]),
            "serenc_logmel_ex": (C.c_int, [vp, vp, C.c_int, pi64, pi32, C.c_int, vp, vp, vp]),
            "serenc_encode_whisper_ex": (C.c_int, [vp, vp, C.c_int, u64, C.c_int, pf, pi32, vp, vp, vp, sz, vp]),
            "serenc_text_workspace_bytes": (C.c_int, [vp, C.c_int, C.c_int, C.POINTER(sz)]),
            "serenc_encode_text": (C.c_int, [vp, vp, pi32, C.c_int, C.c_int, u64, C.c_int, pf, vp, vp, vp, sz, vp]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


def check(status: int) -> None:
    if status != 0:
        msg = load_library().serenc_last_error()
        raise SerencError(status, msg.decode("utf-8", "replace") if msg else "")


def i64_array(values):
    arr = (C.c_int64 * len(values))(*[int(v) for v in values])
    return arr


def i32_array(values):
    arr = (C.c_int32 * len(values))(*[int(v) for v in values])
    return arr


def f32_array(values):
    return (C.c_float * len(values))(*[float(v) for v in values])
