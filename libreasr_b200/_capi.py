"""ctypes binding of the C ABI declared in include/rnnt_b200.h.

The CUDA library is the product: importing this module never falls back to a CPU
implementation.  If the shared object is missing it is built with nvcc
(``libreasr_b200.build``); if that fails, or no sm_100 GPU is present when a handle is
created, a ``RuntimeError`` is raised.
"""
import ctypes as C
import os

from . import build as _build

LIB_PATH = _build.LIB_PATH


class Config(C.Structure):
    """Mirror of ``rnnt_b200_config`` (include/rnnt_b200.h)."""

    _fields_ = [
        ("sample_rate", C.c_int32), ("n_fft", C.c_int32), ("win_length", C.c_int32), ("hop_length", C.c_int32),
        ("n_mels", C.c_int32), ("n_stack", C.c_int32), ("downsample", C.c_int32),
        ("enc_layers", C.c_int32), ("pred_layers", C.c_int32),
        ("hidden_sz", C.c_int32), ("embed_sz", C.c_int32), ("joint_sz", C.c_int32), ("vocab_sz", C.c_int32),
        ("blank", C.c_int32), ("bos", C.c_int32), ("device", C.c_int32), ("gemm_mode", C.c_int32),
        ("lm_layers", C.c_int32), ("lm_hidden_sz", C.c_int32), ("lm_embed_sz", C.c_int32),
        ("log_offset", C.c_float), ("ln_eps", C.c_float), ("bn_eps", C.c_float),
        ("lm_alpha", C.c_float), ("lm_theta", C.c_float),
    ]


GEMM_FP32_SIMT, GEMM_TC_FP16X3, GEMM_TC_FP16 = 0, 1, 2

_vp, _i32, _i64, _cp = C.c_void_p, C.c_int32, C.c_int64, C.c_char_p

# symbol -> (restype, argtypes); every function include/rnnt_b200.h declares
SIGNATURES = {
    "rnnt_b200_abi_version": (_i32, []),
    "rnnt_b200_default_config": (_i32, [C.POINTER(Config)]),
    "rnnt_b200_create": (_i32, [C.POINTER(Config), C.POINTER(_vp)]),
    "rnnt_b200_destroy": (_i32, [_vp]),
    "rnnt_b200_last_error": (_cp, [_vp]),
    "rnnt_b200_set_weight": (_i32, [_vp, _cp, _vp, _i64]),
    "rnnt_b200_finalize": (_i32, [_vp, _vp]),
    "rnnt_b200_reserve": (_i32, [_vp, _i32, _i64]),
    "rnnt_b200_num_frames": (_i64, [_vp, _i64]),
    "rnnt_b200_num_steps": (_i64, [_vp, _i64]),
    "rnnt_b200_features": (_i32, [_vp, _vp, _vp, _i32, _i64, _vp, _vp]),
    "rnnt_b200_logmel": (_i32, [_vp, _vp, _vp, _i32, _i64, _vp, _vp]),
    "rnnt_b200_features_stream": (_i32, [_vp, _vp, _i32, _i64, _vp, _vp]),
    "rnnt_b200_resample_len": (_i64, [_vp, _i64, _i32]),
    "rnnt_b200_resample": (_i32, [_vp, _vp, _i32, _i64, _i32, _vp, _vp]),
    "rnnt_b200_encode": (_i32, [_vp, _vp, _vp, _i32, _i32, _vp, _vp, _i32, _vp, _vp]),
    "rnnt_b200_predict": (_i32, [_vp, _vp, _i32, _vp, _i32, _vp, _vp]),
    "rnnt_b200_joint": (_i32, [_vp, _vp, _vp, _i32, _vp, _vp]),
    "rnnt_b200_lm_state_bytes": (_i32, [_vp, _i32, C.POINTER(_i64)]),
    "rnnt_b200_set_lm_state": (_i32, [_vp, _vp, _i32]),
    "rnnt_b200_decode_greedy": (_i32, [_vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp, _i32, _vp, _i32, _vp, _vp, _vp, _vp, _i32, _vp]),
    "rnnt_b200_transcribe": (_i32, [_vp, _vp, _vp, _i32, _i64, _i32, _vp, _i32, _vp, _vp, _vp, _vp]),
    "rnnt_b200_transcribe_host": (_i32, [_vp, _vp, _vp, _i32, _i64, _i32, _vp, _i32, _vp, _vp, _vp]),
    "rnnt_b200_pipeline_submit": (_i32, [_vp, _vp, _i32, _vp, _i32, _i64, _i32, _i32, _vp, _i32, _vp, _vp, _vp]),
    "rnnt_b200_pipeline_collect": (_i32, [_vp, _i32]),
    "rnnt_b200_pipeline_query": (_i32, [_vp, _i32, C.POINTER(_i32), C.POINTER(_i32)]),
    "rnnt_b200_stream_open": (_i32, [_vp, _i32, _i32, _i32, _i32, _i32, C.POINTER(_vp)]),
    "rnnt_b200_stream_push": (_i32, [_vp, _vp, _i32, _vp, _vp, _i32, _vp, C.POINTER(_i32), _vp]),
    "rnnt_b200_decode_beam": (_i32, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _i32, _vp, _vp, _vp]),
    "rnnt_b200_forward_loss": (_i32, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp]),
    "rnnt_b200_rnnt_loss": (_i32, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp]),
    "rnnt_b200_stream_reset": (_i32, [_vp, _i32]),
    "rnnt_b200_stream_reset_state": (_i32, [_vp, _i32]),
    "rnnt_b200_stream_close": (_i32, [_vp]),
    "rnnt_b200_selftest_gemm": (_i32, [_vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _vp]),
    "rnnt_b200_kernel_launches": (_i64, [_vp]),
    "rnnt_b200_fp32_decode_launches": (_i64, [_vp]),
    "rnnt_b200_set_profiling": (_i32, [_vp, _i32]),
    "rnnt_b200_stage_times_ms": (_i32, [_vp, _vp]),
}

_lib = None


def load_library(build_if_missing=True):
    """Loads (building first if stale/missing) librnnt_b200.so and types its symbols."""
    global _lib
    if _lib is not None:
        return _lib
    if build_if_missing and _build.needs_build():
        _build.build()
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: the CUDA library is required (there is no CPU fallback); "
                           "run `python -m libreasr_b200.build`")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


class RnntError(RuntimeError):
    pass


def check(lib, handle, status):
    """Raises with ``rnnt_b200_last_error`` on a non-zero status; ERR_INVALID maps to
    ``ValueError`` like the reference's shape checks (haste/base_rnn.py:81-117)."""
    if status == 0:
        return
    msg = lib.rnnt_b200_last_error(handle)
    msg = msg.decode() if msg else "unknown error"
    if status == -1:
        raise ValueError(msg)
    raise RnntError(f"rnnt_b200 status {status}: {msg}")
