"""ctypes binding of libegx (include/egx.h).  Fails loudly when the library is missing:
there is no PyTorch/CPU fallback for the pose path."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libegx.so")

EGX_PREC_FP32, EGX_PREC_TC = 0, 1
EGX_DTYPE_F32, EGX_DTYPE_I64 = 0, 1


class EgxCfg(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "frames", "prior_frames", "pose_dim", "d_model", "d_inner", "n_layers", "n_head", "d_k",
        "d_v", "n_mels", "spec_w", "n_position", "precision")]


# name -> (restype, argtypes); must list every symbol include/egx.h declares
PROTOTYPES = {
    "egx_version": (C.c_int, []),
    "egx_create": (C.c_int, [C.POINTER(EgxCfg), C.c_int, C.POINTER(C.c_void_p)]),
    "egx_destroy": (None, [C.c_void_p]),
    "egx_last_error": (C.c_char_p, [C.c_void_p]),
    "egx_set_weight": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.POINTER(C.c_int64), C.c_int,
                                 C.c_int]),
    "egx_finalize_weights": (C.c_int, [C.c_void_p]),
    "egx_logmel": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                             C.c_void_p, C.c_void_p]),
    "egx_debug_logmel_global_tile": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                             C.c_void_p, C.c_void_p]),
    "egx_audio_pcm16_to_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "egx_audio_fixed_length": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "egx_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int]),
    "egx_generator_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_size_t, C.c_void_p]),
    "egx_get_tap": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t,
                              C.POINTER(C.c_size_t), C.c_void_p]),
    "egx_debug_trunk": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t,
                                  C.POINTER(C.c_size_t), C.c_void_p, C.c_size_t, C.c_void_p]),
    "egx_fgd_accumulate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p,
                                     C.c_void_p, C.c_void_p]),
    "egx_debug_linear_tc": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                      C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "egx_debug_linear_ln_tc": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "egx_debug_ffn_tc": (C.c_int, [C.c_void_p] * 8 + [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "egx_debug_conv_tc": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                    C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "egx_debug_attention_tc": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                         C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "egx_cvae_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_void_p]),
    "egx_cvae_sample": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "egx_cvae3_sample": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "egx_pose_features": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                   C.c_void_p]),
    "egx_pose_feature_dim": (C.c_int, [C.c_void_p, C.c_int]),
    "egx_row_features_workspace": (C.c_size_t, [C.c_void_p, C.c_int64]),
    "egx_row_features": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p,
                                  C.c_size_t, C.c_void_p]),
    "egx_emotion_net_workspace": (C.c_size_t, [C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "egx_emotion_net_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                         C.c_void_p, C.c_size_t, C.c_void_p]),
    "egx_skeleton_workspace": (C.c_size_t, [C.c_void_p, C.c_int]),
    "egx_skeleton_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_size_t, C.c_void_p]),
    "egx_skeleton_dims": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                    C.POINTER(C.c_int)]),
    "egx_beat_align": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double,
                                 C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "egx_profile_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "egx_profile_read": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.c_int]),
    "egx_launch_count": (C.c_int64, [C.c_void_p]),
}

_lib = None


def load_library(path: str | None = None) -> C.CDLL:
    """dlopen libegx.so (built in-tree by `python -m emotiongestures_b200.build`)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("EGX_LIBRARY", LIB_PATH)
    if not os.path.exists(p):
        raise RuntimeError(
            f"libegx.so not found at {p}: build it with `python -m emotiongestures_b200.build` "
            "(nvcc, sm_100a). There is no fallback path.")
    lib = C.CDLL(p)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)          # AttributeError if the export is missing
        fn.restype = res
        fn.argtypes = args
    if lib.egx_version() != 1:
        raise RuntimeError("libegx.so version mismatch; rebuild")
    if path is None:
        _lib = lib
    return lib
