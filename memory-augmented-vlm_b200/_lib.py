"""ctypes binding of libmavlm.so (the C ABI declared in include/mavlm.h).

The library is built in-tree by ``build.py`` (nvcc, sm_100a).  There is no fallback of any kind:
if the shared object is missing or a symbol is absent, importing / calling raises immediately.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_int64, c_size_t, c_ulonglong, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmavlm.so")

F32, BF16, F16 = 0, 1, 2
ACT_NONE, ACT_GELU_ERF, ACT_RELU = 0, 1, 2
POOL_BILINEAR, POOL_AVERAGE, POOL_MAX = 0, 1, 2
E_INVALID, E_CUDA, E_ARCH, E_INDEX, E_WORKSPACE = -1, -2, -3, -4, -5


class GemmDesc(ctypes.Structure):
    """mavlm_gemm_desc of include/mavlm.h (one nn.Linear call as a tiled problem)."""
    _fields_ = [("A", c_void_p), ("lda", c_int64), ("W", c_void_p), ("ldw", c_int64), ("bias", c_void_p),
                ("resid", c_void_p), ("ldr", c_int64), ("addvec", c_void_p), ("pe_table", c_void_p),
                ("frame_idx", c_void_p), ("tokens_per_frame", ctypes.c_int32), ("C", c_void_p), ("ldc", c_int64),
                ("M", ctypes.c_int32), ("N", ctypes.c_int32), ("K", ctypes.c_int32), ("act", ctypes.c_int32),
                ("out_dtype", ctypes.c_int32)]


class MavlmError(RuntimeError):
    """Raised when a C-ABI call returns a negative status (shape errors surface as RuntimeError,
    like the reference's view/reshape failures)."""


# name -> (restype, argtypes); must list EVERY symbol of include/mavlm.h (tests check this)
PROTOTYPES = {
    "mavlm_version": (c_int, []),
    "mavlm_last_error_string": (c_char_p, []),
    "mavlm_launch_count": (c_ulonglong, []),
    "mavlm_check_device": (c_int, [c_int]),
    "mavlm_pool_pe_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                  c_int, c_void_p]),
    "mavlm_add_pe_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "mavlm_gemm_bias_act_fwd": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_void_p,
                                        c_void_p, c_int64, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "mavlm_gemm_bias_pe_fwd": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int, c_void_p,
                                       c_int64, c_int, c_int, c_int, c_int, c_void_p]),
    "mavlm_gather_rows_fwd": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int,
                                      c_int, c_void_p]),
    "mavlm_gemm_num_tiles": (c_int, [c_void_p]),
    "mavlm_gemm_tiles_fwd": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p]),
    "mavlm_gemm_fill_fwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_void_p]),
    "mavlm_cast_fwd": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p]),
    "mavlm_layernorm_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_int, c_int,
                                    c_void_p]),
    "mavlm_xattn_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int, c_int]),
    "mavlm_xattn_fwd": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int64,
                                c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                c_float, c_int, c_void_p, c_size_t, c_void_p]),
    "mavlm_xattn_colsum": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_int, c_int,
                                   c_int, c_int, c_int, c_float, c_int, c_void_p]),
    "mavlm_assemble_fwd": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "mavlm_gemm_ex": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_int64, c_int, c_void_p, c_int64, c_int, c_int, c_int,
                              c_float, c_int, c_int, c_int, c_void_p, c_int, c_int, c_void_p]),
    "mavlm_colsum": (c_int, [c_void_p, c_int64, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "mavlm_layernorm_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_float,
                                    c_int, c_void_p]),
    "mavlm_act_fwd": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p]),
    "mavlm_act_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p]),
    "mavlm_xattn_bwd_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int, c_int]),
    "mavlm_xattn_bwd": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int64,
                                c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_int64,
                                c_int64, c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int64, c_int, c_int, c_int,
                                c_int, c_int, c_float, c_int, c_void_p, c_size_t, c_void_p]),
    "mavlm_resize_coeffs": (c_int, [c_int, c_int, c_void_p, c_void_p, c_int]),
    "mavlm_frames_preprocess_fwd": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                            c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_double, c_void_p,
                                            c_void_p, c_int, c_void_p]),
    "mavlm_stream_compress_workspace_bytes": (c_size_t, [c_int64, c_int, c_int, c_int]),
    "mavlm_stream_compress_fwd": (c_int, [c_void_p, c_int64, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                          c_void_p, c_size_t, c_int, c_void_p]),
    "mavlm_stream_compress_batched_workspace_bytes": (c_size_t, [c_int, c_int64, c_int, c_int, c_int]),
    "mavlm_stream_compress_batched_fwd": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int, c_int, c_void_p, c_int64, c_void_p,
                                                  c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_void_p]),
    "mavlm_frame_mean_fwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "mavlm_adjacent_cosine_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int]),
    "mavlm_adjacent_cosine_fwd": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_float, c_void_p, c_void_p, c_size_t, c_int,
                                          c_void_p]),
    "mavlm_depth_scores_fwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "mavlm_avg_pool_fwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "mavlm_kmeans_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int, c_int]),
    "mavlm_kmeans_iter_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64,
                                      c_int64, c_int, c_void_p, c_size_t, c_int, c_void_p]),
    "mavlm_row_distance_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int64, c_void_p, c_size_t, c_int, c_void_p]),
    "mavlm_ntm_softmax_fwd": (c_int, [c_void_p, c_int64, c_int64, c_int, c_float, c_float, c_void_p, c_int64, c_int, c_void_p,
                                      c_void_p, c_int, c_int, c_void_p]),
    "mavlm_debug_force_gemm_bn": (c_int, [c_int]),
    "mavlm_debug_set_flags": (c_int, [c_int]),
    "mavlm_debug_force_attn_groups": (c_int, [c_int]),
    "mavlm_debug_attn_trace": (c_int, [c_void_p]),
    "mavlm_debug_attn_tc_trace": (c_int, [c_void_p]),
}

_lib = None


def load() -> ctypes.CDLL:
    """dlopen libmavlm.so and bind every prototype.  Fails loudly when the CUDA extension is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: the CUDA extension is not built.  Run `python -c 'import __graft_entry__ as g; "
            "g.build()'` (needs nvcc); there is no CPU / PyTorch fallback for this path.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int, what: str = "") -> None:
    if status == 0:
        return
    msg = load().mavlm_last_error_string().decode("utf-8", "replace")
    if status == E_INDEX:
        raise ValueError(msg)
    raise MavlmError(f"{what}: {msg} (status {status})" if what else f"{msg} (status {status})")
