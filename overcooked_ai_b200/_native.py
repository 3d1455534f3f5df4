"""ctypes binding of the C ABI in include/ovc_b200.h (csrc/libovc_b200.so).

There is no CPU fallback anywhere in this package: if the CUDA library is missing or a CUDA
device is not available, the calls below raise.  Build it with ``python -m overcooked_ai_b200.build``
(or ``__graft_entry__.build()``).
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("OVC_B200_LIB") or os.path.join(_HERE, "csrc", "libovc_b200.so")  # OVC_B200_LIB: an experiment build (build.py)

ABI_VERSION = 5
F_AUTO_RESET = 1
F_PDL = 2
F_ACT_U8 = 4
F_OUT_NARROW = 8
F_OUT_PACKED = 16
F_OUT_CODES = 32
F_ACT_PACKED = 64
F_OUT_STREAM = 128
F_STREAM_CAP_SHIFT = 16
STREAM_CAP_MAX = 0xFFFF
F_IO_SHIFT = 8
IO_DEFAULT, IO_TMA_TENSOR, IO_TMA_BULK, IO_DIRECT = 0, 1, 2, 3
DT_F32, DT_U8, DT_I32, DT_BF16 = 0, 1, 2, 3



class RandomStart(ctypes.Structure):
    """ovc_random_start_t (include/ovc_b200.h): random start states, passed by host pointer."""
    _fields_ = [("seed", ctypes.c_uint64), ("obj_threshold", ctypes.c_uint32), ("random_start_pos", ctypes.c_int32),
                ("random_layout", ctypes.c_int32), ("reserved", ctypes.c_int32)]


class PipelineDesc(ctypes.Structure):
    """ovc_pipeline_desc_t (include/ovc_b200.h)"""
    _fields_ = [("layouts", ctypes.c_void_p), ("n_layouts", ctypes.c_int32), ("state_words", ctypes.c_int32),
                ("start_records", ctypes.c_void_p), ("state", ctypes.c_void_p), ("n_envs", ctypes.c_int64),
                ("horizon", ctypes.c_int32), ("flags", ctypes.c_int32), ("chunk", ctypes.c_int32),
                ("has_random_start", ctypes.c_int32), ("random_start", RandomStart),
                ("d_actions", ctypes.c_void_p * 2), ("d_sparse", ctypes.c_void_p * 2), ("d_shaped", ctypes.c_void_p * 2),
                ("d_done", ctypes.c_void_p * 2), ("d_events", ctypes.c_void_p * 2),
                ("stream_cap", ctypes.c_int32), ("reserved", ctypes.c_int32), ("d_codes_full", ctypes.c_void_p * 2)]


_lib = None


class NativeLibraryError(RuntimeError):
    pass


def lib():
    """Load (once) and return the CUDA library; raises NativeLibraryError if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryError(
            "%s not found: the CUDA extension is not built (python -m overcooked_ai_b200.build). "
            "This engine has no CPU fallback." % LIB_PATH
        )
    L = ctypes.CDLL(LIB_PATH)
    vp, i32, i64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64
    L.ovc_abi_version.restype = i32
    L.ovc_layout_table_size.restype = ctypes.c_size_t
    L.ovc_feat_lut_entry_size.restype = ctypes.c_size_t
    L.ovc_last_error.restype = ctypes.c_char_p
    L.ovc_step.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp, vp, i64, i32, i32, i32, vp, vp]
    L.ovc_rollout.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp, vp, i64, i32, i32, i32, i32, vp, vp]
    L.ovc_reset.argtypes = [vp, i32, vp, vp, vp, vp, i64, i32, vp, vp]
    L.ovc_encode_lossless.argtypes = [vp, i32, vp, vp, vp, i32, i64, i32, i32, i32, i32, vp]
    L.ovc_encode_linear.argtypes = [vp, i32, vp, vp, vp, vp, vp, i64, i32, i32, i32, i32, i32, ctypes.c_float, vp]
    L.ovc_sample_actions.argtypes = [vp, i32, i32, i64, ctypes.c_uint64, vp, vp, vp]
    L.ovc_accumulate_returns.argtypes = [vp, vp, ctypes.c_float, i64, vp, vp, vp]
    L.ovc_policy_tail.argtypes = [vp, i64, i32, ctypes.c_float, vp, vp, vp, vp, i32, vp, vp, ctypes.c_float, i32, ctypes.c_uint64, vp, vp, vp, vp, vp]
    L.ovc_wide_layers.argtypes = [vp, i64, i32, vp, vp, i32, vp, vp, i32, ctypes.c_float, vp, vp]
    L.ovc_featurize.argtypes = [vp, i32, vp, vp, vp, vp, i64, i32, i32, vp]
    L.ovc_potential.argtypes = [vp, i32, vp, vp, vp, i32, vp, vp, i64, i32, vp]
    L.ovc_potential_table_size.restype = ctypes.c_size_t
    L.ovc_expand_codes_host.argtypes = [vp, i64, i64, vp, vp, i32, vp, vp, vp, vp, i32]
    L.ovc_expand_stream_host.argtypes = [vp, vp, i64, i64, i64, i64, vp, vp, i32, vp, vp, vp, vp, i32, ctypes.POINTER(i64)]
    L.ovc_pipeline_create.argtypes = [ctypes.POINTER(PipelineDesc), ctypes.POINTER(vp)]
    L.ovc_pipeline_run.argtypes = [vp, vp, vp, vp, vp, vp, i32, vp, i32, ctypes.POINTER(i64)]
    L.ovc_pipeline_wait.argtypes = [vp, i64]
    L.ovc_pipeline_join.argtypes = [vp, vp]
    L.ovc_pipeline_destroy.argtypes = [vp]
    L.ovc_pipeline_destroy.restype = None
    for f in (L.ovc_step, L.ovc_rollout, L.ovc_reset, L.ovc_encode_lossless, L.ovc_encode_linear, L.ovc_sample_actions, L.ovc_accumulate_returns, L.ovc_policy_tail, L.ovc_wide_layers, L.ovc_featurize, L.ovc_potential,
              L.ovc_expand_codes_host, L.ovc_expand_stream_host, L.ovc_pipeline_create, L.ovc_pipeline_run, L.ovc_pipeline_wait, L.ovc_pipeline_join):
        f.restype = i32
    if L.ovc_abi_version() != ABI_VERSION:
        raise NativeLibraryError("ABI version mismatch: library %d, binding %d" % (L.ovc_abi_version(), ABI_VERSION))
    _lib = L
    return L


EXPORTED_SYMBOLS = (
    "ovc_abi_version", "ovc_layout_table_size", "ovc_feat_lut_entry_size", "ovc_last_error",
    "ovc_step", "ovc_rollout", "ovc_reset", "ovc_encode_lossless", "ovc_encode_linear", "ovc_sample_actions", "ovc_accumulate_returns", "ovc_policy_tail", "ovc_wide_layers", "ovc_featurize", "ovc_potential",
    "ovc_potential_table_size", "ovc_expand_codes_host", "ovc_expand_stream_host",
    "ovc_pipeline_create", "ovc_pipeline_run", "ovc_pipeline_wait", "ovc_pipeline_join", "ovc_pipeline_destroy",
)


def check(rc):
    if rc != 0:
        raise RuntimeError("ovc native call failed (%d): %s" % (rc, lib().ovc_last_error().decode()))
