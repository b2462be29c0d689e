"""ctypes binding of libstorygen_b200.so (the C ABI declared in include/storygen_b200.h).

The library is built in-tree by `__graft_entry__.build()` / `make -C storygen_b200/csrc`.  There is no
fallback: if the shared object is missing or its ABI does not match, importing the ops fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# STORYGEN_B200_LIB: developer override for A/B runs of an alternative in-tree build of the same library
LIB_PATH = os.environ.get("STORYGEN_B200_LIB") or os.path.join(_HERE, "libstorygen_b200.so")

SG_ABI_VERSION = 9
SG_MAX_SRC = 4
SG_MAX_SEGS = 24
SG_EPI_LINEAR = 0
SG_EPI_GEGLU = 1
SG_EPI_NCHW_F32 = 2

EXPORTS = [
    "sg_abi_version", "sg_sizeof_gemm_args", "sg_sizeof_attn_args", "sg_launch_count", "sg_error_string",
    "sg_gemm", "sg_attention", "sg_groupnorm_ws_floats", "sg_groupnorm_stats", "sg_groupnorm_apply",
    "sg_layernorm", "sg_timestep_features", "sg_small_linear", "sg_upsample2x", "sg_im2col3x3_nchw",
    "sg_add_noise", "sg_cfg_ddim_step", "sg_gemm_row_stat_parts", "sg_groupnorm_finalize", "sg_groupnorm_apply_ms",
    "sg_attention_bwd",
]


class SgSrc(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("dim_c", C.c_int64), ("dim_w", C.c_int64), ("dim_h", C.c_int64),
                ("dim_n", C.c_int64), ("stride_w", C.c_int64), ("stride_h", C.c_int64), ("stride_n", C.c_int64)]


class SgSeg(C.Structure):
    _fields_ = [("src", C.c_int32), ("dx", C.c_int32), ("dy", C.c_int32), ("c0", C.c_int32), ("nkb", C.c_int32)]


class SgGemmArgs(C.Structure):
    _fields_ = [("num_src", C.c_int32), ("num_segs", C.c_int32), ("src", SgSrc * SG_MAX_SRC),
                ("segs", SgSeg * SG_MAX_SEGS), ("out_w", C.c_int32), ("out_h", C.c_int32), ("out_n", C.c_int32),
                ("bw", C.c_int32), ("bh", C.c_int32), ("bn", C.c_int32), ("b", C.c_void_p), ("N", C.c_int64),
                ("Ktot", C.c_int64), ("mode", C.c_int32), ("out_channels", C.c_int32), ("res_scale", C.c_float), ("bias", C.c_void_p),
                ("lane_bias", C.c_void_p), ("ld_lane_bias", C.c_int64), ("residual", C.c_void_p),
                ("ldr", C.c_int64), ("out", C.c_void_p), ("ldo", C.c_int64), ("debug_timeline", C.c_void_p),
                ("row_stats_out", C.c_void_p), ("ln_stats", C.c_void_p), ("ln_csum", C.c_void_p),
                ("ln_parts", C.c_int32), ("ln_eps", C.c_float), ("col_stats_out", C.c_void_p),
                ("out_stride_w", C.c_int64), ("out_stride_h", C.c_int64), ("out_stride_n", C.c_int64)]


class SgGnParts(C.Structure):
    _fields_ = [("parts", C.c_void_p), ("c", C.c_int32), ("tiles_xy", C.c_int32), ("imgs_per_tile", C.c_int32),
                ("reps", C.c_int32), ("m_tiles", C.c_int32)]


class SgAttnArgs(C.Structure):
    _fields_ = [("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p), ("o", C.c_void_p),
                ("q_stride_n", C.c_int64), ("q_stride_t", C.c_int64), ("k_stride_n", C.c_int64),
                ("k_stride_t", C.c_int64), ("v_stride_n", C.c_int64), ("v_stride_t", C.c_int64),
                ("o_stride_n", C.c_int64), ("o_stride_t", C.c_int64), ("n_lanes", C.c_int32),
                ("n_heads", C.c_int32), ("head_dim", C.c_int32), ("nq", C.c_int32), ("nk", C.c_int32),
                ("n_kv_lanes", C.c_int32), ("kv_map", C.c_void_p), ("scale", C.c_float),
                ("debug_timeline", C.c_void_p), ("lse", C.c_void_p)]


class SgAttnBwdArgs(C.Structure):
    _fields_ = [("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p), ("o", C.c_void_p), ("d_o", C.c_void_p),
                ("lse", C.c_void_p), ("delta", C.c_void_p), ("dq_accum", C.c_void_p), ("dk", C.c_void_p), ("dv", C.c_void_p),
                ("q_stride_n", C.c_int64), ("q_stride_t", C.c_int64), ("k_stride_n", C.c_int64), ("k_stride_t", C.c_int64),
                ("v_stride_n", C.c_int64), ("v_stride_t", C.c_int64), ("o_stride_n", C.c_int64), ("o_stride_t", C.c_int64),
                ("dk_stride_n", C.c_int64), ("dk_stride_t", C.c_int64), ("n_lanes", C.c_int32), ("n_heads", C.c_int32),
                ("head_dim", C.c_int32), ("nq", C.c_int32), ("nk", C.c_int32), ("scale", C.c_float)]


class LibraryError(RuntimeError):
    pass


_lib = None


def load():
    """Load (once) and type the shared library.  Raises LibraryError when it is absent or stale."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LibraryError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU or PyTorch fallback for the StoryGen hot path)")
    lib = C.CDLL(LIB_PATH)
    for name in EXPORTS:
        if not hasattr(lib, name):
            raise LibraryError(f"{LIB_PATH} does not export {name}")
    lib.sg_launch_count.restype = C.c_int64
    lib.sg_error_string.restype = C.c_char_p
    lib.sg_error_string.argtypes = [C.c_int]
    if lib.sg_abi_version() != SG_ABI_VERSION:
        raise LibraryError(f"ABI mismatch: library {lib.sg_abi_version()} vs binding {SG_ABI_VERSION}; rebuild")
    if lib.sg_sizeof_gemm_args() != C.sizeof(SgGemmArgs) or lib.sg_sizeof_attn_args() != C.sizeof(SgAttnArgs):
        raise LibraryError("argument struct size mismatch between the library and the ctypes mirror")
    vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
    lib.sg_gemm.argtypes = [C.POINTER(SgGemmArgs), vp]
    lib.sg_gemm_row_stat_parts.argtypes = [C.POINTER(SgGemmArgs)]
    lib.sg_attention.argtypes = [C.POINTER(SgAttnArgs), vp]
    lib.sg_attention_bwd.argtypes = [C.POINTER(SgAttnBwdArgs), vp]
    lib.sg_groupnorm_ws_floats.argtypes = [i32, i32]
    lib.sg_groupnorm_stats.argtypes = [vp, i32, vp, i32, i32, i32, i32, f32, vp, vp]
    lib.sg_groupnorm_apply.argtypes = [vp, i32, vp, i32, i32, i32, i32, f32, vp, vp, vp, i32, vp, vp]
    lib.sg_groupnorm_finalize.argtypes = [C.POINTER(SgGnParts), C.POINTER(SgGnParts), i32, i32, i32, f32, vp, vp]
    lib.sg_groupnorm_apply_ms.argtypes = [vp, i32, vp, i32, i32, i32, i32, vp, vp, vp, i32, vp, vp]
    lib.sg_layernorm.argtypes = [vp, i64, i32, f32, vp, vp, vp, vp, vp, vp, vp]
    lib.sg_timestep_features.argtypes = [vp, i32, i32, vp, vp]
    lib.sg_small_linear.argtypes = [vp, i32, i32, vp, vp, i32, i32, i32, vp, vp]
    lib.sg_upsample2x.argtypes = [vp, i32, i32, i32, i32, vp, vp]
    lib.sg_im2col3x3_nchw.argtypes = [vp, i32, i32, i32, i32, vp, vp]
    lib.sg_add_noise.argtypes = [vp, vp, vp, vp, i32, i32, vp, vp]
    lib.sg_cfg_ddim_step.argtypes = [vp, vp, i32, i32, f32, f32, vp, vp]
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().sg_error_string(rc).decode()
        raise RuntimeError(f"{what} failed: {msg} (code {rc})")


def launch_count() -> int:
    return int(load().sg_launch_count())
