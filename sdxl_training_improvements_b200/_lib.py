"""ctypes binding of libsdxl_b200.so (include/sdxl_b200.h).  No CPU fallback: a missing library raises."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libsdxl_b200.so")

c_p = C.c_void_p
i32, i64, f32, f64, u64 = C.c_int32, C.c_int64, C.c_float, C.c_double, C.c_uint64


class GemmArgs(C.Structure):
    _fields_ = [
        ("A", c_p), ("B", c_p), ("D", c_p), ("bias", c_p), ("residual", c_p),
        ("M", i32), ("N", i32), ("K", i32),
        ("nb_lo", i32), ("nb_hi", i32),
        ("a_mn", i32), ("b_mn", i32),
        ("lda", i64), ("ldb", i64), ("ldd", i64), ("ldr", i64),
        ("a_bs_lo", i64), ("a_bs_hi", i64), ("b_bs_lo", i64), ("b_bs_hi", i64),
        ("d_bs_lo", i64), ("d_bs_hi", i64), ("r_bs_lo", i64), ("r_bs_hi", i64),
        ("alpha", f32), ("accumulate", i32), ("out_fp32", i32),
        ("bias_rows_per_group", i32), ("bias_group_stride", i64),
        ("tile_n", i32), ("allow_split_k", i32),
    ]


class AttnArgs(C.Structure):
    _fields_ = [
        ("Q", c_p), ("K", c_p), ("V", c_p), ("O", c_p), ("LSE", c_p), ("dO", c_p), ("D", c_p),
        ("dQ", c_p), ("dK", c_p), ("dV", c_p),
        ("B", i32), ("H", i32), ("n_q", i32), ("n_k", i32),
        ("ldq", i64), ("ldk", i64), ("ldv", i64), ("ldo", i64), ("lddo", i64), ("lddq", i64), ("lddk", i64), ("lddv", i64),
        ("q_bs", i64), ("k_bs", i64), ("v_bs", i64), ("o_bs", i64), ("do_bs", i64), ("dq_bs", i64), ("dk_bs", i64),
        ("dv_bs", i64),
        ("scale", f32), ("flags", i32),
    ]


class ConvArgs(C.Structure):
    _fields_ = [
        ("x", c_p), ("w", c_p), ("y", c_p), ("bias", c_p), ("residual", c_p),
        ("mode", i32), ("B", i32), ("H", i32), ("W", i32), ("Cin", i32), ("Cout", i32),
        ("ldx", i64), ("ldy", i64), ("ldr", i64),
        ("bias_per_sample", i32), ("accumulate", i32),
    ]


# name -> argtypes (return type is int unless listed in _RESTYPES); mirrors include/sdxl_b200.h one to one.
SIGNATURES = {
    "b2_version": [],
    "b2_last_error": [],
    "b2_launch_count": [],
    "b2_gemm": [C.POINTER(GemmArgs), c_p],
    "b2_attn_lse_rows": [i32],
    "b2_attn_fwd": [C.POINTER(AttnArgs), c_p],
    "b2_attn_bwd": [C.POINTER(AttnArgs), c_p],
    "b2_attn_set_debug": [c_p],
    "b2_xattn_bwd_ok": [i32, i32, i32, i32],
    "b2_xattn_bwd": [C.POINTER(AttnArgs), c_p],
    "b2_xattn_q_core_ok": [i32, i32, i32, i32],
    "b2_linear_dgrad_geglu_ok": [i32, i32, i32],
    "b2_linear_dgrad_geglu": [c_p, c_p, c_p, c_p, i32, i32, i32, i64, i64, i64, i64, c_p],
    "b2_gemm2_set_debug": [c_p],
    "b2_xattn_set_debug": [c_p],
    "b2_xattn_q_core": [c_p, c_p, c_p, c_p, c_p, c_p, c_p, i32, i32, i32, i32, i64, i64, i64, i64, i64, i64, i64, i64, f32, c_p],
    "b2_conv3x3_implicit_ok": [i32, i32, i32, i32, i32],
    "b2_conv3x3": [C.POINTER(ConvArgs), c_p],
    "b2_im2col3x3": [c_p, c_p, i32, i32, i32, i32, i32, i32, i64, c_p],
    "b2_col2im3x3": [c_p, c_p, i32, i32, i32, i32, i32, i32, i64, i32, c_p],
    "b2_gn_stats": [c_p, i32, i32, i32, i32, f32, c_p, c_p, c_p, c_p],
    "b2_gn_apply": [c_p, c_p, c_p, c_p, c_p, c_p, i32, i32, i32, i32, i32, c_p],
    "b2_gn_bwd": [c_p, c_p, c_p, c_p, c_p, c_p, c_p, i32, i32, i32, i32, i32, c_p, c_p, i32, c_p],
    "b2_ln_fwd": [c_p, c_p, c_p, c_p, c_p, c_p, i32, i32, f32, c_p],
    "b2_ln_bwd": [c_p, c_p, c_p, c_p, c_p, c_p, c_p, i32, i32, i32, c_p],
    "b2_ln_bwd_parts": [c_p, c_p, c_p, c_p, c_p, c_p, c_p, i32, i32, i32, i32, c_p],
    "b2_softmax_fwd": [c_p, c_p, i64, i32, i64, i64, c_p],
    "b2_softmax_bwd": [c_p, c_p, c_p, i64, i32, i64, i64, f32, c_p],
    "b2_geglu_fwd": [c_p, c_p, i64, i32, c_p],
    "b2_linear_geglu_ok": [i32, i32, i32],
    "b2_linear_geglu": [c_p, c_p, c_p, c_p, c_p, i32, i32, i32, i64, i64, i64, i64, c_p],
    "b2_geglu_bwd": [c_p, c_p, c_p, i64, i32, c_p],
    "b2_geglu_bwd_bias": [c_p, c_p, c_p, c_p, i64, i32, c_p],
    "b2_silu_fwd": [c_p, c_p, i64, c_p],
    "b2_silu_bwd": [c_p, c_p, c_p, i64, i32, c_p],
    "b2_add": [c_p, c_p, c_p, i64, c_p],
    "b2_copy2d": [c_p, c_p, i64, i64, i64, i64, i32, c_p],
    "b2_copy2d_any": [c_p, c_p, i64, i64, i64, i64, i32, c_p],
    "b2_colsum": [c_p, c_p, i64, i32, i64, i32, c_p, c_p],
    "b2_accum_f32_to_bf16": [c_p, c_p, i64, i32, c_p],
    "b2_colsum_f32": [c_p, c_p, i64, i32, i64, c_p],
    "b2_colsum_groups": [c_p, c_p, i32, i64, i32, i64, i32, c_p, c_p],
    "b2_flush_small_grads": [c_p, c_p, c_p, i32, c_p],
    "b2_nchw_to_nhwc": [c_p, i32, c_p, i32, i32, i32, i32, c_p],
    "b2_nhwc_to_nchw": [c_p, c_p, i32, i32, i32, i32, i32, c_p],
    "b2_timestep_embedding": [c_p, c_p, i32, i32, i64, c_p],
    "b2_randn": [c_p, i64, c_p, u64, i32, c_p],
    "b2_philox_advance": [c_p, u64, c_p],
    "b2_make_noisy": [c_p, c_p, c_p, i32, i32, i32, c_p, c_p, i32, i32, i32, i32, c_p],
    "b2_mse_loss": [c_p, c_p, c_p, c_p, c_p, f32, i32, i32, i32, i32, c_p],
    "b2_finalize_loss": [c_p, f64, f32, c_p, c_p, c_p, i64, c_p],
    "b2_abs_sq_sums": [c_p, i32, i64, i32, i32, c_p, c_p],
    "b2_scale_bf16": [c_p, i64, c_p, f32, c_p],
    "b2_sumsq": [c_p, i64, c_p, c_p],
    "b2_adamw_bf16": [c_p, c_p, c_p, c_p, c_p, i64, f64, f64, f64, f64, i32, c_p, f32, f32, c_p, i32, i32, c_p, i32, c_p],
    "b2_axpy_bf16": [c_p, c_p, i64, f32, c_p],
    "b2_adamw_denom_test": [c_p, c_p, c_p, i32, f32, c_p],
    "b2_dpx_ipc_export": [c_p, c_p, C.POINTER(i64)],
    "b2_dpx_ipc_import": [c_p, i64, C.POINTER(c_p)],
    "b2_dpx_alloc_flags": [C.POINTER(c_p)],
    "b2_dpx_max_chunks": [],
    "b2_dpx_create": [i32, i32, i32, C.POINTER(c_p), C.POINTER(c_p), C.POINTER(c_p), i64, i32, C.POINTER(c_p)],
    "b2_dpx_exchange": [c_p, i32, C.c_uint32, i32, C.POINTER(i64), C.POINTER(i64), i64, c_p],
    "b2_dpx_memcpy_async": [c_p, c_p, i64, c_p],
    "b2_dpx_finish": [c_p, C.c_uint32, c_p],
    "b2_dpx_finish_norm": [c_p, C.c_uint32, c_p, c_p],
    "b2_dpx_norm_supported": [c_p],
    "b2_dpx_destroy": [c_p],
    "b2_adamw": [c_p, c_p, c_p, c_p, c_p, i64, f32, f32, f32, f32, f32, i32, c_p, f32, f32, c_p, c_p],
}
_RESTYPES = {"b2_last_error": C.c_char_p, "b2_launch_count": C.c_longlong}

_lib = None


def load():
    """Load the shared library (once).  Raises RuntimeError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU / PyTorch fallback for the sm_100a kernels)")
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, C.c_int)
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().b2_last_error()
        raise RuntimeError(f"libsdxl_b200 {what} failed ({rc}): {msg.decode() if msg else ''}")


def launch_count() -> int:
    return int(load().b2_launch_count())
