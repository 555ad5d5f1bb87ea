"""ctypes binding of libmtsb200.so — the C ABI declared in include/mts_b200.h.

This is the reference-side binding a maintainer would add: the reference is
Python, so the FFI is ctypes.  There is no fallback: if the library cannot be loaded, importing any
compute entry point raises, and every non-zero status becomes an `MtsError`.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import os

# MTS_LIB_PATH: load another build of the library (A/B timing of kernel changes, tools/bench_attn.py)
LIB_PATH = Path(os.environ.get("MTS_LIB_PATH") or Path(__file__).resolve().parent / "lib" / "libmtsb200.so")

MTS_OK = 0
MTS_BF16, MTS_F32 = 0, 1
EPI_STORE, EPI_RESID_ADD, EPI_GELU_NEW, EPI_SWIGLU, EPI_ROPE_QK = 0, 1, 2, 3, 4
BIAS_NONE, BIAS_N, BIAS_M = 0, 1, 2


class MtsError(RuntimeError):
    """A libmtsb200 entry point returned a non-zero status."""


class GemmArgs(C.Structure):
    _fields_ = [
        ("a", C.c_void_p), ("b", C.c_void_p), ("d", C.c_void_p), ("bias", C.c_void_p), ("c", C.c_void_p),
        ("lda", C.c_int64), ("ldb", C.c_int64), ("ldd", C.c_int64),
        ("a_batch_stride", C.c_int64), ("b_batch_stride", C.c_int64), ("d_batch_stride", C.c_int64),
        ("m", C.c_int32), ("n", C.c_int32), ("k", C.c_int32), ("batch", C.c_int32),
        ("d_dtype", C.c_int32), ("epilogue", C.c_int32), ("bias_axis", C.c_int32),
        ("d_transposed", C.c_int32), ("block_n", C.c_int32), ("alpha", C.c_float),
        ("rope_cos", C.c_void_p), ("rope_sin", C.c_void_p),
        ("rope_L", C.c_int32), ("rope_hd", C.c_int32), ("rope_cols", C.c_int32), ("rope_prefix", C.c_int32),
        ("aux", C.c_void_p), ("ld_aux", C.c_int64),
        ("ab_dtype", C.c_int32), ("round_tf32", C.c_int32),
        ("a_lo", C.c_void_p), ("b_lo", C.c_void_p),
        ("drop_p", C.c_float), ("drop_seed", C.c_uint64),
        ("sk_workspace", C.c_void_p), ("sk_workspace_bytes", C.c_int64), ("sk_flags", C.c_void_p),
        ("sk_flags_len", C.c_int32), ("sk_epoch", C.c_int32),
    ]


_p, _i, _i64, _f = C.c_void_p, C.c_int, C.c_int64, C.c_float

# name -> argtypes; every function returns int status unless listed in _SPECIAL
SIGNATURES = {
    "mts_revin_patch_embed": [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _f, _p],
    "mts_patch_gather": [_p, _p, _i, _i, _i, _i, _i, _p],
    "mts_revin_patch_embed_bwd": [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p],
    "mts_revin_denorm": [_p, _p, _p, _i, _i, _i, _p],
    "mts_gemm": [C.POINTER(GemmArgs), _p],
    "mts_pack_gate_up": [_p, _p, _p, _i, _i, _p],
    "mts_cast_f32_bf16": [_p, _p, _i64, _p],
    "mts_cast_bf16_f32": [_p, _p, _i64, _p],
    "mts_transpose_f32_bf16": [_p, _p, _i, _i, _p],
    "mts_transpose_bf16": [_p, _p, _i, _i, _p],
    "mts_rmsnorm": [_p, _i64, _p, _p, _p, _i, _i, _f, _p],
    "mts_layernorm": [_p, _i64, _p, _p, _p, _p, _i, _i, _f, _p],
    "mts_attn_causal": [_p, _p, _p, _p, _p, _i, _i, _i, _i, _f, _p],
    "mts_softmax_rows": [_p, _p, _i64, _i, _f, _p],
    "mts_prompt_gather": [_p, _p, _p, _p, _i, _i, _i, _i, _i, _p],
    "mts_swiglu": [_p, _i64, _p, _i64, _i, _p],
    "mts_dropout": [_p, _p, _i, _i64, _f, C.c_uint64, _p],
    "mts_sigmoid": [_p, _i64, _p],
    "mts_softmax_lastdim": [_p, _i64, _i, _p],
    "mts_rmsnorm_bwd": [_p, _i64, _p, _p, _p, _p, _i, _i, _f, _i, _p],
    "mts_layernorm_bwd": [_p, _i64, _p, _p, _p, _p, _i, _i, _f, _i, _p],
    "mts_attn_causal_bwd": [_p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _f, _i, _p],
    "mts_rope_qk": [_p, _p, _p, _i, _i, _i, _i, _p],
    "mts_swiglu_blk": [_p, _i64, _p, _i64, _i, _i, _p],
    "mts_swiglu_bwd": [_p, _i64, _p, _p, _i64, _i, _i, _p],
    "mts_gelu_new": [_p, _p, _p, _i64, _p],
    "mts_softmax_bwd_rows": [_p, _p, _p, _i64, _i, _f, _p],
    "mts_colsum": [_p, _i, _i64, _p, _i, _i, _p],
    "mts_rowsum_f32": [_p, _i64, _p, _i, _i, _p],
    "mts_transpose_strided": [_p, _i, _i64, _i64, _p, _i64, _i, _i, _i, _p],
    "mts_cast_rows_f32_bf16": [_p, _i64, _i64, _p, _i64, _i64, _i, _i, _i, _p],
    "mts_group_reduce": [_p, _p, _p, _p, _i64, _i, _i, _i64, _i, _p],
    "mts_group_reduce_bwd": [_p, _i64, _p, _p, _p, _p, _p, _i, _i, _i64, _p],
    "mts_merge_end": [_p, _p, _p, _p, _i, _i, _i, _i, _p],
    "mts_merge_end_bwd": [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p],
    "mts_revin_denorm_bwd": [_p, _p, _p, _i, _i, _i, _p],
    "mts_attn_causal_shared": [_p, _p, _p, _i, _i, _i, _i, _i, _f, _p],
    "mts_attn_causal_shared_bwd": [_p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _f, _p],
    "mts_attn_causal_shared_bwd_full": [_p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _f, _p],
    "mts_rope_qk_shared": [_p, _p, _p, _i, _i, _i, _i, _i, _p],
    "mts_prompt_gather_shared": [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p],
    "mts_gpt4ts_embed": [_p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _f, _p],
    "mts_gpt4ts_conv_wgrad": [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i64, _i64, _i64, _p],
    "mts_norm_wgrad_partial": [_p, _i64, _p, _p, _i, _i, _f, _i, _p],
    "mts_round_tf32": [_p, _p, _i64, _p],
    "mts_split_tf32": [_p, _p, _p, _i64, _p],
    "mts_softmax_rows_f32": [_p, _p, _i64, _i, _f, _p],
    "mts_attn_causal_f32": [_p, _p, _i, _i, _i, _i, _i, _f, _i, _p],
    "mts_attn_causal_tf32": [_p, _p, _i, _i, _i, _i, _i, _f, _i, _p],
    "mts_input_stats": [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p],
    "mts_attn_causal_dropout": [_p, _p, _p, _i, _i, _i, _i, _f, _f, C.c_uint64, _p],
    "mts_attn_causal_dropout_bwd": [_p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _f, _f, C.c_uint64, _p],
    "mts_clear_caches": [],
    "mts_set_option": [C.c_char_p, _i],
}
_SPECIAL = {
    "mts_version": ([], C.c_int),
    "mts_last_error": ([], C.c_char_p),
    "mts_launch_count": ([], C.c_int64),
}

EXPORTED_SYMBOLS = sorted(list(SIGNATURES) + list(_SPECIAL))

_lib = None


def load() -> C.CDLL:
    """Loads the shared library (raises MtsError with the build hint if it is missing)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise MtsError(
            f"{LIB_PATH} not found: the CUDA extension is not built. Run `python -c 'import "
            f"__graft_entry__ as g; g.build()'` (needs nvcc). There is no CPU fallback.")
    lib = C.CDLL(str(LIB_PATH))
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int
    for name, (argtypes, restype) in _SPECIAL.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = restype
    _lib = lib
    return lib


def call(name: str, *args) -> None:
    """Calls a status-returning entry point; raises MtsError on failure."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != MTS_OK:
        msg = lib.mts_last_error().decode("utf-8", "replace")
        raise MtsError(f"{name} failed (status {rc}): {msg}")


_replayed = 0


def launch_count() -> int:
    """Kernels launched by the library so far: launches it issued itself (counted in C, including the ones
    recorded into a CUDA graph at capture time) plus the kernel nodes of every graph replay since."""
    return int(load().mts_launch_count()) + _replayed


def note_replay(n_kernels: int) -> None:
    global _replayed
    _replayed += int(n_kernels)


def version() -> int:
    return int(load().mts_version())


def set_option(name: str, value: int) -> None:
    call("mts_set_option", name.encode(), int(value))
