"""Tensor-level wrappers over the C ABI: torch owns memory and streams, libmtsb200 does the work.

Every function here launches hand-written sm_100a kernels through ctypes on the current CUDA stream.
Nothing falls back to torch arithmetic: a CPU tensor or a missing library raises.
"""
from __future__ import annotations

import ctypes as C
import math

import torch

from . import _lib
from ._lib import BIAS_N, BIAS_NONE, EPI_STORE, EPI_SWIGLU, MTS_BF16, MTS_F32, GemmArgs, MtsError


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _chk(t: torch.Tensor, dtype=None, name="tensor"):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise MtsError(f"{name} must be a CUDA tensor (libmtsb200 has no CPU path)")
    if dtype is not None and t.dtype != dtype:
        raise MtsError(f"{name} must be {dtype}, got {t.dtype}")
    return t


def _ptr(t):
    return 0 if t is None else t.data_ptr()


# ------------------------------------------------------------------------------------------------
# GEMM
# ------------------------------------------------------------------------------------------------
_SK_SCRATCH: dict = {}
_STREAMK_MODE = 0


def set_streamk(mode: int) -> None:
    """0 off (default) | 1 auto | 2 whenever legal — mts_set_option("streamk") plus the scratch the library needs for it
    (only lent while the mode is on: experimental, see profiles/r02_streamk_gemm.md)."""
    global _STREAMK_MODE
    _lib.set_option("streamk", int(mode))
    _STREAMK_MODE = int(mode)


def _streamk_scratch(device, stream):
    """Scratch the library may use for stream-K (include/mts_b200.h, mts_gemm_args.sk_workspace): one partial-tile slot
    and one flag per SM, per (device, stream) — launches on one stream are ordered, so they can share it."""
    key = (device.index, stream)
    hit = _SK_SCRATCH.get(key)
    if hit is None:
        sms = torch.cuda.get_device_properties(device).multi_processor_count
        hit = (torch.empty(sms * 128 * 256, device=device, dtype=torch.float32),
               torch.zeros(sms, device=device, dtype=torch.int32))
        _SK_SCRATCH[key] = hit
    return hit


def gemm(a, b, d, *, m, n, k, batch=1, lda=None, ldb=None, ldd=None, a_bs=0, b_bs=0, d_bs=0,
         bias=None, bias_axis=BIAS_NONE, epilogue=EPI_STORE, alpha=1.0, d_transposed=False,
         block_n=0, a_off=0, b_off=0, d_off=0, c=None, rope=None, rope_L=0, rope_hd=0, rope_cols=0,
         rope_prefix=0, aux=None, a_lo=None, b_lo=None, round_tf32=False, drop_p=0.0, drop_seed=0):
    """Raw mts_gemm: D[b] = epi(alpha * A[b] @ B[b]^T + bias).  Offsets/strides in elements.
    a / b bf16 (default path) or both fp32 (evaluation parity modes: tcgen05 kind::tf32; with `a_lo` / `b_lo` the
    3xTF32 fp32-grade contraction on split operands)."""
    ab = a.dtype
    if ab not in (torch.bfloat16, torch.float32):
        raise MtsError("a must be bf16 or fp32")
    _chk(a, ab, "a"); _chk(b, ab, "b"); _chk(d, None, "d")
    es = a.element_size()
    if d.dtype not in (torch.bfloat16, torch.float32):
        raise MtsError("d must be bf16 or fp32")
    args = GemmArgs()
    args.a = a.data_ptr() + es * a_off
    args.b = b.data_ptr() + es * b_off
    args.d = d.data_ptr() + d.element_size() * d_off
    args.bias = _ptr(bias)
    args.c = 0 if c is None else _chk(c, torch.float32, "c").data_ptr() + 4 * d_off
    if bias is not None:
        _chk(bias, torch.float32, "bias")
    args.lda = k if lda is None else lda
    args.ldb = k if ldb is None else ldb
    n_store = n // 2 if epilogue == EPI_SWIGLU else n
    args.ldd = (m if d_transposed else n_store) if ldd is None else ldd
    args.a_batch_stride, args.b_batch_stride, args.d_batch_stride = a_bs, b_bs, d_bs
    args.m, args.n, args.k, args.batch = m, n, k, batch
    args.d_dtype = MTS_F32 if d.dtype == torch.float32 else MTS_BF16
    args.epilogue = epilogue
    args.bias_axis = bias_axis
    args.d_transposed = 1 if d_transposed else 0
    args.block_n = block_n
    args.alpha = alpha
    args.ab_dtype = MTS_F32 if ab == torch.float32 else MTS_BF16
    args.round_tf32 = 1 if round_tf32 else 0
    args.drop_p, args.drop_seed = float(drop_p), int(drop_seed) & (2 ** 64 - 1)
    stream = _stream()
    if batch == 1 and _STREAMK_MODE:
        ws, flags = _streamk_scratch(d.device, stream)
        args.sk_workspace, args.sk_workspace_bytes = ws.data_ptr(), ws.numel() * 4
        args.sk_flags, args.sk_flags_len, args.sk_epoch = flags.data_ptr(), flags.numel(), 1
    if (a_lo is None) != (b_lo is None):
        raise MtsError("a_lo and b_lo go together (3xTF32 split)")
    if a_lo is not None:
        _chk(a_lo, torch.float32, "a_lo"); _chk(b_lo, torch.float32, "b_lo")
        if a_lo.shape != a.shape or b_lo.shape != b.shape or a_lo.stride() != a.stride() or b_lo.stride() != b.stride():
            raise MtsError("a_lo / b_lo must be laid out like a / b")
        args.a_lo = a_lo.data_ptr() + 4 * a_off
        args.b_lo = b_lo.data_ptr() + 4 * b_off
    if rope is not None:
        args.rope_cos, args.rope_sin = rope[0].data_ptr(), rope[1].data_ptr()
        args.rope_L, args.rope_hd, args.rope_cols, args.rope_prefix = rope_L, rope_hd, rope_cols, rope_prefix
        if rope[0].shape[0] < rope_prefix + rope_L:
            raise MtsError("rope tables shorter than the sequence")
    if aux is not None:
        _chk(aux, torch.bfloat16, "aux")
        args.aux, args.ld_aux = aux.data_ptr(), aux.shape[-1]
    _lib.call("mts_gemm", C.byref(args), stream)
    return d


def linear_bf16(x, w, bias=None, *, out=None, out_dtype=torch.bfloat16, epilogue=EPI_STORE,
                block_n=0):
    """y = x @ w^T (+ bias): x bf16 [..., K] contiguous, w bf16 [N, K] contiguous."""
    _chk(x, torch.bfloat16, "x"); _chk(w, torch.bfloat16, "w")
    if not x.is_contiguous() or not w.is_contiguous():
        raise MtsError("linear_bf16 needs contiguous operands")
    K = x.shape[-1]
    N = w.shape[0]
    M = x.numel() // K
    n_out = N // 2 if epilogue == EPI_SWIGLU else N
    if out is None:
        out = torch.empty(*x.shape[:-1], n_out, device=x.device, dtype=out_dtype)
    gemm(x, w, out, m=M, n=N, k=K, bias=bias, bias_axis=BIAS_N if bias is not None else BIAS_NONE,
         epilogue=epilogue, block_n=block_n)
    return out


# ------------------------------------------------------------------------------------------------
# front end
# ------------------------------------------------------------------------------------------------
def n_patches(T, P, S):
    return (T + S - P) // S + 1


def revin_patch_embed(x, w_conv, P, S, *, concat, eps=1e-5, want_bf16=True, want_f32=False):
    """Fused RevIN-norm + pad + unfold + TokenEmbedding conv.  Returns (out_bf16, out_f32, mean, std)."""
    _chk(x, torch.float32, "x"); _chk(w_conv, torch.float32, "w_conv")
    x = x.contiguous(); w_conv = w_conv.contiguous()
    B, T, Cc = x.shape
    dm = w_conv.shape[0]
    N = n_patches(T, P, S)
    shape = (B, N, Cc * dm) if concat else (B * Cc, N, dm)
    mean = torch.empty(B, Cc, device=x.device, dtype=torch.float32)
    std = torch.empty(B, Cc, device=x.device, dtype=torch.float32)
    ob = torch.empty(shape, device=x.device, dtype=torch.bfloat16) if want_bf16 else None
    of = torch.empty(shape, device=x.device, dtype=torch.float32) if want_f32 else None
    _lib.call("mts_revin_patch_embed", x.data_ptr(), w_conv.data_ptr(), mean.data_ptr(),
              std.data_ptr(), _ptr(ob), _ptr(of), B, T, Cc, P, S, dm, 1 if concat else 0, eps,
              _stream())
    return ob, of, mean, std


def patch_gather(x, P, S):
    _chk(x, torch.float32, "x")
    x = x.contiguous()
    B, T, Cc = x.shape
    N = n_patches(T, P, S)
    out = torch.empty(B * Cc, N, P, device=x.device, dtype=torch.float32)
    _lib.call("mts_patch_gather", x.data_ptr(), out.data_ptr(), B, T, Cc, P, S, _stream())
    return out


def revin_patch_embed_bwd(x, mean, std, dout, P, S, dm, *, concat, out=None):
    _chk(x, torch.float32, "x"); _chk(dout, torch.float32, "dout")
    x = x.contiguous(); dout = dout.contiguous()
    B, T, Cc = x.shape
    dw = torch.empty(dm, P, 3, device=x.device, dtype=torch.float32) if out is None else out
    if dw.shape != (dm, P, 3) or dw.dtype != torch.float32 or not dw.is_contiguous():
        raise MtsError("revin_patch_embed_bwd: out must be contiguous fp32 [d_model, P, 3]")
    _lib.call("mts_revin_patch_embed_bwd", x.data_ptr(), mean.data_ptr(), std.data_ptr(),
              dout.data_ptr(), dw.data_ptr(), B, T, Cc, P, S, dm, 1 if concat else 0, _stream())
    return dw


def revin_denorm(y, mean, std):
    """In-place y[b,t,c] = y*std[b,c] + mean[b,c]."""
    _chk(y, torch.float32, "y")
    if not y.is_contiguous():
        raise MtsError("revin_denorm needs a contiguous tensor")
    B, T, Cc = y.shape
    _lib.call("mts_revin_denorm", y.data_ptr(), mean.data_ptr(), std.data_ptr(), B, T, Cc, _stream())
    return y


# ------------------------------------------------------------------------------------------------
# row kernels
# ------------------------------------------------------------------------------------------------
def rmsnorm(x, w, eps, *, out_dtype=torch.bfloat16, out=None):
    _chk(x, torch.float32, "x"); _chk(w, torch.float32, "w")
    D = x.shape[-1]
    rows = x.numel() // D
    if not x.is_contiguous():
        raise MtsError("rmsnorm needs contiguous input")
    if out is None:
        out = torch.empty(x.shape, device=x.device, dtype=out_dtype)
    yb, yf = (out.data_ptr(), 0) if out.dtype == torch.bfloat16 else (0, out.data_ptr())
    _lib.call("mts_rmsnorm", x.data_ptr(), D, w.data_ptr(), yb, yf, rows, D, eps, _stream())
    return out


def layernorm(x, w, b, eps, *, out_dtype=torch.bfloat16, out=None):
    _chk(x, torch.float32, "x"); _chk(w, torch.float32, "w"); _chk(b, torch.float32, "b")
    D = x.shape[-1]
    rows = x.numel() // D
    if not x.is_contiguous():
        raise MtsError("layernorm needs contiguous input")
    if out is None:
        out = torch.empty(x.shape, device=x.device, dtype=out_dtype)
    yb, yf = (out.data_ptr(), 0) if out.dtype == torch.bfloat16 else (0, out.data_ptr())
    _lib.call("mts_layernorm", x.data_ptr(), D, w.data_ptr(), b.data_ptr(), yb, yf, rows, D, eps,
              _stream())
    return out


def attn_causal(qkv, Bp, L, H, hd, *, rope=None, scale=None, want_lse=False, out=None):
    """qkv bf16 [Bp*L, 3*H*hd] -> out bf16 [Bp*L, H*hd]; rope = (cos, sin) fp32 [L, hd/2] or None."""
    _chk(qkv, torch.bfloat16, "qkv")
    if not qkv.is_contiguous():
        raise MtsError("attn_causal needs contiguous qkv")
    if scale is None:
        scale = 1.0 / math.sqrt(hd)
    if out is None:
        out = torch.empty(Bp * L, H * hd, device=qkv.device, dtype=torch.bfloat16)
    lse = torch.empty(Bp, H, L, device=qkv.device, dtype=torch.float32) if want_lse else None
    cos = sin = None
    if rope is not None:
        cos, sin = rope
        _chk(cos, torch.float32, "rope cos"); _chk(sin, torch.float32, "rope sin")
        if cos.shape[0] < L or cos.shape[1] != hd // 2 or not cos.is_contiguous() or not sin.is_contiguous():
            raise MtsError("rope tables must be contiguous [>=L, hd/2]")
    _lib.call("mts_attn_causal", qkv.data_ptr(), _ptr(cos), _ptr(sin), out.data_ptr(), _ptr(lse),
              Bp, L, H, hd, scale, _stream())
    return (out, lse) if want_lse else out


def attn_causal_dropout(qkv, Bp, L, H, hd, p, seed, *, scale=None, out=None):
    """attn_causal with dropout (probability p, counter-based mask from `seed`) on the attention probabilities; q / k
    already rotated.  Returns (out, lse) — training only."""
    _chk(qkv, torch.bfloat16, "qkv")
    if not qkv.is_contiguous() or qkv.shape != (Bp * L, 3 * H * hd):
        raise MtsError("attn_causal_dropout needs contiguous qkv [Bp*L, 3*H*hd]")
    if scale is None:
        scale = 1.0 / math.sqrt(hd)
    if out is None:
        out = torch.empty(Bp * L, H * hd, device=qkv.device, dtype=torch.bfloat16)
    lse = torch.empty(Bp, H, L, device=qkv.device, dtype=torch.float32)
    _lib.call("mts_attn_causal_dropout", qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), Bp, L, H, hd, scale, float(p),
              int(seed) & (2 ** 64 - 1), _stream())
    return out, lse


def attn_causal_dropout_bwd(qkv, out, dout, lse, Bp, L, H, hd, p, seed, *, rope=None, scale=None):
    _chk(qkv, torch.bfloat16, "qkv"); _chk(out, torch.bfloat16, "out"); _chk(dout, torch.bfloat16, "dout")
    _chk(lse, torch.float32, "lse")
    if scale is None:
        scale = 1.0 / math.sqrt(hd)
    dqkv = torch.empty_like(qkv)
    delta = torch.empty(Bp, H, L, device=qkv.device, dtype=torch.float32)
    cos, sin = rope if rope is not None else (None, None)
    _lib.call("mts_attn_causal_dropout_bwd", qkv.data_ptr(), _ptr(cos), _ptr(sin), out.data_ptr(), dout.data_ptr(),
              lse.data_ptr(), delta.data_ptr(), dqkv.data_ptr(), Bp, L, H, hd, scale, float(p), int(seed) & (2 ** 64 - 1),
              _stream())
    return dqkv


def attn_causal_shared(qkv, Bp, Lc, Ls, H, hd, *, scale=None, want_lse=False, out=None):
    """Causal attention on the shared-prefix row layout (include/mts_b200.h): qkv bf16 [Lc + Bp*Ls, 3*H*hd]
    with q / k already rotated -> out bf16 [Lc + Bp*Ls, H*hd]; lse fp32 [H*Lc + Bp*H*Ls] (prefix [H, Lc] first, then
    the own tokens [Bp, H, Ls]: `lse_own_view`)."""
    _chk(qkv, torch.bfloat16, "qkv")
    M = Lc + Bp * Ls
    if not qkv.is_contiguous() or qkv.shape != (M, 3 * H * hd):
        raise MtsError("attn_causal_shared needs contiguous qkv [Lc + Bp*Ls, 3*H*hd]")
    if scale is None:
        scale = 1.0 / math.sqrt(hd)
    if out is None:
        out = torch.empty(M, H * hd, device=qkv.device, dtype=torch.bfloat16)
    lse = torch.empty(H * Lc + Bp * H * Ls, device=qkv.device, dtype=torch.float32) if want_lse else None
    _lib.call("mts_attn_causal_shared", qkv.data_ptr(), out.data_ptr(), _ptr(lse), Bp, Lc, Ls, H, hd, scale, _stream())
    return (out, lse) if want_lse else out


def lse_own_view(lse, Bp, Lc, Ls, H):
    """The own-token part [Bp, H, Ls] of the lse buffer written by attn_causal_shared."""
    return lse[H * Lc:].view(Bp, H, Ls)


def attn_causal_shared_bwd_full(qkv, out, dout, lse, Bp, Lc, Ls, H, hd, *, rope=None, scale=None):
    """Gradient w.r.t. ALL rows of the (un-rotated) qkv projection on the shared-prefix layout: bf16
    [Lc + Bp*Ls, 3*H*hd] (LoRA training: the prefix rows matter too)."""
    M = Lc + Bp * Ls
    _chk(qkv, torch.bfloat16, "qkv"); _chk(out, torch.bfloat16, "out"); _chk(dout, torch.bfloat16, "dout")
    _chk(lse, torch.float32, "lse")
    for t, cols in ((out, H * hd), (dout, H * hd), (qkv, 3 * H * hd)):
        if not t.is_contiguous() or t.shape != (M, cols):
            raise MtsError("attn_causal_shared_bwd_full: tensors must be contiguous with Lc + Bp*Ls rows")
    if not lse.is_contiguous() or lse.numel() != H * Lc + Bp * H * Ls:
        raise MtsError("attn_causal_shared_bwd_full: lse must be the full buffer of attn_causal_shared")
    if scale is None:
        scale = 1.0 / math.sqrt(hd)
    dqkv = torch.empty_like(qkv)
    delta = torch.empty_like(lse)
    cos, sin = rope if rope is not None else (None, None)
    if cos is not None and cos.shape[0] < Lc + Ls:
        raise MtsError("rope tables shorter than the sequence")
    _lib.call("mts_attn_causal_shared_bwd_full", qkv.data_ptr(), _ptr(cos), _ptr(sin), out.data_ptr(), dout.data_ptr(),
              lse.data_ptr(), delta.data_ptr(), dqkv.data_ptr(), Bp, Lc, Ls, H, hd, scale, _stream())
    return dqkv


def attn_causal_shared_bwd(qkv, out_own, dout_own, lse_own, Bp, Lc, Ls, H, hd, *, rope=None, scale=None):
    """Gradient w.r.t. the own-token rows of the (un-rotated) qkv projection: bf16 [Bp*Ls, 3*H*hd]."""
    _chk(qkv, torch.bfloat16, "qkv"); _chk(out_own, torch.bfloat16, "out"); _chk(dout_own, torch.bfloat16, "dout")
    _chk(lse_own, torch.float32, "lse")
    for t, cols in ((out_own, H * hd), (dout_own, H * hd)):
        if not t.is_contiguous() or t.shape != (Bp * Ls, cols):
            raise MtsError("attn_causal_shared_bwd: own-row tensors must be contiguous [Bp*Ls, H*hd]")
    if not qkv.is_contiguous() or not lse_own.is_contiguous() or lse_own.numel() != Bp * H * Ls:
        raise MtsError("attn_causal_shared_bwd: contiguous qkv and lse [Bp, H, Ls] required")
    if scale is None:
        scale = 1.0 / math.sqrt(hd)
    dqkv = torch.empty(Bp * Ls, 3 * H * hd, device=qkv.device, dtype=torch.bfloat16)
    delta = torch.empty(Bp, H, Ls, device=qkv.device, dtype=torch.float32)
    cos, sin = rope if rope is not None else (None, None)
    if cos is not None and cos.shape[0] < Lc + Ls:
        raise MtsError("rope tables shorter than the sequence")
    _lib.call("mts_attn_causal_shared_bwd", qkv.data_ptr(), _ptr(cos), _ptr(sin), out_own.data_ptr(), dout_own.data_ptr(),
              lse_own.data_ptr(), delta.data_ptr(), dqkv.data_ptr(), Bp, Lc, Ls, H, hd, scale, _stream())
    return dqkv


def softmax_rows(s, scale, out=None):
    _chk(s, torch.float32, "s")
    if not s.is_contiguous():
        raise MtsError("softmax_rows needs contiguous input")
    n = s.shape[-1]
    rows = s.numel() // n
    if out is None:
        out = torch.empty(s.shape, device=s.device, dtype=torch.bfloat16)
    _lib.call("mts_softmax_rows", s.data_ptr(), out.data_ptr(), rows, n, scale, _stream())
    return out


def prompt_gather(ids, emb, wpe, x, *, rep, Lp, L, Lc=0, B=None):
    """Fills x fp32 [B*rep, L, D]: rows <Lp from emb[ids] (+wpe), rows >=Lp zero (+wpe).  Lc > 0: shared-prefix
    row layout, x fp32 [Lc + B*rep*(L-Lc), D] (B must then be given)."""
    _chk(emb, torch.float32, "emb"); _chk(x, torch.float32, "x")
    D = emb.shape[1]
    if Lc > 0:
        _chk(ids, torch.int32, "ids")
        if not ids.is_contiguous() or ids.shape != (B, Lp) or not x.is_contiguous() or not emb.is_contiguous() \
                or x.numel() != (Lc + B * rep * (L - Lc)) * D:
            raise MtsError("prompt_gather (shared prefix): ids [B, Lp] and x [Lc + B*rep*(L-Lc), D] contiguous required")
        _lib.call("mts_prompt_gather_shared", ids.data_ptr(), emb.data_ptr(), _ptr(wpe), x.data_ptr(), B, rep, Lp, L,
                  Lc, D, _stream())
        return x
    if B is None:
        B = x.shape[0] // rep
    if x.numel() != B * rep * L * D:
        raise MtsError("prompt_gather: x must hold B*rep*L rows of D")
    if Lp > 0:
        _chk(ids, torch.int32, "ids")
        if not ids.is_contiguous() or ids.shape != (B, Lp):
            raise MtsError("ids must be contiguous int32 [B, Lp]")
    if not x.is_contiguous() or not emb.is_contiguous():
        raise MtsError("prompt_gather needs contiguous tensors")
    _lib.call("mts_prompt_gather", _ptr(ids) if Lp > 0 else 0, emb.data_ptr(), _ptr(wpe),
              x.data_ptr(), B, rep, Lp, L, D, _stream())
    return x


def cast_bf16(x, out=None):
    _chk(x, torch.float32, "x")
    x = x.contiguous()
    if out is None:
        out = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
    _lib.call("mts_cast_f32_bf16", x.data_ptr(), out.data_ptr(), x.numel(), _stream())
    return out


def cast_f32(x, out=None):
    _chk(x, torch.bfloat16, "x")
    x = x.contiguous()
    if out is None:
        out = torch.empty(x.shape, device=x.device, dtype=torch.float32)
    _lib.call("mts_cast_bf16_f32", x.data_ptr(), out.data_ptr(), x.numel(), _stream())
    return out


def transpose_to_bf16(x):
    """[rows, cols] fp32 or bf16 -> [cols, rows] bf16."""
    _chk(x, None, "x")
    x = x.contiguous()
    rows, cols = x.shape
    out = torch.empty(cols, rows, device=x.device, dtype=torch.bfloat16)
    fn = "mts_transpose_f32_bf16" if x.dtype == torch.float32 else "mts_transpose_bf16"
    if x.dtype not in (torch.float32, torch.bfloat16):
        raise MtsError("transpose_to_bf16: fp32 or bf16 input")
    _lib.call(fn, x.data_ptr(), out.data_ptr(), rows, cols, _stream())
    return out


def pack_gate_up(gate, up):
    _chk(gate, torch.bfloat16, "gate"); _chk(up, torch.bfloat16, "up")
    gate = gate.contiguous(); up = up.contiguous()
    I, K = gate.shape
    out = torch.empty(((I + 127) // 128) * 256, K, device=gate.device, dtype=torch.bfloat16)
    _lib.call("mts_pack_gate_up", gate.data_ptr(), up.data_ptr(), out.data_ptr(), I, K, _stream())
    return out


def swiglu(gu, I):
    _chk(gu, torch.bfloat16, "gu")
    rows = gu.numel() // gu.shape[-1]
    out = torch.empty(*gu.shape[:-1], I, device=gu.device, dtype=torch.bfloat16)
    _lib.call("mts_swiglu", gu.data_ptr(), gu.shape[-1], out.data_ptr(), rows, I, _stream())
    return out


def sigmoid_(y):
    _chk(y, torch.float32, "y")
    if not y.is_contiguous():
        raise MtsError("sigmoid_ needs a contiguous tensor")
    _lib.call("mts_sigmoid", y.data_ptr(), y.numel(), _stream())
    return y


def softmax_lastdim_(y):
    _chk(y, torch.float32, "y")
    if not y.is_contiguous():
        raise MtsError("softmax_lastdim_ needs a contiguous tensor")
    n = y.shape[-1]
    _lib.call("mts_softmax_lastdim", y.data_ptr(), y.numel() // n, n, _stream())
    return y


# ------------------------------------------------------------------------------------------------
# training-path kernels
# ------------------------------------------------------------------------------------------------
def _dt(t):
    return MTS_F32 if t.dtype == torch.float32 else MTS_BF16


def rmsnorm_bwd(x, w, dy, dx, eps, accumulate=True, dx_bf16=None):
    """dx (+)= J_rmsnorm(x)^T (w*dy); x fp32 [rows,D], dy bf16, dx fp32; dx_bf16: optional bf16 copy of the result."""
    _chk(x, torch.float32, "x"); _chk(dy, torch.bfloat16, "dy"); _chk(dx, torch.float32, "dx")
    D = x.shape[-1]
    _lib.call("mts_rmsnorm_bwd", x.data_ptr(), D, w.data_ptr(), dy.data_ptr(), dx.data_ptr(), _ptr(dx_bf16),
              x.numel() // D, D, eps, 1 if accumulate else 0, _stream())
    return dx


def layernorm_bwd(x, w, dy, dx, eps, accumulate=True, dx_bf16=None):
    _chk(x, torch.float32, "x"); _chk(dy, torch.bfloat16, "dy"); _chk(dx, torch.float32, "dx")
    D = x.shape[-1]
    _lib.call("mts_layernorm_bwd", x.data_ptr(), D, w.data_ptr(), dy.data_ptr(), dx.data_ptr(), _ptr(dx_bf16),
              x.numel() // D, D, eps, 1 if accumulate else 0, _stream())
    return dx


def rope_qk_(qkv, Bp, L, H, hd, rope):
    """In-place rotate-half RoPE on the q and k sections of qkv bf16 [Bp*L, 3*H*hd]."""
    _chk(qkv, torch.bfloat16, "qkv")
    cos, sin = rope
    if not qkv.is_contiguous() or cos.shape[0] < L or cos.shape[1] != hd // 2:
        raise MtsError("rope_qk_: contiguous qkv and [>=L, hd/2] tables required")
    _lib.call("mts_rope_qk", qkv.data_ptr(), cos.data_ptr(), sin.data_ptr(), Bp, L, H, hd, _stream())
    return qkv


def norm_wgrad(x, dy, eps, layernorm=True):
    """(dgamma, dbeta) fp32 [D] of a LayerNorm / RMSNorm whose output gradient is dy bf16 [rows, D] (x fp32 [rows, D])."""
    _chk(x, torch.float32, "x"); _chk(dy, torch.bfloat16, "dy")
    rows, D = dy.shape
    if not dy.is_contiguous() or x.shape != (rows, D) or x.stride(1) != 1:
        raise MtsError("norm_wgrad: x fp32 [rows, D] (unit column stride) and contiguous dy bf16 [rows, D] required")
    partial = torch.empty((rows + 31) // 32, 2 * D, device=x.device, dtype=torch.float32)
    _lib.call("mts_norm_wgrad_partial", x.data_ptr(), x.stride(0), dy.data_ptr(), partial.data_ptr(), rows, D, eps,
              1 if layernorm else 0, _stream())
    both = colsum(partial)
    return both[:D], both[D:]


def rope_qk_shared_(qkv, Bp, Lc, Ls, H, hd, rope):
    """rope_qk_ on the shared-prefix row layout (qkv bf16 [Lc + Bp*Ls, 3*H*hd])."""
    _chk(qkv, torch.bfloat16, "qkv")
    cos, sin = rope
    if not qkv.is_contiguous() or cos.shape[0] < Lc + Ls or cos.shape[1] != hd // 2:
        raise MtsError("rope_qk_shared_: contiguous qkv and [>=Lc+Ls, hd/2] tables required")
    _lib.call("mts_rope_qk_shared", qkv.data_ptr(), cos.data_ptr(), sin.data_ptr(), Bp, Lc, Ls, H, hd, _stream())
    return qkv


def attn_causal_bwd(qkv, out, dout, lse, Bp, L, H, hd, *, rope=None, scale=None, pre_roped=False):
    _chk(qkv, torch.bfloat16, "qkv"); _chk(out, torch.bfloat16, "out"); _chk(dout, torch.bfloat16, "dout")
    _chk(lse, torch.float32, "lse")
    if scale is None:
        scale = 1.0 / math.sqrt(hd)
    dqkv = torch.empty_like(qkv)
    delta = torch.empty(Bp, H, L, device=qkv.device, dtype=torch.float32)
    cos, sin = rope if rope is not None else (None, None)
    _lib.call("mts_attn_causal_bwd", qkv.data_ptr(), _ptr(cos), _ptr(sin), out.data_ptr(), dout.data_ptr(),
              lse.data_ptr(), delta.data_ptr(), dqkv.data_ptr(), Bp, L, H, hd, scale, 1 if pre_roped else 0,
              _stream())
    return dqkv


def swiglu_blk(gu, I, blk):
    _chk(gu, torch.bfloat16, "gu")
    rows = gu.numel() // gu.shape[-1]
    out = torch.empty(*gu.shape[:-1], I, device=gu.device, dtype=torch.bfloat16)
    _lib.call("mts_swiglu_blk", gu.data_ptr(), gu.shape[-1], out.data_ptr(), rows, I, blk, _stream())
    return out


def swiglu_bwd(gu, dact, I, blk):
    _chk(gu, torch.bfloat16, "gu"); _chk(dact, torch.bfloat16, "dact")
    rows = gu.numel() // gu.shape[-1]
    dgu = torch.empty_like(gu)
    _lib.call("mts_swiglu_bwd", gu.data_ptr(), gu.shape[-1], dact.data_ptr(), dgu.data_ptr(), rows, I, blk,
              _stream())
    return dgu


def gelu_new(pre, dact=None):
    _chk(pre, torch.bfloat16, "pre")
    out = torch.empty_like(pre)
    _lib.call("mts_gelu_new", pre.data_ptr(), _ptr(dact), out.data_ptr(), pre.numel(), _stream())
    return out


def softmax_bwd_rows(p, dp, scale):
    _chk(p, torch.bfloat16, "p"); _chk(dp, torch.float32, "dp")
    n = p.shape[-1]
    ds = torch.empty_like(p)
    _lib.call("mts_softmax_bwd_rows", p.data_ptr(), dp.data_ptr(), ds.data_ptr(), p.numel() // n, n, scale,
              _stream())
    return ds


def colsum(x, rows=None, cols=None, ld=None, out=None):
    """fp32 [cols] = column sums of a (possibly strided) 2-D view of x."""
    _chk(x, None, "x")
    if rows is None:
        cols = x.shape[-1]; rows = x.numel() // cols
    ld = cols if ld is None else ld
    if out is None:
        out = torch.empty(cols, device=x.device, dtype=torch.float32)
    elif out.numel() != cols or out.dtype != torch.float32 or not out.is_contiguous():
        raise MtsError("colsum: out must be contiguous fp32 [cols]")
    _lib.call("mts_colsum", x.data_ptr(), _dt(x), ld, out.data_ptr(), rows, cols, _stream())
    return out


def ceil8(n):
    return (n + 7) // 8 * 8


def transpose_strided(x, *, rows, cols, ld_in=None, batch=1, in_bs=0, in_off=0, out=None, ld_out=None):
    """out[c, b*rows + r] = bf16(x[b][r][c]); out is [cols, ld_out] with ld_out = ceil8(batch*rows),
    zero padded (so it can be a K-major GEMM operand with K = batch*rows)."""
    _chk(x, None, "x")
    ld_in = cols if ld_in is None else ld_in
    k = batch * rows
    if ld_out is None:
        ld_out = ceil8(k)
    if out is None:
        out = (torch.zeros if ld_out != k else torch.empty)(cols, ld_out, device=x.device, dtype=torch.bfloat16)
    _lib.call("mts_transpose_strided", x.data_ptr() + x.element_size() * in_off, _dt(x), ld_in, in_bs,
              out.data_ptr(), ld_out, batch, rows, cols, _stream())
    return out


def cast_rows(x, *, rows, cols, ld_in=None, batch=1, in_bs=0, in_off=0, ld_out=None, out=None, out_bs=0,
              out_off=0):
    """out[(b*rows + r), :cols] = bf16(x[b][r][:cols]); out [batch*rows, ld_out] (zero padded columns)."""
    _chk(x, torch.float32, "x")
    ld_in = cols if ld_in is None else ld_in
    if ld_out is None:
        ld_out = ceil8(cols)
    if out is None:
        out = (torch.zeros if ld_out != cols else torch.empty)(batch * rows, ld_out, device=x.device,
                                                               dtype=torch.bfloat16)
    _lib.call("mts_cast_rows_f32_bf16", x.data_ptr() + 4 * in_off, ld_in, in_bs, out.data_ptr() + 2 * out_off, ld_out,
              out_bs, batch, rows, cols, _stream())
    return out


def revin_denorm_bwd(dy, std):
    _chk(dy, torch.float32, "dy")
    dy = dy.contiguous()
    B, T, Cc = dy.shape
    out = torch.empty_like(dy)
    _lib.call("mts_revin_denorm_bwd", dy.data_ptr(), std.data_ptr(), out.data_ptr(), B, T, Cc, _stream())
    return out


def rowsum(x):
    """fp32 [rows] = row sums of a contiguous fp32 [rows, cols]."""
    _chk(x, torch.float32, "x")
    rows, cols = x.shape
    out = torch.empty(rows, device=x.device, dtype=torch.float32)
    _lib.call("mts_rowsum_f32", x.data_ptr(), cols, out.data_ptr(), rows, cols, _stream())
    return out


# ------------------------------------------------------------------------------------------------
# covariate merges over the feature axis
# ------------------------------------------------------------------------------------------------
def group_reduce(x, B, C, R, *, w=None, bias=None, out=None, out_bs=None, out_off=0, accumulate=False):
    """out[b, r] = sum_c w[c] * x[b, c, r] (+ bias); w None = mean.  `out` may be a strided destination."""
    _chk(x, torch.float32, "x")
    if out is None:
        out = torch.empty(B, R, device=x.device, dtype=torch.float32)
    out_bs = R if out_bs is None else out_bs
    _lib.call("mts_group_reduce", x.data_ptr(), _ptr(w), _ptr(bias), out.data_ptr() + 4 * out_off, out_bs, B, C, R,
              1 if accumulate else 0, _stream())
    return out


def group_reduce_bwd(dout, B, C, R, *, w=None, x=None, dout_bs=None, dout_off=0):
    """Returns (din [B*C, R], dw [C] or None, dbias [1] or None)."""
    _chk(dout, torch.float32, "dout")
    din = torch.empty(B * C, R, device=dout.device, dtype=torch.float32)
    dw = dbias = None
    if x is not None:
        dw = torch.empty(C, device=dout.device, dtype=torch.float32)
        dbias = torch.empty(1, device=dout.device, dtype=torch.float32)
    _lib.call("mts_group_reduce_bwd", dout.data_ptr() + 4 * dout_off, R if dout_bs is None else dout_bs, _ptr(w),
              _ptr(x), din.data_ptr(), _ptr(dw), _ptr(dbias), B, C, R, _stream())
    return din, dw, dbias


def merge_end(h, W, bias, B, C, P, O):
    _chk(h, torch.float32, "h")
    y = torch.empty(B, P, O, device=h.device, dtype=torch.float32)
    _lib.call("mts_merge_end", h.data_ptr(), W.data_ptr(), bias.data_ptr(), y.data_ptr(), B, C, P, O, _stream())
    return y


def merge_end_bwd(dy, h, W, B, C, P, O):
    _chk(dy, torch.float32, "dy")
    dh = torch.empty_like(h)
    dW = torch.empty_like(W)
    db = torch.empty(O, device=dy.device, dtype=torch.float32)
    _lib.call("mts_merge_end_bwd", dy.data_ptr(), h.data_ptr(), W.data_ptr(), dh.data_ptr(), dW.data_ptr(),
              db.data_ptr(), B, C, P, O, _stream())
    return dh, dW, db


def dropout(x, p, seed, out=None):
    """y = keep(seed, i) ? x / (1-p) : 0 (bf16 or fp32).  Same (p, seed) on the gradient = the backward."""
    _chk(x, None, "x")
    if x.dtype not in (torch.bfloat16, torch.float32) or not x.is_contiguous():
        raise MtsError("dropout: contiguous bf16 / fp32 tensor required")
    if out is None:
        out = torch.empty_like(x)
    _lib.call("mts_dropout", x.data_ptr(), out.data_ptr(), _dt(x), x.numel(), float(p), int(seed) & (2 ** 64 - 1),
              _stream())
    return out


# ------------------------------------------------------------------------------------------------
# evaluation parity modes (fp32 activations, kind::tf32 contractions)
# ------------------------------------------------------------------------------------------------
def round_tf32(x, out=None):
    """Nearest TF32-representable fp32 values (in place when out is x)."""
    _chk(x, torch.float32, "x")
    if not x.is_contiguous():
        raise MtsError("round_tf32 needs a contiguous tensor")
    if out is None:
        out = torch.empty_like(x)
    _lib.call("mts_round_tf32", x.data_ptr(), out.data_ptr(), x.numel(), _stream())
    return out


def split_tf32(x):
    """(hi, lo) with x ~ hi + lo, both TF32-representable: operands of the 3xTF32 GEMM."""
    _chk(x, torch.float32, "x")
    if not x.is_contiguous():
        raise MtsError("split_tf32 needs a contiguous tensor")
    hi, lo = torch.empty_like(x), torch.empty_like(x)
    _lib.call("mts_split_tf32", x.data_ptr(), hi.data_ptr(), lo.data_ptr(), x.numel(), _stream())
    return hi, lo


def softmax_rows_f32(s, scale, out=None):
    _chk(s, torch.float32, "s")
    if not s.is_contiguous():
        raise MtsError("softmax_rows_f32 needs contiguous input")
    n = s.shape[-1]
    if out is None:
        out = torch.empty_like(s)
    _lib.call("mts_softmax_rows_f32", s.data_ptr(), out.data_ptr(), s.numel() // n, n, scale, _stream())
    return out


def attn_causal_f32(qkv, Bp, Lc, Ls, H, hd, *, scale=None, out=None, round_out=False, tf32=False):
    """fp32 causal attention: qkv fp32 [Lc + Bp*Ls, 3*H*hd] (q / k rotated) -> out fp32 [Lc + Bp*Ls, H*hd].
    tf32=True: both contractions on the tensor cores in TF32 (the "tf32" mode); False: plain fp32 FMA arithmetic."""
    _chk(qkv, torch.float32, "qkv")
    M = Lc + Bp * Ls
    if not qkv.is_contiguous() or qkv.shape != (M, 3 * H * hd):
        raise MtsError("attn_causal_f32 needs contiguous qkv [Lc + Bp*Ls, 3*H*hd]")
    if scale is None:
        scale = 1.0 / math.sqrt(hd)
    if out is None:
        out = torch.empty(M, H * hd, device=qkv.device, dtype=torch.float32)
    _lib.call("mts_attn_causal_tf32" if tf32 else "mts_attn_causal_f32", qkv.data_ptr(), out.data_ptr(), Bp, Lc, Ls, H, hd,
              scale, 1 if round_out else 0, _stream())
    return out


# ------------------------------------------------------------------------------------------------
# prompt statistics
# ------------------------------------------------------------------------------------------------
def input_stats(x, *, f0=0, n_features=None, n_lags=5):
    """Device side of build_input_stats_prompt (models/medtsllm.py:441-495): x fp32 [B, T, C] ->
    (stats fp32 [B, C_sel, 4] = min / max / median / trend, lags int32 [B, n_lags]), both still on the device."""
    _chk(x, torch.float32, "x")
    x = x.contiguous()
    B, T, Cc = x.shape
    C_sel = Cc - f0 if n_features is None else n_features
    stats = torch.empty(B, C_sel, 4, device=x.device, dtype=torch.float32)
    corr = torch.empty(B, C_sel, T, device=x.device, dtype=torch.float64)
    lags = torch.empty(B, n_lags, device=x.device, dtype=torch.int32)
    _lib.call("mts_input_stats", x.data_ptr(), stats.data_ptr(), corr.data_ptr(), lags.data_ptr(), B, T, Cc, f0, C_sel,
              n_lags, _stream())
    return stats, lags
