"""Per-kernel numerics: each libmtsb200 entry point (called through the C ABI) against a plain
PyTorch fp32 reference of the same op on identical inputs.  Tolerances are stated per test.

Integer/index work (patch gather) is compared bit-exactly.
"""
import math
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops(cuda):
    from medtsllm_b200 import ops as _ops
    return _ops


def _rel_l2(a, b):
    a = a.float(); b = b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _bf16_gemm_ref(a, b):
    # fp32 accumulation over bf16-rounded operands (the kernel's arithmetic contract)
    return a.float() @ b.float().t()


# --------------------------------------------------------------------------------------------- GEMM
@pytest.mark.parametrize("m,n,k,bn", [
    (128, 256, 64, 256), (128, 128, 128, 128), (128, 64, 64, 64),
    (256, 512, 512, 0), (300, 264, 200, 0), (1, 8, 8, 64), (77, 1024, 96, 0),
    (2048, 4096, 512, 0), (1024, 768, 4096, 128), (4096, 4096, 1024, 256),
])
def test_gemm_store_bf16(ops, cuda, m, n, k, bn):
    g = torch.Generator(device="cpu").manual_seed(m * 7 + n * 3 + k)
    a = torch.randn(m, k, generator=g).to(cuda, torch.bfloat16)
    b = torch.randn(n, k, generator=g).to(cuda, torch.bfloat16)
    d = torch.full((m, n), float("nan"), device=cuda, dtype=torch.bfloat16)
    ops.gemm(a, b, d, m=m, n=n, k=k, block_n=bn)
    ref = _bf16_gemm_ref(a, b)
    # bf16 output rounding: 2^-9 relative per element; accumulation-order noise ~1e-6
    torch.testing.assert_close(d.float(), ref, rtol=8e-3, atol=2e-2 * math.sqrt(k) / 8)
    assert _rel_l2(d, ref) < 3e-3


def test_gemm_store_f32_bias_alpha(ops, cuda):
    m, n, k = 384, 320, 328
    g = torch.Generator().manual_seed(1)
    a = torch.randn(m, k, generator=g).to(cuda, torch.bfloat16)
    b = torch.randn(n, k, generator=g).to(cuda, torch.bfloat16)
    bias = torch.randn(n, generator=g).to(cuda)
    d = torch.empty(m, n, device=cuda, dtype=torch.float32)
    ops.gemm(a, b, d, m=m, n=n, k=k, bias=bias, bias_axis=1, alpha=0.5)
    ref = 0.5 * _bf16_gemm_ref(a, b) + bias
    torch.testing.assert_close(d, ref, rtol=1e-4, atol=1e-3)   # fp32 out: only summation order differs
    bias_m = torch.randn(m, generator=g).to(cuda)
    ops.gemm(a, b, d, m=m, n=n, k=k, bias=bias_m, bias_axis=2)
    torch.testing.assert_close(d, _bf16_gemm_ref(a, b) + bias_m[:, None], rtol=1e-4, atol=1e-3)


def test_gemm_resid_add(ops, cuda):
    m, n, k = 640, 512, 256
    g = torch.Generator().manual_seed(2)
    a = torch.randn(m, k, generator=g).to(cuda, torch.bfloat16)
    b = torch.randn(n, k, generator=g).to(cuda, torch.bfloat16)
    bias = torch.randn(n, generator=g).to(cuda)
    r0 = torch.randn(m, n, generator=g).to(cuda)
    r = r0.clone()
    ops.gemm(a, b, r, m=m, n=n, k=k, bias=bias, bias_axis=1, epilogue=1)
    torch.testing.assert_close(r, r0 + _bf16_gemm_ref(a, b) + bias, rtol=1e-4, atol=1e-3)


def test_gemm_gelu_new(ops, cuda):
    m, n, k = 256, 384, 192
    g = torch.Generator().manual_seed(3)
    a = (torch.randn(m, k, generator=g) * 0.3).to(cuda, torch.bfloat16)
    b = (torch.randn(n, k, generator=g) * 0.3).to(cuda, torch.bfloat16)
    bias = torch.randn(n, generator=g).to(cuda)
    d = torch.empty(m, n, device=cuda, dtype=torch.bfloat16)
    ops.gemm(a, b, d, m=m, n=n, k=k, bias=bias, bias_axis=1, epilogue=2)
    x = _bf16_gemm_ref(a, b) + bias
    ref = 0.5 * x * (1 + torch.tanh(math.sqrt(2 / math.pi) * (x + 0.044715 * x ** 3)))
    torch.testing.assert_close(d.float(), ref, rtol=8e-3, atol=4e-3)


def test_gemm_swiglu_packed(ops, cuda):
    m, I, k = 200, 320, 256          # I not a multiple of 128: exercises the zero-padded block
    g = torch.Generator().manual_seed(4)
    x = (torch.randn(m, k, generator=g) * 0.5).to(cuda, torch.bfloat16)
    wg = (torch.randn(I, k, generator=g) * 0.1).to(cuda, torch.bfloat16)
    wu = (torch.randn(I, k, generator=g) * 0.1).to(cuda, torch.bfloat16)
    packed = ops.pack_gate_up(wg, wu)
    n = packed.shape[0]
    assert n == 3 * 256
    d = torch.full((m, n // 2), float("nan"), device=cuda, dtype=torch.bfloat16)
    ops.gemm(x, packed, d, m=m, n=n, k=k, epilogue=3)
    gate = _bf16_gemm_ref(x, wg); up = _bf16_gemm_ref(x, wu)
    ref = torch.nn.functional.silu(gate) * up
    torch.testing.assert_close(d[:, :I].float(), ref, rtol=8e-3, atol=4e-3)
    assert torch.all(d[:, I:] == 0)


def test_gemm_batched_strided_transposed(ops, cuda):
    # per-head scores: A = Q[:, h] (row stride H*dk), B = K[:, h], D [H, M, S] fp32
    M, S, H, dk = 192, 1024, 8, 64
    g = torch.Generator().manual_seed(5)
    q = torch.randn(M, H * dk, generator=g).to(cuda, torch.bfloat16)
    kk = torch.randn(S, H * dk, generator=g).to(cuda, torch.bfloat16)
    d = torch.empty(H, M, S, device=cuda, dtype=torch.float32)
    ops.gemm(q, kk, d, m=M, n=S, k=dk, batch=H, lda=H * dk, ldb=H * dk, a_bs=dk, b_bs=dk, d_bs=M * S)
    ref = torch.einsum("mhe,she->hms", q.float().view(M, H, dk), kk.float().view(S, H, dk))
    torch.testing.assert_close(d, ref, rtol=1e-4, atol=1e-3)
    # transposed store with a shared (batch-stride 0) B operand
    a = torch.randn(3, 64, 128, generator=g).to(cuda, torch.bfloat16)
    w = torch.randn(72, 128, generator=g).to(cuda, torch.bfloat16)
    dt = torch.empty(3, 72, 64, device=cuda, dtype=torch.bfloat16)
    ops.gemm(a, w, dt, m=64, n=72, k=128, batch=3, a_bs=64 * 128, b_bs=0, d_bs=72 * 64,
             d_transposed=True, ldd=64)
    ref_t = torch.einsum("bmk,nk->bnm", a.float(), w.float())
    torch.testing.assert_close(dt.float(), ref_t, rtol=8e-3, atol=6e-2)


def test_gemm_rejects_bad_args(ops, cuda):
    from medtsllm_b200 import MtsError
    a = torch.zeros(8, 12, device=cuda, dtype=torch.bfloat16)
    with pytest.raises(MtsError):
        ops.gemm(a, a, torch.zeros(8, 8, device=cuda, dtype=torch.bfloat16), m=8, n=8, k=12)  # lda % 8
    with pytest.raises(MtsError):
        ops.gemm(a.cpu(), a, a, m=8, n=8, k=8)


# ---------------------------------------------------------------------------------------- front end
@pytest.mark.parametrize("B,T,C", [(8, 96, 7), (64, 100, 25), (16, 336, 2), (32, 512, 3), (16, 1024, 12), (3, 17, 1)])
def test_patch_gather_bit_exact(ops, cuda, B, T, C):
    P, S = 16, 8
    g = torch.Generator().manual_seed(T)
    x = torch.randn(B, T, C, generator=g).to(cuda)
    got = ops.patch_gather(x, P, S)
    xp = x.permute(0, 2, 1)
    xp = torch.cat([xp, xp[:, :, -1:].repeat(1, 1, S)], dim=-1)          # ReplicationPad1d((0,S))
    ref = xp.unfold(-1, P, S).reshape(B * C, -1, P)
    assert got.shape == ref.shape
    assert torch.equal(got, ref)                                          # index work: bit exact


@pytest.mark.parametrize("B,T,C,concat", [(8, 96, 7, True), (64, 100, 25, True), (32, 512, 3, True),
                                          (16, 1024, 12, False), (4, 336, 2, False)])
def test_revin_patch_embed(ops, cuda, B, T, C, concat):
    P, S, dm = 16, 8, 32
    g = torch.Generator().manual_seed(B + T)
    x = (torch.randn(B, T, C, generator=g) * (torch.rand(C, generator=g) * 4.5 + 0.5)
         + (torch.rand(C, generator=g) * 20 - 10)).to(cuda)
    w = torch.randn(dm, P, 3, generator=g).to(cuda) * 0.2
    ob, of, mean, std = ops.revin_patch_embed(x, w, P, S, concat=concat, want_f32=True)
    mu = x.mean(1, keepdim=True)
    sd = torch.sqrt(x.var(1, keepdim=True, unbiased=False) + 1e-5)
    xn = ((x - mu) / sd).permute(0, 2, 1)
    xp = torch.cat([xn, xn[:, :, -1:].repeat(1, 1, S)], dim=-1).unfold(-1, P, S).reshape(B * C, -1, P)
    N = xp.shape[1]
    ref = torch.nn.functional.conv1d(
        torch.nn.functional.pad(xp.permute(0, 2, 1), (1, 1), mode="circular"), w).transpose(1, 2)
    if concat:
        ref = ref.reshape(B, C, N, dm).permute(0, 2, 1, 3).reshape(B, N, C * dm)
    torch.testing.assert_close(mean, mu.squeeze(1), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(std, sd.squeeze(1), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(of, ref, rtol=1e-4, atol=2e-5)             # fp32: summation order only
    torch.testing.assert_close(ob.float(), ref, rtol=8e-3, atol=1e-3)     # + bf16 rounding


def test_revin_patch_embed_bwd_and_denorm(ops, cuda):
    B, T, C, P, S, dm = 6, 100, 5, 16, 8, 32
    g = torch.Generator().manual_seed(9)
    x = torch.randn(B, T, C, generator=g).to(cuda) * 3 + 1
    w = (torch.randn(dm, P, 3, generator=g) * 0.2).to(cuda).requires_grad_(True)
    _, of, mean, std = ops.revin_patch_embed(x, w.detach(), P, S, concat=True, want_f32=True, want_bf16=False)
    dout = torch.randn(of.shape, generator=g).to(cuda)
    dw = ops.revin_patch_embed_bwd(x, mean, std, dout, P, S, dm, concat=True)
    xn = ((x - mean[:, None]) / std[:, None]).permute(0, 2, 1)
    xp = torch.cat([xn, xn[:, :, -1:].repeat(1, 1, S)], dim=-1).unfold(-1, P, S).reshape(B * C, -1, P)
    N = xp.shape[1]
    ref = torch.nn.functional.conv1d(
        torch.nn.functional.pad(xp.permute(0, 2, 1), (1, 1), mode="circular"), w).transpose(1, 2)
    ref = ref.reshape(B, C, N, dm).permute(0, 2, 1, 3).reshape(B, N, C * dm)
    (ref * dout).sum().backward()
    torch.testing.assert_close(dw, w.grad, rtol=1e-4, atol=1e-3)
    y = torch.randn(B, 7, C, generator=g).to(cuda)
    y0 = y.clone()
    ops.revin_denorm(y, mean, std)
    torch.testing.assert_close(y, y0 * std[:, None] + mean[:, None], rtol=1e-6, atol=1e-6)


# ---------------------------------------------------------------------------------------- row ops
@pytest.mark.parametrize("rows,D", [(37, 4096), (512, 1024), (5, 768), (3, 64)])
def test_rmsnorm_layernorm(ops, cuda, rows, D):
    g = torch.Generator().manual_seed(D)
    x = torch.randn(rows, D, generator=g).to(cuda) * 2 + 0.3
    w = torch.randn(D, generator=g).to(cuda)
    b = torch.randn(D, generator=g).to(cuda)
    ref = x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + 1e-5) * w
    torch.testing.assert_close(ops.rmsnorm(x, w, 1e-5, out_dtype=torch.float32), ref, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(ops.rmsnorm(x, w, 1e-5).float(), ref, rtol=8e-3, atol=1e-3)
    ref_ln = torch.nn.functional.layer_norm(x, (D,), w, b, 1e-5)
    torch.testing.assert_close(ops.layernorm(x, w, b, 1e-5, out_dtype=torch.float32), ref_ln, rtol=1e-5, atol=2e-5)
    torch.testing.assert_close(ops.layernorm(x, w, b, 1e-5).float(), ref_ln, rtol=8e-3, atol=2e-3)


def test_softmax_rows(ops, cuda):
    s = torch.randn(8 * 77, 1024, device=cuda) * 5
    p = ops.softmax_rows(s, 0.125)
    ref = torch.softmax(s * 0.125, -1)
    torch.testing.assert_close(p.float(), ref, rtol=8e-3, atol=1e-6)
    torch.testing.assert_close(p.float().sum(-1), torch.ones(s.shape[0], device=cuda), rtol=0, atol=5e-3)


def test_casts_transpose_swiglu_gather(ops, cuda):
    g = torch.Generator().manual_seed(11)
    x = torch.randn(1000, 37, generator=g).to(cuda)
    assert torch.equal(ops.cast_bf16(x), x.to(torch.bfloat16))
    assert torch.equal(ops.cast_f32(x.to(torch.bfloat16)), x.to(torch.bfloat16).float())
    assert torch.equal(ops.transpose_to_bf16(x), x.t().contiguous().to(torch.bfloat16))
    xb = x.to(torch.bfloat16)
    assert torch.equal(ops.transpose_to_bf16(xb), xb.t().contiguous())
    gu = torch.randn(50, 2 * 64, generator=g).to(cuda, torch.bfloat16)
    ref = torch.nn.functional.silu(gu[:, :64].float()) * gu[:, 64:].float()
    torch.testing.assert_close(ops.swiglu(gu, 64).float(), ref, rtol=8e-3, atol=1e-3)
    # prompt gather with repeat and position embedding
    V, D, B, Lp, L, rep = 50, 64, 3, 5, 9, 2
    emb = torch.randn(V, D, generator=g).to(cuda)
    wpe = torch.randn(L, D, generator=g).to(cuda)
    ids = torch.randint(0, V, (B, Lp), generator=g).to(cuda, torch.int32)
    xo = torch.full((B * rep, L, D), float("nan"), device=cuda)
    ops.prompt_gather(ids, emb, wpe, xo, rep=rep, Lp=Lp, L=L)
    ref = torch.zeros(B, L, D, device=cuda)
    ref[:, :Lp] = emb[ids.long()]
    ref = (ref + wpe[None]).repeat_interleave(rep, 0)
    assert torch.equal(xo, ref)
    xo2 = torch.full((B, L, D), float("nan"), device=cuda)
    ops.prompt_gather(ids, emb, None, xo2, rep=1, Lp=Lp, L=L)
    assert torch.equal(xo2[:, :Lp], emb[ids.long()]) and torch.all(xo2[:, Lp:] == 0)


# ---------------------------------------------------------------------------------------- attention
def _rope_tables(L, hd, device, theta=10000.0):
    inv = 1.0 / (theta ** (torch.arange(0, hd, 2, dtype=torch.float32) / hd))
    f = torch.arange(L, dtype=torch.float32)[:, None] * inv[None]
    return f.cos().to(device).contiguous(), f.sin().to(device).contiguous()


@pytest.mark.parametrize("Bp,L,H,hd,rope", [(2, 192, 4, 128, True), (3, 140, 4, 64, False),
                                            (1, 64, 2, 128, True), (2, 257, 2, 64, False), (1, 7, 1, 64, True)])
def test_attn_causal(ops, cuda, Bp, L, H, hd, rope):
    g = torch.Generator().manual_seed(L + hd)
    D = H * hd
    qkv = torch.randn(Bp * L, 3 * D, generator=g).to(cuda, torch.bfloat16)
    tabs = _rope_tables(L, hd, cuda) if rope else None
    out, lse = ops.attn_causal(qkv, Bp, L, H, hd, rope=tabs, want_lse=True)
    q, k, v = (t.view(Bp, L, H, hd).transpose(1, 2) for t in qkv.float().split(D, dim=-1))
    if rope:
        cos = torch.cat([tabs[0], tabs[0]], -1)[None, None]
        sin = torch.cat([tabs[1], tabs[1]], -1)[None, None]
        rot = lambda t: torch.cat([-t[..., hd // 2:], t[..., :hd // 2]], -1)
        q = q * cos + rot(q) * sin
        k = k * cos + rot(k) * sin
    s = (q @ k.transpose(-1, -2)) / math.sqrt(hd)
    mask = torch.ones(L, L, device=cuda, dtype=torch.bool).tril()
    s = s.masked_fill(~mask, float("-inf"))
    ref = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(Bp * L, D)
    # bf16 rounding of rotated q/k, P and the output: a few 1e-3 relative
    assert _rel_l2(out, ref) < 8e-3
    torch.testing.assert_close(out.float(), ref, rtol=2e-2, atol=2e-2)
    torch.testing.assert_close(lse, torch.logsumexp(s, -1), rtol=1e-2, atol=2e-2)


@pytest.mark.parametrize("Bp,Lc,Ls,H,hd", [
    (5, 128, 64, 2, 128),      # BIDMC: two samples share a tile
    (11, 128, 12, 3, 64),      # PSM: eight samples per tile, ragged last group
    (7, 128, 42, 2, 128),      # Ventilator: own rows not a multiple of 16
    (3, 128, 128, 1, 128),     # LUDB: 256 key columns
    (4, 37, 12, 3, 64),        # prefix not a multiple of 16: dummy columns
    (5, 130, 33, 1, 128),      # prefix in two tiles
    (2, 32, 150, 2, 64),       # own rows in two tiles
    (3, 0, 192, 2, 128),       # plain layout, two tiles per sample
    (6, 0, 49, 2, 64),         # plain layout, two samples per tile
    (2, 0, 256, 1, 128),
    (2, 0, 5, 1, 64),          # tiny: sixteen-column segments mostly padding
    (3, 16, 1, 2, 128),        # one own token per sample
    (1, 0, 129, 1, 64),        # one row in the second tile
    (40, 48, 3, 1, 64),        # samples-per-tile limited by the 256 key columns
])
def test_attn_tensor_memory_kernel(ops, cuda, Bp, Lc, Ls, H, hd):
    """The tcgen05 / TMEM forward (attention_tc.cu) against fp64 attention and against the mma.sync kernels it replaces,
    on every tiling case: several samples per 128-row tile, tiled prefix, tiled own rows, dummy columns."""
    from medtsllm_b200 import _lib
    g = torch.Generator().manual_seed(Lc * 5 + Ls + hd)
    D, L = H * hd, Lc + Ls
    M = Lc + Bp * Ls
    qkv = (torch.randn(M, 3 * D, generator=g) * 0.8).to(cuda, torch.bfloat16)

    def run():
        if Lc:
            out, lse_full = ops.attn_causal_shared(qkv, Bp, Lc, Ls, H, hd, want_lse=True)
            return out, lse_full
        return ops.attn_causal(qkv, Bp, L, H, hd, rope=None, want_lse=True)

    _lib.set_option("attn_tc", 2)             # 2 = whenever the shape fits (auto leaves small head-dim-64 problems alone)
    try:
        out, lse = run()
        out2, lse2 = run()                                                  # deterministic
        assert torch.equal(out, out2) and torch.equal(lse, lse2)
        if Lc:   # bit-identical to the per-sample layout (prefix repeated in front of every sample) on the same kernel
            out_p, lse_p = ops.attn_causal(_expand_shared(qkv, Bp, Lc, Ls), Bp, L, H, hd, rope=None, want_lse=True)
            assert torch.equal(_expand_shared(out, Bp, Lc, Ls), out_p)
            assert torch.equal(ops.lse_own_view(lse, Bp, Lc, Ls, H), lse_p[:, :, Lc:])
        _lib.set_option("attn_tc", 0)
        out_old, lse_old = run()
    finally:
        _lib.set_option("attn_tc", 1)
    # fp64 reference on the per-sample view
    full = _expand_shared(qkv, Bp, Lc, Ls) if Lc else qkv
    q, k, v = (t.view(Bp, L, H, hd).transpose(1, 2) for t in full.double().split(D, dim=-1))
    sc = (q @ k.transpose(-1, -2)) / math.sqrt(hd)
    mask = torch.ones(L, L, device=cuda, dtype=torch.bool).tril()
    sc = sc.masked_fill(~mask, float("-inf"))
    ref = (torch.softmax(sc, -1) @ v).transpose(1, 2).reshape(Bp, L, D)
    ref_lse = torch.logsumexp(sc, -1)                                   # [Bp, H, L]
    got = (_expand_shared(out, Bp, Lc, Ls) if Lc else out).view(Bp, L, D)
    assert torch.isfinite(out.float()).all()
    assert _rel_l2(got, ref) < 6e-3                                     # bf16 P and output
    assert _rel_l2(out, out_old) < 6e-3
    if Lc:
        own = ops.lse_own_view(lse, Bp, Lc, Ls, H)
        torch.testing.assert_close(own.double(), ref_lse[:, :, Lc:], rtol=1e-4, atol=2e-4)
        torch.testing.assert_close(lse[:H * Lc].view(H, Lc).double(), ref_lse[0, :, :Lc], rtol=1e-4, atol=2e-4)
        torch.testing.assert_close(lse, lse_old, rtol=1e-4, atol=2e-4)
    else:
        torch.testing.assert_close(lse.double(), ref_lse, rtol=1e-4, atol=2e-4)


@pytest.mark.parametrize("rope", [False, True])
@pytest.mark.parametrize("Bp,Lc,Ls,H,hd", [
    (5, 128, 64, 2, 128),      # BIDMC: two samples per tile, ragged last group
    (11, 128, 12, 3, 64),      # PSM: eight samples per tile
    (7, 128, 42, 2, 128),      # Ventilator: own rows not a multiple of 16
    (3, 128, 128, 1, 128),     # LUDB: 256 key columns
    (4, 37, 12, 3, 64),        # prefix not a multiple of 64: masked gap columns
    (5, 100, 33, 1, 128),
    (3, 16, 1, 2, 128),        # one own token per sample
    (9, 64, 128, 1, 64),       # a full tile of own rows behind a short prefix
])
def test_attn_tensor_memory_backward(ops, cuda, Bp, Lc, Ls, H, hd, rope):
    """The tcgen05 / TMEM backward of the shared-prefix layout (own rows only: frozen backbone) against fp64 autograd and
    against the mma.sync kernels it replaces."""
    from medtsllm_b200 import _lib
    g = torch.Generator().manual_seed(Lc * 3 + Ls + hd)
    D, L = H * hd, Lc + Ls
    M = Lc + Bp * Ls
    qkv = (torch.randn(M, 3 * D, generator=g) * 0.7).to(cuda, torch.bfloat16)       # q / k are taken as already rotated
    dout = torch.randn(Bp * Ls, D, generator=g).to(cuda, torch.bfloat16)
    tabs = _rope_tables(L, hd, cuda) if rope else None
    out, lse_full = ops.attn_causal_shared(qkv, Bp, Lc, Ls, H, hd, want_lse=True)
    lse = ops.lse_own_view(lse_full, Bp, Lc, Ls, H)

    def run():
        return ops.attn_causal_shared_bwd(qkv, out[Lc:], dout, lse, Bp, Lc, Ls, H, hd, rope=tabs)

    _lib.set_option("attn_tc", 2)
    try:
        got = run()
        assert torch.equal(run(), got)                                              # deterministic
        _lib.set_option("attn_tc", 0)
        old = run()
    finally:
        _lib.set_option("attn_tc", 1)
    # fp64 reference on the per-sample view; the gradient w.r.t. the ROTATED q / k is rotated back like the kernels do
    x = _expand_shared(qkv, Bp, Lc, Ls).double().requires_grad_(True)
    q, k, v = (t.view(Bp, L, H, hd).transpose(1, 2) for t in x.split(D, dim=-1))
    sc = (q @ k.transpose(-1, -2)) / math.sqrt(hd)
    mask = torch.ones(L, L, device=cuda, dtype=torch.bool).tril()
    o = (torch.softmax(sc.masked_fill(~mask, float("-inf")), -1) @ v).transpose(1, 2).reshape(Bp, L, D)
    dfull = torch.zeros(Bp, L, D, device=cuda, dtype=torch.float64)
    dfull[:, Lc:] = dout.double().view(Bp, Ls, D)
    (gx,) = torch.autograd.grad(o, x, dfull)
    ref = gx.view(Bp, L, 3 * D)[:, Lc:].reshape(Bp * Ls, 3 * D)
    if rope:
        cos = torch.cat([tabs[0], tabs[0]], -1).double()[Lc:]                       # positions of the own tokens
        sin = torch.cat([tabs[1], tabs[1]], -1).double()[Lc:]
        rot_t = lambda t: torch.cat([t[..., hd // 2:], -t[..., :hd // 2]], -1)      # transpose of rotate-half
        for sl in (slice(0, D), slice(D, 2 * D)):
            gsec = ref[:, sl].view(Bp, Ls, H, hd)
            ref[:, sl] = (gsec * cos[None, :, None] + rot_t(gsec * sin[None, :, None])).reshape(Bp * Ls, D)
    assert torch.isfinite(got.float()).all()
    for name, sl in (("dq", slice(0, D)), ("dk", slice(D, 2 * D)), ("dv", slice(2 * D, 3 * D))):
        e_ref, e_old = _rel_l2(got[:, sl], ref[:, sl]), _rel_l2(got[:, sl], old[:, sl])
        assert e_ref < 1.5e-2, (name, e_ref)        # bf16 P / dS / outputs
        assert e_old < 1.5e-2, (name, e_old)
    # full backward (LoRA: the prefix rows' gradients too): own rows from the tensor-memory kernel, which leaves delta
    # behind for the prefix-key kernels; against the all-mma.sync path
    dout_all = torch.cat([torch.randn(Lc, D, generator=g).to(cuda, torch.bfloat16), dout])
    _lib.set_option("attn_tc", 2)
    try:
        full_tc = ops.attn_causal_shared_bwd_full(qkv, out, dout_all, lse_full, Bp, Lc, Ls, H, hd, rope=tabs)
        _lib.set_option("attn_tc", 0)
        full_old = ops.attn_causal_shared_bwd_full(qkv, out, dout_all, lse_full, Bp, Lc, Ls, H, hd, rope=tabs)
    finally:
        _lib.set_option("attn_tc", 1)
    assert torch.equal(full_tc[Lc:], got)                       # same own-row kernel, same inputs
    assert _rel_l2(full_tc[:Lc], full_old[:Lc]) < 1e-5          # prefix rows: same kernels fed the new delta


def test_full_attention_backward_first_call_in_a_fresh_process(cuda):
    """The full (LoRA) backward as the very first attention call of a process: every kernel it launches must have its
    shared-memory attribute set by that call itself, not by an earlier call of another route (regression: the own rows
    going to the tensor-memory kernel skipped the set-up the prefix-row kernels rely on)."""
    import subprocess
    import sys
    code = (
        "import sys, torch; sys.path.insert(0, %r); from medtsllm_b200 import ops\n"
        "Bp, Lc, Ls, H, hd = 4, 128, 42, 2, 128; D = H * hd; M = Lc + Bp * Ls\n"
        "dev = torch.device('cuda', 0); g = torch.Generator().manual_seed(0)\n"
        "qkv = (torch.randn(M, 3 * D, generator=g) * 0.5).to(dev, torch.bfloat16)\n"
        "out = (torch.randn(M, D, generator=g) * 0.5).to(dev, torch.bfloat16)\n"
        "dout = torch.randn(M, D, generator=g).to(dev, torch.bfloat16)\n"
        "lse = torch.randn(H * Lc + Bp * H * Ls, generator=g).abs().to(dev) + 5\n"
        "d = ops.attn_causal_shared_bwd_full(qkv, out, dout, lse, Bp, Lc, Ls, H, hd)\n"
        "torch.cuda.synchronize(); assert torch.isfinite(d.float()).all(); print('ok')\n"
    ) % str(Path(__file__).resolve().parent.parent / "med-ts-llm_b200")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]


# ------------------------------------------------------------------------------- training-path kernels
def test_gemm_resid_out_of_place_and_scalar_epilogue(ops, cuda):
    g = torch.Generator().manual_seed(21)
    m, n, k = 200, 130, 72                      # n % 4 != 0 -> scalar epilogue path
    a = torch.randn(m, k, generator=g).to(cuda, torch.bfloat16)
    b = torch.randn(n, k, generator=g).to(cuda, torch.bfloat16)
    c = torch.randn(m, n, generator=g).to(cuda)
    d = torch.full((m, n), float("nan"), device=cuda)
    ops.gemm(a, b, d, m=m, n=n, k=k, epilogue=1, c=c)
    torch.testing.assert_close(d, c + _bf16_gemm_ref(a, b), rtol=1e-4, atol=1e-3)
    d16 = torch.full((m, n), float("nan"), device=cuda, dtype=torch.bfloat16)
    ops.gemm(a, b, d16, m=m, n=n, k=k)
    torch.testing.assert_close(d16.float(), _bf16_gemm_ref(a, b), rtol=8e-3, atol=6e-2)
    # odd K with padded leading dimensions (GPT-2 vocabulary 50257)
    k2, ld = 77, 80
    a2 = torch.zeros(m, ld, device=cuda, dtype=torch.bfloat16); a2[:, :k2] = torch.randn(m, k2, generator=g).to(cuda)
    b2 = torch.zeros(64, ld, device=cuda, dtype=torch.bfloat16); b2[:, :k2] = torch.randn(64, k2, generator=g).to(cuda)
    a2[:, k2:] = 7.0; b2[:, k2:] = 7.0          # must be ignored: TMA extent is k, not ld
    d2 = torch.empty(m, 64, device=cuda)
    ops.gemm(a2, b2, d2, m=m, n=64, k=k2, lda=ld, ldb=ld)
    torch.testing.assert_close(d2, a2[:, :k2].float() @ b2[:, :k2].float().t(), rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("rows,D", [(33, 4096), (64, 256)])
def test_norm_bwd(ops, cuda, rows, D):
    g = torch.Generator().manual_seed(D)
    x = (torch.randn(rows, D, generator=g) * 1.5 + 0.2).to(cuda).requires_grad_(True)
    w = torch.randn(D, generator=g).to(cuda)
    b = torch.randn(D, generator=g).to(cuda)
    dy = torch.randn(rows, D, generator=g).to(cuda, torch.bfloat16)
    base = torch.randn(rows, D, generator=g).to(cuda)
    y = x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + 1e-5) * w
    (gx,) = torch.autograd.grad(y, x, dy.float())
    dx = base.clone()
    ops.rmsnorm_bwd(x.detach(), w, dy, dx, 1e-5, accumulate=True)
    torch.testing.assert_close(dx, base + gx, rtol=1e-4, atol=1e-4)
    y = torch.nn.functional.layer_norm(x, (D,), w, b, 1e-5)
    (gx,) = torch.autograd.grad(y, x, dy.float())
    dx = torch.full_like(base, float("nan"))
    ops.layernorm_bwd(x.detach(), w, dy, dx, 1e-5, accumulate=False)
    torch.testing.assert_close(dx, gx, rtol=1e-4, atol=1e-4)


def test_swiglu_gelu_softmax_bwd(ops, cuda):
    g = torch.Generator().manual_seed(31)
    rows, I = 40, 256
    # packed layout, blk = 128
    gate = torch.randn(rows, I, generator=g).to(cuda, torch.bfloat16)
    up = torch.randn(rows, I, generator=g).to(cuda, torch.bfloat16)
    packed = torch.stack([gate.view(rows, 2, 128), up.view(rows, 2, 128)], dim=2).reshape(rows, 2 * I).contiguous()
    dact = torch.randn(rows, I, generator=g).to(cuda, torch.bfloat16)
    gf, uf = gate.float().requires_grad_(True), up.float().requires_grad_(True)
    act = torch.nn.functional.silu(gf) * uf
    torch.testing.assert_close(ops.swiglu_blk(packed, I, 128).float(), act.detach(), rtol=8e-3, atol=2e-3)
    dg, du = torch.autograd.grad(act, (gf, uf), dact.float())
    dgu = ops.swiglu_bwd(packed, dact, I, 128).float().view(rows, 2, 2, 128)
    torch.testing.assert_close(dgu[:, :, 0].reshape(rows, I), dg, rtol=1e-2, atol=4e-3)
    torch.testing.assert_close(dgu[:, :, 1].reshape(rows, I), du, rtol=1e-2, atol=4e-3)
    # gelu_new
    pre = torch.randn(rows, I, generator=g).to(cuda, torch.bfloat16)
    pf = pre.float().requires_grad_(True)
    ref = 0.5 * pf * (1 + torch.tanh(math.sqrt(2 / math.pi) * (pf + 0.044715 * pf ** 3)))
    torch.testing.assert_close(ops.gelu_new(pre).float(), ref.detach(), rtol=8e-3, atol=2e-3)
    (dp,) = torch.autograd.grad(ref, pf, dact.float())
    torch.testing.assert_close(ops.gelu_new(pre, dact).float(), dp, rtol=1e-2, atol=4e-3)
    # softmax backward
    s = (torch.randn(24, 512, generator=g) * 3).to(cuda).requires_grad_(True)
    p = torch.softmax(s * 0.25, -1)
    dpv = torch.randn(24, 512, generator=g).to(cuda)
    (ds_ref,) = torch.autograd.grad(p, s, dpv)
    pb = p.detach().to(torch.bfloat16)
    ds = ops.softmax_bwd_rows(pb, dpv, 0.25)
    pf32 = pb.float()
    ds_exact = 0.25 * pf32 * (dpv - (dpv * pf32).sum(-1, keepdim=True))     # same bf16-rounded P
    torch.testing.assert_close(ds.float(), ds_exact, rtol=8e-3, atol=1e-5)
    assert _rel_l2(ds, ds_ref) < 1e-2


def test_colsum_transpose_cast_rows_denorm_bwd(ops, cuda):
    g = torch.Generator().manual_seed(41)
    x = torch.randn(300, 77, generator=g).to(cuda)
    torch.testing.assert_close(ops.colsum(x), x.sum(0), rtol=1e-5, atol=1e-4)
    xb = x.to(torch.bfloat16)
    torch.testing.assert_close(ops.colsum(xb), xb.float().sum(0), rtol=1e-5, atol=1e-4)
    # batched strided transpose with zero padded K
    src = torch.randn(3, 10, 50, generator=g).to(cuda)           # rows 4..9 of each batch are skipped
    out = ops.transpose_strided(src, batch=3, rows=5, cols=50, ld_in=50, in_bs=500, in_off=50 * 4)
    assert out.shape == (50, 16)
    ref = src[:, 4:9, :].reshape(15, 50).t().to(torch.bfloat16)
    assert torch.equal(out[:, :15], ref) and torch.all(out[:, 15:] == 0)
    # strided cast with padded output rows
    c = ops.cast_rows(src, batch=3, rows=5, cols=50, ld_in=50, in_bs=500, in_off=50 * 4)
    assert c.shape == (15, 56)
    assert torch.equal(c[:, :50], src[:, 4:9, :].reshape(15, 50).to(torch.bfloat16)) and torch.all(c[:, 50:] == 0)
    odd = torch.randn(7, 13, generator=g).to(cuda)
    assert torch.equal(ops.cast_rows(odd, rows=7, cols=13)[:, :13], odd.to(torch.bfloat16))
    dy = torch.randn(4, 9, 3, generator=g).to(cuda)
    std = torch.rand(4, 3, generator=g).to(cuda) + 0.5
    torch.testing.assert_close(ops.revin_denorm_bwd(dy, std), dy * std[:, None], rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("Bp,L,H,hd,rope", [(2, 192, 2, 128, True), (2, 140, 3, 64, False), (1, 70, 2, 64, True)])
def test_attn_causal_bwd(ops, cuda, Bp, L, H, hd, rope):
    g = torch.Generator().manual_seed(L * 3 + hd)
    D = H * hd
    qkv = (torch.randn(Bp * L, 3 * D, generator=g) * 0.7).to(cuda, torch.bfloat16)
    dout = torch.randn(Bp * L, D, generator=g).to(cuda, torch.bfloat16)
    tabs = _rope_tables(L, hd, cuda) if rope else None
    out, lse = ops.attn_causal(qkv, Bp, L, H, hd, rope=tabs, want_lse=True)
    dqkv = ops.attn_causal_bwd(qkv, out, dout, lse, Bp, L, H, hd, rope=tabs)
    x = qkv.float().requires_grad_(True)
    q, k, v = (t.view(Bp, L, H, hd).transpose(1, 2) for t in x.split(D, dim=-1))
    if rope:
        cos = torch.cat([tabs[0], tabs[0]], -1)[None, None]
        sin = torch.cat([tabs[1], tabs[1]], -1)[None, None]
        rot = lambda t: torch.cat([-t[..., hd // 2:], t[..., :hd // 2]], -1)
        q = q * cos + rot(q) * sin
        k = k * cos + rot(k) * sin
    s = (q @ k.transpose(-1, -2)) / math.sqrt(hd)
    mask = torch.ones(L, L, device=cuda, dtype=torch.bool).tril()
    ref = (torch.softmax(s.masked_fill(~mask, float("-inf")), -1) @ v).transpose(1, 2).reshape(Bp * L, D)
    (gx,) = torch.autograd.grad(ref, x, dout.float())
    for name, sl in (("dq", slice(0, D)), ("dk", slice(D, 2 * D)), ("dv", slice(2 * D, 3 * D))):
        err = _rel_l2(dqkv[:, sl], gx[:, sl])
        assert err < 1.5e-2, (name, err)     # bf16 P / dS / outputs


@pytest.mark.parametrize("Bp,L,H,hd", [(2, 192, 3, 128), (1, 100, 2, 64), (2, 256, 1, 128)])
def test_rope_kernel_and_seq_attention_match_fused_path(ops, cuda, Bp, L, H, hd):
    """mts_rope_qk + the single-staging forward (and the pre-roped backward) against the kernels that
    rotate while staging: same arithmetic, so the results agree to bf16 rounding of the outputs."""
    g = torch.Generator().manual_seed(L + hd + 1)
    D = H * hd
    qkv = (torch.randn(Bp * L, 3 * D, generator=g) * 0.8).to(cuda, torch.bfloat16)
    dout = torch.randn(Bp * L, D, generator=g).to(cuda, torch.bfloat16)
    tabs = _rope_tables(L, hd, cuda)
    out1, lse1 = ops.attn_causal(qkv, Bp, L, H, hd, rope=tabs, want_lse=True)
    dqkv1 = ops.attn_causal_bwd(qkv, out1, dout, lse1, Bp, L, H, hd, rope=tabs)
    roped = ops.rope_qk_(qkv.clone(), Bp, L, H, hd, tabs)
    # the rotation itself: fp32 math on bf16 inputs, rounded once
    q = qkv[:, :D].float().view(Bp, L, H, hd)
    cos = torch.cat([tabs[0], tabs[0]], -1)[None, :, None]; sin = torch.cat([tabs[1], tabs[1]], -1)[None, :, None]
    rot = torch.cat([-q[..., hd // 2:], q[..., :hd // 2]], -1)
    torch.testing.assert_close(roped[:, :D].float().view(Bp, L, H, hd), q * cos + rot * sin, rtol=8e-3, atol=1e-5)
    assert torch.equal(roped[:, 2 * D:], qkv[:, 2 * D:])                     # v untouched
    out2, lse2 = ops.attn_causal(roped, Bp, L, H, hd, rope=None, want_lse=True)
    torch.testing.assert_close(out2.float(), out1.float(), rtol=2e-2, atol=4e-3)
    torch.testing.assert_close(lse2, lse1, rtol=1e-4, atol=1e-4)
    dqkv2 = ops.attn_causal_bwd(roped, out2, dout, lse2, Bp, L, H, hd, rope=tabs, pre_roped=True)
    assert _rel_l2(dqkv2, dqkv1) < 5e-3


def test_gemm_bf16_accumulate_epilogue(ops, cuda):
    """RESID_ADD with a bf16 destination: D = bf16(D + alpha * A B^T), strided into a wider buffer (the LoRA
    side GEMMs accumulate into the q / v columns of qkv)."""
    g = torch.Generator().manual_seed(51)
    m, n, k, ld = 300, 128, 8, 3 * 128
    a = torch.randn(m, 16, generator=g).to(cuda, torch.bfloat16)          # lda 16, uses columns 8..15
    b = torch.randn(n, k, generator=g).to(cuda, torch.bfloat16)
    d0 = torch.randn(m, ld, generator=g).to(cuda, torch.bfloat16)
    d = d0.clone()
    ops.gemm(a, b, d, m=m, n=n, k=k, lda=16, a_off=8, ldb=k, ldd=ld, d_off=2 * 128, alpha=0.37, epilogue=1)
    ref = d0.float()
    ref[:, 256:] += 0.37 * (a[:, 8:].float() @ b.float().t())
    assert torch.equal(d[:, :256], d0[:, :256])
    torch.testing.assert_close(d[:, 256:].float(), ref[:, 256:], rtol=8e-3, atol=8e-3)


def test_rowsum(ops, cuda):
    x = torch.randn(1024, 333, device=cuda)
    torch.testing.assert_close(ops.rowsum(x), x.sum(1), rtol=1e-5, atol=1e-4)


def test_gemm_cta_pair_and_single_cta_kernels_agree(ops, cuda):
    """The cta_group::2 kernel (default for 256-wide tiles) and the 1-CTA kernel compute the same tiles: with the
    same operands they must agree bit for bit (same MMA shape per k-step, same fp32 accumulation order)."""
    _pair_vs_single(ops, cuda, 900, 1536, 640)       # 24 pair tiles: one partial wave, no tail split
    _pair_vs_single(ops, cuda, 4864, 1024, 256)      # 76 pair tiles on 74 pairs: tail of 2 split into 4 column slices
    _pair_vs_single(ops, cuda, 5000, 2304, 192)      # 180 tiles: tail 32 -> 2 slices; ragged M


def _pair_vs_single(ops, cuda, m, n, k):
    from medtsllm_b200 import _lib
    g = torch.Generator().manual_seed(61)
    a = torch.randn(m, k, generator=g).to(cuda, torch.bfloat16)
    b = torch.randn(n, k, generator=g).to(cuda, torch.bfloat16)
    bias = torch.randn(n, generator=g).to(cuda)
    outs = []
    try:
        for flag in (1, 0):
            _lib.set_option("gemm_2cta", flag)
            d = torch.full((m, n), float("nan"), device=cuda, dtype=torch.bfloat16)
            ops.gemm(a, b, d, m=m, n=n, k=k, block_n=256, bias=bias, bias_axis=1)
            r = torch.randn(m, n, generator=torch.Generator().manual_seed(7)).to(cuda)
            ops.gemm(a, b, r, m=m, n=n, k=k, block_n=256, epilogue=1)
            outs.append((d, r))
    finally:
        _lib.set_option("gemm_2cta", 1)
    torch.testing.assert_close(outs[0][0].float(), a.float() @ b.float().t() + bias, rtol=8e-3, atol=8e-2)
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


@pytest.mark.parametrize("Bp,L,H,hd,pair", [(2, 192, 4, 128, 1), (3, 70, 6, 64, 1), (2, 192, 4, 128, 0)])
def test_gemm_rope_epilogue(ops, cuda, Bp, L, H, hd, pair):
    """Fused qkv projection + RoPE (MTS_EPI_ROPE_QK): q | k columns rotated in fp32 from the accumulators, v untouched."""
    from medtsllm_b200 import _lib
    g = torch.Generator().manual_seed(hd + L)
    D, M = H * hd, Bp * L
    x = (torch.randn(M, D, generator=g) * 0.5).to(cuda, torch.bfloat16)
    w = (torch.randn(3 * D, D, generator=g) * 0.05).to(cuda, torch.bfloat16)
    tabs = _rope_tables(max(L, 512), hd, cuda)
    out = torch.full((M, 3 * D), float("nan"), device=cuda, dtype=torch.bfloat16)
    try:
        _lib.set_option("gemm_2cta", pair)
        ops.gemm(x, w, out, m=M, n=3 * D, k=D, epilogue=4, rope=tabs, rope_L=L, rope_hd=hd, rope_cols=2 * D)
    finally:
        _lib.set_option("gemm_2cta", 1)
    ref = (x.float() @ w.float().t()).view(Bp, L, 3, H, hd)
    cos = torch.cat([tabs[0][:L], tabs[0][:L]], -1)[None, :, None, None]
    sin = torch.cat([tabs[1][:L], tabs[1][:L]], -1)[None, :, None, None]
    qk = ref[:, :, :2]
    rot = torch.cat([-qk[..., hd // 2:], qk[..., :hd // 2]], -1)
    ref = torch.cat([qk * cos + rot * sin, ref[:, :, 2:]], dim=2).reshape(M, 3 * D)
    torch.testing.assert_close(out.float(), ref, rtol=8e-3, atol=2e-2)
    assert _rel_l2(out, ref) < 3e-3


def test_swiglu_epilogue_keeps_preactivations_and_norm_bwd_bf16_copy(ops, cuda):
    g = torch.Generator().manual_seed(71)
    m, I, k = 300, 256, 192
    x = (torch.randn(m, k, generator=g) * 0.5).to(cuda, torch.bfloat16)
    wg = (torch.randn(I, k, generator=g) * 0.1).to(cuda, torch.bfloat16)
    wu = (torch.randn(I, k, generator=g) * 0.1).to(cuda, torch.bfloat16)
    packed = ops.pack_gate_up(wg, wu)
    act = torch.empty(m, I, device=cuda, dtype=torch.bfloat16)
    pre = torch.full((m, 2 * I), float("nan"), device=cuda, dtype=torch.bfloat16)
    ops.gemm(x, packed, act, m=m, n=2 * I, k=k, epilogue=3, aux=pre)
    plain = torch.empty(m, 2 * I, device=cuda, dtype=torch.bfloat16)
    ops.gemm(x, packed, plain, m=m, n=2 * I, k=k, block_n=256)
    assert torch.equal(pre, plain)                                    # same bf16 pre-activations as a plain GEMM
    assert torch.equal(act, ops.swiglu_blk(plain, I, 128))            # and the activation computed from them
    # norm backward emitting the bf16 copy of the accumulated gradient
    rows, D = 40, 512
    xx = torch.randn(rows, D, generator=g).to(cuda); w = torch.randn(D, generator=g).to(cuda)
    dy = torch.randn(rows, D, generator=g).to(cuda, torch.bfloat16)
    dx = torch.randn(rows, D, generator=g).to(cuda); dxb = torch.empty(rows, D, device=cuda, dtype=torch.bfloat16)
    ops.rmsnorm_bwd(xx, w, dy, dx, 1e-5, accumulate=True, dx_bf16=dxb)
    assert torch.equal(dxb, dx.to(torch.bfloat16))


# ------------------------------------------------------------------------------- shared-prefix row layout
def _expand_shared(t, Bp, Lc, Ls):
    """[Lc + Bp*Ls, W] (prefix once, then own rows per sample) -> [Bp*(Lc+Ls), W] plain layout."""
    W = t.shape[1]
    return torch.cat([t[:Lc].unsqueeze(0).expand(Bp, Lc, W), t[Lc:].view(Bp, Ls, W)], dim=1).reshape(Bp * (Lc + Ls), W).contiguous()


@pytest.mark.parametrize("Bp,Lc,Ls,H,hd", [(3, 128, 64, 2, 128), (4, 37, 12, 3, 64), (2, 16, 100, 2, 64), (5, 130, 33, 1, 128)])
def test_attn_causal_shared_prefix_matches_plain(ops, cuda, request, Bp, Lc, Ls, H, hd):
    """The prefix rows are stored once; results must equal the plain kernels run on the layout with the
    prefix repeated in front of every sample (same kernel, same key order: bit-identical forward)."""
    g = torch.Generator().manual_seed(Lc * 7 + Ls)
    D, L = H * hd, Lc + Ls
    M = Lc + Bp * Ls
    qkv = (torch.randn(M, 3 * D, generator=g) * 0.8).to(cuda, torch.bfloat16)
    out, lse = ops.attn_causal_shared(qkv, Bp, Lc, Ls, H, hd, want_lse=True)
    qkv_p = _expand_shared(qkv, Bp, Lc, Ls)
    out_p, lse_p = ops.attn_causal(qkv_p, Bp, L, H, hd, rope=None, want_lse=True)
    assert torch.equal(_expand_shared(out, Bp, Lc, Ls), out_p)
    lse_full = lse
    lse = ops.lse_own_view(lse_full, Bp, Lc, Ls, H)
    assert torch.equal(lse, lse_p[:, :, Lc:])
    assert torch.equal(lse_full[:H * Lc].view(H, Lc), lse_p[0, :, :Lc])
    # backward: gradient only enters through the samples' own rows.  The mma.sync kernels of both layouts share their
    # arithmetic (1e-6); the tensor-memory backward of the shared layout is checked in test_attn_tensor_memory_backward.
    from medtsllm_b200 import _lib
    _lib.set_option("attn_tc", 0)
    request.addfinalizer(lambda: _lib.set_option("attn_tc", 1))
    tabs = _rope_tables(L, hd, cuda)
    dout = torch.randn(Bp * Ls, D, generator=g).to(cuda, torch.bfloat16)
    dqkv = ops.attn_causal_shared_bwd(qkv, out[Lc:], dout, lse, Bp, Lc, Ls, H, hd, rope=tabs)
    dout_p = torch.zeros(Bp, L, D, device=cuda, dtype=torch.bfloat16)
    dout_p[:, Lc:] = dout.view(Bp, Ls, D)
    dqkv_p = ops.attn_causal_bwd(qkv_p, out_p, dout_p.view(Bp * L, D), lse_p, Bp, L, H, hd, rope=tabs, pre_roped=True)
    own = dqkv_p.view(Bp, L, 3 * D)[:, Lc:].reshape(Bp * Ls, 3 * D)
    for name, sl in (("dq", slice(0, D)), ("dk", slice(D, 2 * D)), ("dv", slice(2 * D, 3 * D))):
        assert _rel_l2(dqkv[:, sl], own[:, sl]) < 1e-6, name
    for _ in range(3):      # deterministic (no atomics; the job-to-warp assignment must not matter)
        assert torch.equal(ops.attn_causal_shared_bwd(qkv, out[Lc:], dout, lse, Bp, Lc, Ls, H, hd, rope=tabs), dqkv)
    # full backward (LoRA): gradient also enters at the prefix rows' outputs, and the prefix rows' q/k/v receive what
    # all the samples send back.  Equivalent plain computation: Bp private copies of the prefix; the gradient of the
    # shared rows is the sum over the copies (the upstream gradient of the prefix outputs is given to copy 0).
    dout_all = torch.randn(M, D, generator=g).to(cuda, torch.bfloat16)
    dqkv_all = ops.attn_causal_shared_bwd_full(qkv, out, dout_all, lse_full, Bp, Lc, Ls, H, hd, rope=tabs)
    dout_p = torch.zeros(Bp, L, D, device=cuda, dtype=torch.bfloat16)
    dout_p[:, Lc:] = dout_all[Lc:].view(Bp, Ls, D)
    dout_p[0, :Lc] = dout_all[:Lc]
    ref = ops.attn_causal_bwd(qkv_p, out_p, dout_p.view(Bp * L, D), lse_p, Bp, L, H, hd, rope=tabs, pre_roped=True)
    ref = ref.view(Bp, L, 3 * D).float()
    for name, sl in (("dq", slice(0, D)), ("dk", slice(D, 2 * D)), ("dv", slice(2 * D, 3 * D))):
        assert _rel_l2(dqkv_all[Lc:, sl], ref[:, Lc:, sl].reshape(Bp * Ls, -1)) < 1e-6, name
        e = _rel_l2(dqkv_all[:Lc, sl], ref[:, :Lc, sl].sum(0))
        assert e < 1.5e-2, (name, e)          # the plain side rounds each copy's share to bf16 before the sum


def test_prompt_gather_and_rope_shared_prefix(ops, cuda):
    g = torch.Generator().manual_seed(5)
    B, rep, Lp, Lc, N, D, V = 3, 2, 20, 17, 6, 64, 50
    L, Ls = Lp + N, Lp + N - Lc
    ids = torch.randint(0, V, (B, Lp), generator=g, dtype=torch.int32)
    ids[:, :Lc] = ids[0, :Lc]
    ids = ids.to(cuda)
    emb = torch.randn(V, D, generator=g).to(cuda)
    wpe = torch.randn(L + 3, D, generator=g).to(cuda)
    for pe in (None, wpe):
        plain = torch.full((B * rep, L, D), float("nan"), device=cuda)
        ops.prompt_gather(ids, emb, pe, plain, rep=rep, Lp=Lp, L=L)
        shared = torch.full((Lc + B * rep * Ls, D), float("nan"), device=cuda)
        ops.prompt_gather(ids, emb, pe, shared, rep=rep, Lp=Lp, L=L, Lc=Lc, B=B)
        assert torch.equal(_expand_shared(shared, B * rep, Lc, Ls).view(B * rep, L, D), plain)
    # RoPE positions: fused GEMM epilogue and the in-place kernel
    H, hd = 2, 64
    Dm = H * hd
    Bp = B * rep
    M = Lc + Bp * Ls
    tabs = _rope_tables(L, hd, cuda)
    a = torch.randn(M, 64, generator=g).to(cuda, torch.bfloat16)
    w = torch.randn(3 * Dm, 64, generator=g).to(cuda, torch.bfloat16)
    qkv = torch.empty(M, 3 * Dm, device=cuda, dtype=torch.bfloat16)
    ops.gemm(a, w, qkv, m=M, n=3 * Dm, k=64, epilogue=4, rope=tabs, rope_L=Ls, rope_hd=hd, rope_cols=2 * Dm, rope_prefix=Lc)
    a_p = _expand_shared(a, Bp, Lc, Ls)
    qkv_p = torch.empty(Bp * L, 3 * Dm, device=cuda, dtype=torch.bfloat16)
    ops.gemm(a_p, w, qkv_p, m=Bp * L, n=3 * Dm, k=64, epilogue=4, rope=tabs, rope_L=L, rope_hd=hd, rope_cols=2 * Dm)
    assert torch.equal(_expand_shared(qkv, Bp, Lc, Ls), qkv_p)
    raw = torch.empty(M, 3 * Dm, device=cuda, dtype=torch.bfloat16)
    ops.gemm(a, w, raw, m=M, n=3 * Dm, k=64)
    raw_p = _expand_shared(raw, Bp, Lc, Ls)
    ops.rope_qk_shared_(raw, Bp, Lc, Ls, H, hd, tabs)
    ops.rope_qk_(raw_p, Bp, L, H, hd, tabs)
    assert torch.equal(_expand_shared(raw, Bp, Lc, Ls), raw_p)


# --------------------------------------------------------------------------------------------- evaluation parity modes
def _tf32(x):
    """Nearest TF32 (ties away from zero), the contract of mts_round_tf32 / cvt.rna.tf32.f32."""
    i = x.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


def test_round_and_split_tf32(ops, cuda):
    g = torch.Generator().manual_seed(20)
    x = (torch.randn(1000 * 37 + 3, generator=g) * 3).to(cuda)
    x[:4] = torch.tensor([0.0, -0.0, 1.0, -2.5], device=cuda)
    r = ops.round_tf32(x)
    assert torch.equal(r, _tf32(x))                                    # bit-exact
    assert (r.view(torch.int32) & 0x1FFF).eq(0).all()
    hi, lo = ops.split_tf32(x)
    assert torch.equal(hi, r) and (lo.view(torch.int32) & 0x1FFF).eq(0).all()
    assert torch.equal(lo, _tf32(x - r))
    assert ((hi.double() + lo.double() - x.double()).abs() <= x.double().abs() * 2.0 ** -21).all()
    y = x.clone()
    assert ops.round_tf32(y, out=y) is y and torch.equal(y, r)         # in place


@pytest.mark.parametrize("m,n,k,bn", [(128, 256, 32, 256), (128, 64, 64, 64), (300, 264, 200, 0), (1, 8, 8, 64),
                                       (77, 1024, 100, 0), (2176, 4096, 1024, 0), (640, 768, 4096, 128)])
def test_gemm_tf32_and_3xtf32(ops, cuda, m, n, k, bn):
    """fp32 operands on tcgen05 kind::tf32.  Operands pre-rounded to TF32 => the only error left is the tensor core's
    accumulation: every K = 8 instruction adds its products into the fp32 TMEM accumulator with truncation, so the error
    against fp64 grows with the number of accumulation steps (measured ~1e-5 at K = 4096; bound asserted: 1e-6 * max(4,
    K/64)).  With the 3xTF32 split on full-precision operands the operand rounding (plain TF32: ~3e-4) disappears and the
    same accumulation floor is what is left (three times as many steps)."""
    g = torch.Generator().manual_seed(m + 3 * n + 7 * k)
    a = torch.randn(m, k, generator=g).to(cuda)
    b = torch.randn(n, k, generator=g).to(cuda)
    if k % 4:
        pytest.skip("lda must be a multiple of 4")
    ar, br = ops.round_tf32(a), ops.round_tf32(b)
    d = torch.full((m, n), float("nan"), device=cuda)
    ops.gemm(ar, br, d, m=m, n=n, k=k, block_n=bn)
    ref = (ar.double() @ br.double().t())
    floor = 1e-6 * max(4.0, k / 64)
    e0 = _rel_l2(d.double(), ref)
    assert e0 < floor, (e0, floor)
    torch.testing.assert_close(d.double(), ref, rtol=1e-4, atol=1e-4 * math.sqrt(k))
    # 3xTF32 on the unrounded operands
    a_hi, a_lo = ops.split_tf32(a)
    b_hi, b_lo = ops.split_tf32(b)
    d3 = torch.empty(m, n, device=cuda)
    ops.gemm(a_hi, b_hi, d3, m=m, n=n, k=k, block_n=bn, a_lo=a_lo, b_lo=b_lo)
    ref3 = a.double() @ b.double().t()
    e3, e1 = _rel_l2(d3.double(), ref3), _rel_l2(d.double(), ref3)
    print(f"\n[tf32 gemm] m={m} n={n} k={k}: exact-operand accumulation error {e0:.2e}, 3xTF32 {e3:.2e}, plain TF32 {e1:.2e}")
    assert e3 < 3 * floor, (e3, e1)
    if k >= 32:
        assert e1 > 4 * e3                                             # what the split buys over plain TF32


def test_gemm_tf32_epilogues(ops, cuda):
    g = torch.Generator().manual_seed(31)
    m, k = 200, 256
    x = ops.round_tf32((torch.randn(m, k, generator=g) * 0.5).to(cuda))
    # SwiGLU (packed gate / up rows), fp32 D rounded to TF32
    I = 320
    wg = ops.round_tf32((torch.randn(I, k, generator=g) * 0.1).to(cuda))
    wu = ops.round_tf32((torch.randn(I, k, generator=g) * 0.1).to(cuda))
    Ip = 384
    gu = torch.zeros(2, Ip, k, device=cuda)
    gu[0, :I], gu[1, :I] = wg, wu
    packed = gu.view(2, Ip // 128, 128, k).permute(1, 0, 2, 3).reshape(2 * Ip, k).contiguous()
    d = torch.full((m, Ip), float("nan"), device=cuda)
    ops.gemm(x, packed, d, m=m, n=2 * Ip, k=k, epilogue=3, round_tf32=True)
    ref = torch.nn.functional.silu(x.double() @ wg.double().t()) * (x.double() @ wu.double().t())
    torch.testing.assert_close(d[:, :I].double(), ref, rtol=6e-4, atol=1e-5)      # TF32 output rounding: 2^-11
    assert (d.view(torch.int32) & 0x1FFF).eq(0).all() and torch.all(d[:, I:] == 0)
    # gelu_new + bias, fp32 D unrounded
    n = 384
    w = ops.round_tf32((torch.randn(n, k, generator=g) * 0.3).to(cuda))
    bias = torch.randn(n, generator=g).to(cuda)
    d = torch.empty(m, n, device=cuda)
    ops.gemm(x, w, d, m=m, n=n, k=k, bias=bias, bias_axis=1, epilogue=2)
    z = x.double() @ w.double().t() + bias.double()
    ref = 0.5 * z * (1 + torch.tanh(math.sqrt(2 / math.pi) * (z + 0.044715 * z ** 3)))
    torch.testing.assert_close(d.double(), ref, rtol=1e-5, atol=1e-5)             # tanhf at library accuracy
    # residual add with bias, transposed store
    r0 = torch.randn(m, n, generator=g).to(cuda)
    r = r0.clone()
    ops.gemm(x, w, r, m=m, n=n, k=k, bias=bias, bias_axis=1, epilogue=1)
    torch.testing.assert_close(r.double(), r0.double() + z, rtol=1e-5, atol=1e-4)
    dt = torch.empty(n, m, device=cuda)
    ops.gemm(x, w, dt, m=m, n=n, k=k, d_transposed=True, ldd=m)
    torch.testing.assert_close(dt.double(), (x.double() @ w.double().t()).t(), rtol=1e-5, atol=1e-4)


@pytest.mark.parametrize("hd,Lc", [(128, 0), (128, 40), (64, 0), (64, 17)])
def test_gemm_tf32_rope_epilogue_and_attn_f32(ops, cuda, hd, Lc):
    """Fused qkv projection + RoPE (fp32 out) and the fp32 causal attention, plain and shared-prefix row layouts,
    against the oracle's fp64 arithmetic."""
    from oracle import medtsllm_oracle as O
    H, Bp, Ls = 2, 3, 45
    D = H * hd
    L = Lc + Ls
    M = Lc + Bp * Ls
    g = torch.Generator().manual_seed(hd + Lc)
    # per-sample sequences [Bp, L, D]; with a shared prefix the first Lc positions are the same rows for every sample
    xs = torch.randn(Bp, L, D, generator=g) * 0.5
    if Lc:
        xs[:, :Lc] = xs[0, :Lc]
    w = torch.randn(3 * D, D, generator=g) * (1.0 / math.sqrt(D))
    rows = torch.cat([xs[0, :Lc], xs[:, Lc:].reshape(Bp * Ls, D)], 0) if Lc else xs.reshape(Bp * L, D)
    x_dev = ops.round_tf32(rows.contiguous().to(cuda))
    w_dev = ops.round_tf32(w.to(cuda))
    cos, sin = O.rope_tables(L, hd)
    qkv = torch.empty(M, 3 * D, device=cuda)
    ops.gemm(x_dev, w_dev, qkv, m=M, n=3 * D, k=D, epilogue=4, rope=(cos.to(cuda).contiguous(), sin.to(cuda).contiguous()),
             rope_L=Ls, rope_hd=hd, rope_cols=2 * D, rope_prefix=Lc)
    out = ops.attn_causal_f32(qkv, Bp, Lc, Ls, H, hd)
    # reference in fp64 on the per-sample view
    expand = (lambda t: torch.cat([t[:Lc].unsqueeze(0).expand(Bp, Lc, -1), t[Lc:].view(Bp, Ls, -1)], 1)) if Lc else \
        (lambda t: t.view(Bp, L, -1))
    xe = expand(x_dev.cpu()).double()
    q, k, v = (xe @ w_dev.cpu().double().t()).split(D, dim=-1)
    q, k, v = (t.view(Bp, L, H, hd).transpose(1, 2) for t in (q, k, v))
    q, k = O.apply_rope(q, cos.double(), sin.double()), O.apply_rope(k, cos.double(), sin.double())
    ref = O.causal_attention(q, k, v, hd ** -0.5).transpose(1, 2).reshape(Bp, L, D)
    got_qkv = expand(qkv.cpu())
    ref_qkv = torch.cat([t.transpose(1, 2).reshape(Bp, L, D) for t in (q, k, v)], -1)
    assert _rel_l2(got_qkv.double(), ref_qkv) < 5e-6                   # q / k / v stay unrounded fp32
    assert _rel_l2(expand(out.cpu()).double(), ref) < 5e-6
    if Lc:   # the shared rows are computed once and identical for every sample by construction
        assert torch.isfinite(out).all()
    # the "tf32" mode's attention: both contractions on the tensor cores with q / k / v / P rounded to nearest TF32
    # (2^-11 relative per operand, averaging out over the head dim and the keys), fp32 softmax and accumulation
    out_t = ops.attn_causal_f32(qkv, Bp, Lc, Ls, H, hd, tf32=True)
    assert torch.isfinite(out_t).all()
    err_t = _rel_l2(expand(out_t.cpu()).double(), ref)
    assert 1e-6 < err_t < 6e-4, err_t
    out_r = ops.attn_causal_f32(qkv, Bp, Lc, Ls, H, hd, tf32=True, round_out=True)
    assert torch.equal(out_r, ops.round_tf32(out_t))


def test_softmax_rows_f32(ops, cuda):
    g = torch.Generator().manual_seed(9)
    s = (torch.randn(37, 1024, generator=g) * 4).to(cuda)
    p = ops.softmax_rows_f32(s, 0.125)
    torch.testing.assert_close(p.double(), torch.softmax(s.double() * 0.125, -1), rtol=1e-5, atol=1e-8)


# --------------------------------------------------------------------------------------------- prompt statistics
@pytest.mark.parametrize("B,T,C,f0,csel", [(5, 64, 3, 0, 3), (4, 512, 3, 1, 1), (3, 1024, 12, 0, 12), (2, 100, 25, 0, 25)])
def test_input_stats_kernel(ops, cuda, B, T, C, f0, csel):
    """mts_input_stats against the reference's own torch expressions (models/medtsllm.py:477-481, :530-538): min / max /
    median / trend exactly; the lags as a set of autocorrelation VALUES (corr[k] == corr[T-k]: the reference's FFT orders
    such pairs by rounding noise, the kernel takes the smaller index)."""
    g = torch.Generator().manual_seed(B * 1000 + T)
    t = torch.arange(T, dtype=torch.float32)
    x = (torch.randn(B, T, C, generator=g) * 0.5 + torch.sin(t / 9.0)[None, :, None] * 2 + torch.randn(B, 1, C, generator=g)).to(cuda)
    x[0, :, 0] = torch.round(x[0, :, 0])                                     # ties for the median
    stats, lags = ops.input_stats(x, f0=f0, n_features=csel, n_lags=5)
    xs = x[:, :, f0:f0 + csel]
    assert torch.equal(stats[:, :, 0], xs.min(dim=1).values) and torch.equal(stats[:, :, 1], xs.max(dim=1).values)
    assert torch.equal(stats[:, :, 2], torch.median(xs, dim=1).values)
    d = xs.double().diff(dim=1).sum(dim=1)
    clear = d.abs() > 1e-4                                                   # away from a zero sum the sign is unambiguous
    assert torch.equal((stats[:, :, 3] > 0.5)[clear], (d > 0)[clear])
    xp = xs.permute(0, 2, 1).double()
    f = torch.fft.rfft(xp, dim=-1)
    corr = torch.fft.irfft(f * torch.conj(f), dim=-1).mean(dim=1)            # [B, T]
    want = torch.topk(corr, 5, dim=-1).values
    got = torch.gather(corr, 1, lags.long())
    torch.testing.assert_close(got, want, rtol=1e-9, atol=1e-9 * corr.abs().max().item())
    assert (lags[:, 0] == 0).all()                                           # lag 0 always carries the energy
    # corr[k] == corr[T-k] exactly: both members of a pair appear, the smaller index first
    lg = lags.long().cpu()
    for b in range(B):
        seen = lg[b].tolist()
        for i, k in enumerate(seen):
            if 0 < k < T // 2 and (T - k) in seen:
                assert seen.index(T - k) > i, seen


# --------------------------------------------------------------------------------------------- stream-K
@pytest.mark.parametrize("m,n,k,epi", [(800, 4096, 4096, 1), (800, 4096, 11008, 1), (896, 1024, 4096, 1), (896, 3072, 1024, 0),
                                        (896, 4096, 1024, 2), (800, 12288, 4096, 0), (300, 520, 2048, 0), (128, 256, 8192, 0),
                                        (130, 4096, 512, 1)])
def test_gemm_streamk_matches_plain_schedule(ops, cuda, m, n, k, epi):
    """Stream-K (mts_gemm_args.sk_workspace; forced here) against the plain whole-tile schedule of the same kernel and
    against the fp32 reference: the (tile, k-block) units are dealt out evenly over one CTA per SM, partial tiles are
    summed by the tile's owner in ascending CTA order — deterministic, and equal to the plain result up to the fp32
    regrouping of the k-sum."""
    from medtsllm_b200 import _lib
    g = torch.Generator().manual_seed(m + n + k)
    a = (torch.randn(m, k, generator=g) * 0.5).to(cuda, torch.bfloat16)
    b = (torch.randn(n, k, generator=g) * 0.05).to(cuda, torch.bfloat16)
    bias = torch.randn(n, generator=g).to(cuda) if epi == 2 else None
    c0 = torch.randn(m, n, generator=g).to(cuda)

    def run(mode):
        ops.set_streamk(mode)
        d = c0.clone() if epi == 1 else torch.full((m, n), float("nan"), device=cuda, dtype=torch.bfloat16)
        ops.gemm(a, b, d, m=m, n=n, k=k, epilogue=epi, bias=bias, bias_axis=1 if bias is not None else 0)
        return d

    try:
        plain = run(0)
        sk1, sk2 = run(2), run(2)
    finally:
        ops.set_streamk(0)
    assert torch.equal(sk1, sk2)                                   # deterministic
    z = a.float() @ b.float().t()
    if epi == 1:
        ref = c0 + z
    elif epi == 2:
        z = z + bias
        ref = 0.5 * z * (1 + torch.tanh(math.sqrt(2 / math.pi) * (z + 0.044715 * z ** 3)))
    else:
        ref = z
    tol = dict(rtol=1e-4, atol=1e-3) if epi == 1 else dict(rtol=8e-3, atol=2e-2)
    torch.testing.assert_close(sk1.float(), ref, **tol)
    torch.testing.assert_close(sk1.float(), plain.float(), **tol)


@pytest.mark.parametrize("m,n,k,epi,bn", [(896, 1024, 4096, 1, 0), (896, 1024, 4096, 1, 256), (896, 1024, 1024, 1, 0),
                                           (300, 520, 2048, 0, 0), (128, 256, 8192, 2, 128), (130, 1024, 1536, 1, 64)])
def test_gemm_even_split_k(ops, cuda, m, n, k, epi, bn):
    """Even split-K (stream-K mode 3): tiles x s CTAs, each tile cut into s equal k ranges on consecutive CTAs, the first of
    which adds the others' partials in ascending order and runs the epilogue.  Deterministic, back-to-back launches reuse
    the flags, and the result equals the plain schedule up to the fp32 regrouping of the k-sum."""
    g = torch.Generator().manual_seed(m + n + k + bn)
    a = (torch.randn(m, k, generator=g) * 0.5).to(cuda, torch.bfloat16)
    b = (torch.randn(n, k, generator=g) * 0.05).to(cuda, torch.bfloat16)
    bias = torch.randn(n, generator=g).to(cuda) if epi == 2 else None
    c0 = torch.randn(m, n, generator=g).to(cuda)

    def run(mode):
        ops.set_streamk(mode)
        d = c0.clone() if epi == 1 else torch.full((m, n), float("nan"), device=cuda, dtype=torch.bfloat16)
        ops.gemm(a, b, d, m=m, n=n, k=k, epilogue=epi, bias=bias, bias_axis=1 if bias is not None else 0, block_n=bn)
        return d

    try:
        plain = run(0)
        outs = [run(3) for _ in range(4)]
    finally:
        ops.set_streamk(0)
    for o in outs[1:]:
        assert torch.equal(o, outs[0])
    z = a.float() @ b.float().t()
    if epi == 1:
        ref = c0 + z
    elif epi == 2:
        z = z + bias
        ref = 0.5 * z * (1 + torch.tanh(math.sqrt(2 / math.pi) * (z + 0.044715 * z ** 3)))
    else:
        ref = z
    tol = dict(rtol=1e-4, atol=1e-3) if epi == 1 else dict(rtol=8e-3, atol=2e-2)
    torch.testing.assert_close(outs[0].float(), ref, **tol)
    torch.testing.assert_close(outs[0].float(), plain.float(), **tol)


@pytest.mark.parametrize("m,n,k,epi,bn,ks", [
    (896, 1024, 4096, 1, 256, 4),     # GPT-2-medium MLP projection on the PSM rows: 28 tiles x 4
    (896, 1024, 4096, 1, 128, 2),
    (896, 1024, 1024, 1, 128, 2),
    (896, 1024, 4096, 1, 0, -1),      # auto
    (300, 520, 2048, 0, 128, 4),      # ragged rows / columns, one chunk per part (bf16 store)
    (300, 520, 2048, 0, 256, 2),
    (128, 256, 8192, 2, 256, 4),      # bias + gelu_new
    (130, 1000, 1536, 1, 64, 2),      # column count not a multiple of 32: staged stores on warp set 0 only
    (77, 96, 512, 0, 64, 2),
])
def test_gemm_cluster_split_k(ops, cuda, m, n, k, epi, bn, ks):
    """Cluster split-K of the single-CTA kernel (mts_set_option("gemm_ksplit")): the CTAs of a cluster run equal k ranges of
    one tile, exchange column parts through distributed shared memory and each finish one part.  Deterministic, and equal
    to the plain schedule up to the fp32 regrouping of the k-sum."""
    from medtsllm_b200 import _lib
    g = torch.Generator().manual_seed(m + n + k + bn)
    a = (torch.randn(m, k, generator=g) * 0.5).to(cuda, torch.bfloat16)
    b = (torch.randn(n, k, generator=g) * 0.05).to(cuda, torch.bfloat16)
    bias = torch.randn(n, generator=g).to(cuda) if epi == 2 else None
    c0 = torch.randn(m, n, generator=g).to(cuda)

    def run(mode):
        _lib.set_option("gemm_ksplit", mode)
        d = c0.clone() if epi == 1 else torch.full((m, n), float("nan"), device=cuda, dtype=torch.bfloat16)
        ops.gemm(a, b, d, m=m, n=n, k=k, epilogue=epi, bias=bias, bias_axis=1 if bias is not None else 0, block_n=bn)
        return d

    try:
        plain = run(0)
        outs = [run(ks) for _ in range(3)]
    finally:
        _lib.set_option("gemm_ksplit", -1)
    for o in outs[1:]:
        assert torch.equal(o, outs[0])
    z = a.float() @ b.float().t()
    if epi == 1:
        ref = c0 + z
    elif epi == 2:
        z = z + bias
        ref = 0.5 * z * (1 + torch.tanh(math.sqrt(2 / math.pi) * (z + 0.044715 * z ** 3)))
    else:
        ref = z
    tol = dict(rtol=1e-4, atol=1e-3) if epi == 1 else dict(rtol=8e-3, atol=2e-2)
    torch.testing.assert_close(outs[0].float(), ref, **tol)
    torch.testing.assert_close(outs[0].float(), plain.float(), **tol)


def test_gemm_streamk_under_graph_replay(ops, cuda):
    """The flags a launch raises are lowered again by their single reader, so a captured graph (same arguments on every
    replay) can be replayed back to back."""
    from medtsllm_b200 import _lib
    m, n, k = 800, 4096, 4096
    g = torch.Generator().manual_seed(1)
    a = (torch.randn(m, k, generator=g) * 0.5).to(cuda, torch.bfloat16)
    b = (torch.randn(n, k, generator=g) * 0.05).to(cuda, torch.bfloat16)
    d = torch.empty(m, n, device=cuda, dtype=torch.bfloat16)
    ops.set_streamk(2)
    try:
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            ops.gemm(a, b, d, m=m, n=n, k=k)
            s.synchronize()
            ref = d.clone()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=s):
                for _ in range(3):
                    ops.gemm(a, b, d, m=m, n=n, k=k)
            for _ in range(4):
                d.zero_()
                graph.replay()
                s.synchronize()
                assert torch.equal(d, ref)
    finally:
        ops.set_streamk(0)
