"""TEST INFRASTRUCTURE — CPU restatement (plain PyTorch fp32) of the reference's MedTsLLM hot path.

NOT part of the product: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import this module, and only as the checker / the timed CPU baseline.
The product path (medtsllm_b200) never routes through it.

Every function restates one piece of flixpar/med-ts-llm (paths relative to the reference tree) or of
its third-party backbone (HF: = transformers 5.5.0, the version installed in this image; the
reference's requirements.txt:13 says `transformers >= 4.39.0`, unpinned).

Parity status: PINNED.  tests/test_oracle.py checks this restatement against (a) golden stage
tensors produced by running the unmodified reference in the build container
(oracle/make_golden.py -> tests/golden/*), and (b) HuggingFace's own LlamaModel / GPT2Model executed
at test time (transformers is in the image on both machines).  The reference itself has no tests or
golden vectors.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------- RevIN (K1)
def revin_stats(x: torch.Tensor, eps: float = 1e-5):
    """models/layers/RevIN.py:37-43 — mean and sqrt(biased var + eps) over time, x [B,T,C]."""
    mean = x.mean(dim=1, keepdim=True)
    stdev = torch.sqrt(x.var(dim=1, keepdim=True, unbiased=False) + eps)
    return mean, stdev


def revin_norm(x, mean, stdev):
    """models/layers/RevIN.py:45-56 (affine=False)."""
    return (x - mean) / stdev


def revin_denorm(y, mean, stdev):
    """models/layers/RevIN.py:58-69 (affine=False)."""
    return y * stdev + mean


# ----------------------------------------------------------------------------- patching (K2)
def n_patches(T: int, P: int, S: int) -> int:
    """models/medtsllm.py:52  int((T - P) / S + 2) == unfold count on the S-padded series."""
    return (T + S - P) // S + 1


def patch_index(T: int, P: int, S: int) -> torch.Tensor:
    """Integer index map [N,P] into the ORIGINAL series: ReplicationPad1d((0,S)) then unfold(P,S)
    (models/layers/embed.py:155-163, :188-189) — padded index t reads sample min(t, T-1)."""
    N = n_patches(T, P, S)
    idx = torch.arange(N)[:, None] * S + torch.arange(P)[None, :]
    return idx.clamp_max(T - 1)


def patchify(x_bct: torch.Tensor, P: int, S: int) -> torch.Tensor:
    """x [B,C,T] -> [B*C, N, P]  (models/layers/embed.py:186-190)."""
    B, C, T = x_bct.shape
    idx = patch_index(T, P, S)
    return x_bct[:, :, idx].reshape(B * C, idx.shape[0], P)


def token_conv(patches: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """TokenEmbedding: Conv1d(P->d, k=3, circular over the patch axis, bias=False)
    (models/layers/embed.py:29-46): out[n] = W[:,:,0] p[n-1] + W[:,:,1] p[n] + W[:,:,2] p[n+1]."""
    prev = torch.roll(patches, 1, dims=1)
    nxt = torch.roll(patches, -1, dims=1)
    return prev @ w[:, :, 0].T + patches @ w[:, :, 1].T + nxt @ w[:, :, 2].T


# ----------------------------------------------------------------------------- reprogramming (K3/K4)
def mapping(word_emb, w, b):
    """models/medtsllm.py:281: mapping_layer(E^T)^T -> source [num_tokens, D]."""
    return w @ word_emb + b[:, None]


def reprogramming(x, source, sd, n_heads: int, prefix="reprogramming_layer.", attn_mask=None):
    """models/medtsllm.py:566-591.  `attn_mask` [B, H, L, S]: the multiplicative dropout mask (0 or 1/(1-p)) applied
    to the attention probabilities in train mode (:587); None = eval / p = 0."""
    B, L, _ = x.shape
    S = source.shape[0]
    q = F.linear(x, sd[prefix + "query_projection.weight"], sd[prefix + "query_projection.bias"]).view(B, L, n_heads, -1)
    k = F.linear(source, sd[prefix + "key_projection.weight"], sd[prefix + "key_projection.bias"]).view(S, n_heads, -1)
    v = F.linear(source, sd[prefix + "value_projection.weight"], sd[prefix + "value_projection.bias"]).view(S, n_heads, -1)
    scale = 1.0 / math.sqrt(q.shape[-1])
    scores = torch.einsum("blhe,she->bhls", q, k)
    A = torch.softmax(scale * scores, dim=-1)
    if attn_mask is not None:
        A = A * attn_mask
    out = torch.einsum("bhls,she->blhe", A, v).reshape(B, L, -1)
    return F.linear(out, sd[prefix + "out_projection.weight"], sd[prefix + "out_projection.bias"])


# ----------------------------------------------------------------------------- prompt (K5)
def assemble_prompt(ids_per_sample, word_emb, pad_id: int, encode_part=None):
    """models/medtsllm.py:299-319, 331-337: embed every part, cat, LEFT-pad with the pad-token embedding, stack.
    A sample's prompt is a flat list whose items are token ids or — `prompting.examples` — time-series tensors
    [1, T_ex, C], which `encode_part` (:313-319 -> encode_ts) turns into their reprogrammed patch rows [N_ex, D]."""
    B = len(ids_per_sample)
    D = word_emb.shape[1]
    rows = []
    for items in ids_per_sample:
        pieces, run = [], []
        for it in items:
            if isinstance(it, torch.Tensor):
                if run:
                    pieces.append(word_emb[torch.as_tensor(run, dtype=torch.long)])
                    run = []
                pieces.append(encode_part(it)[0])
            else:
                run.append(int(it))
        if run:
            pieces.append(word_emb[torch.as_tensor(run, dtype=torch.long)])
        rows.append(torch.cat(pieces, dim=0) if pieces else word_emb.new_zeros(0, D))
    Lp = max((r.shape[0] for r in rows), default=0)
    out = []
    for r in rows:
        pad = Lp - r.shape[0]
        out.append(torch.cat([word_emb[pad_id][None].expand(pad, D), r], dim=0) if pad else r)
    return torch.stack(out) if B else word_emb.new_zeros(0, Lp, D)


# ----------------------------------------------------------------------------- Llama backbone (K6-K10)
def rmsnorm(x, w, eps):
    """HF:models/llama/modeling_llama.py:61-67."""
    var = x.float().pow(2).mean(-1, keepdim=True)
    return w * (x.float() * torch.rsqrt(var + eps)).to(x.dtype)


def rope_tables(L: int, head_dim: int, theta: float = 10000.0):
    """HF:models/llama/modeling_llama.py:107-136 (default rope): cos/sin [L, head_dim/2] fp32."""
    inv_freq = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.int64).float() / head_dim))
    freqs = torch.arange(L, dtype=torch.float32)[:, None] * inv_freq[None, :]
    return freqs.cos(), freqs.sin()


def apply_rope(t, cos, sin):
    """HF:models/llama/modeling_llama.py:139-168 — rotate-half; t [B,H,L,hd]."""
    hd = t.shape[-1]
    cos = torch.cat([cos, cos], -1)[None, None]
    sin = torch.cat([sin, sin], -1)[None, None]
    rot = torch.cat([-t[..., hd // 2:], t[..., : hd // 2]], dim=-1)
    return t * cos + rot * sin


def causal_attention(q, k, v, scale, prob_mask=None):
    """HF:models/llama/modeling_llama.py:199-221 / HF:models/gpt2/modeling_gpt2.py:54-72 — eager,
    causal, no padding mask (models/medtsllm.py:350 passes none).  `prob_mask` [B, H, L, L]: the multiplicative
    dropout mask (0 or 1/(1-p)) HF applies to the probabilities in train mode (attn_dropout, gpt2 :67-68; llama :217)."""
    L = q.shape[-2]
    s = (q @ k.transpose(-1, -2)) * scale
    mask = torch.ones(L, L, dtype=torch.bool, device=q.device).tril()
    s = s.masked_fill(~mask, torch.finfo(s.dtype).min)
    p = torch.softmax(s.float(), dim=-1).to(q.dtype)
    if prob_mask is not None:
        p = p * prob_mask
    return p @ v


def lora_delta(h, lora, layer: int, t: int):
    """peft LoRA (restated from its published arithmetic; peft is absent here, models/medtsllm.py:187-204):
    scale * B (A h), scale = alpha/sqrt(r) with rsLoRA.  `lora` = {"scale", "n_targets", "A": [...], "B": [...]}
    with the pairs ordered layer-major, target-minor."""
    i = layer * lora["n_targets"] + t
    return lora["scale"] * F.linear(F.linear(h, lora["A"][i]), lora["B"][i])


def llama_forward(x, sd, *, n_layers: int, n_heads: int, eps: float, theta: float = 10000.0,
                  return_hidden: bool = False, lora=None, dropout=None):
    """HF LlamaModel.forward on inputs_embeds (HF:models/llama/modeling_llama.py:375-425, decoder
    layer :303-332, attention :251-300, MLP :171-184).  `sd` uses HF parameter names."""
    B, L, D = x.shape
    hd = D // n_heads
    cos, sin = (t.to(x.device) for t in rope_tables(L, hd, theta))
    hidden = [x]
    for i in range(n_layers):
        p = f"layers.{i}."
        h = rmsnorm(x, sd[p + "input_layernorm.weight"], eps)
        q = F.linear(h, sd[p + "self_attn.q_proj.weight"])
        k = F.linear(h, sd[p + "self_attn.k_proj.weight"])
        v = F.linear(h, sd[p + "self_attn.v_proj.weight"])
        if lora is not None:                      # peft default targets for Llama: q_proj, v_proj
            q = q + lora_delta(h, lora, i, 0)
            v = v + lora_delta(h, lora, i, 1)
        q, k, v = (t.view(B, L, n_heads, hd).transpose(1, 2) for t in (q, k, v))
        q, k = apply_rope(q, cos, sin), apply_rope(k, cos, sin)
        a = causal_attention(q, k, v, hd ** -0.5, dropout["attn"][i] if dropout is not None else None)
        a = a.transpose(1, 2).reshape(B, L, D)
        x = x + F.linear(a, sd[p + "self_attn.o_proj.weight"])
        h = rmsnorm(x, sd[p + "post_attention_layernorm.weight"], eps)
        g = F.linear(h, sd[p + "mlp.gate_proj.weight"])
        u = F.linear(h, sd[p + "mlp.up_proj.weight"])
        x = x + F.linear(F.silu(g) * u, sd[p + "mlp.down_proj.weight"])
        hidden.append(x)
    out = rmsnorm(x, sd["norm.weight"], eps)
    return (out, hidden) if return_hidden else out


# ----------------------------------------------------------------------------- GPT-2 backbone
def gelu_new(x):
    """HF:activations.py:59-66."""
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * torch.pow(x, 3.0))))


def conv1d_hf(x, w, b):
    """HF:pytorch_utils.py:97-123 — Conv1D: addmm(bias, x, W) with W [in, out]."""
    return x @ w + b


def gpt2_forward(x, sd, *, n_layers: int, n_heads: int, eps: float = 1e-5, return_hidden: bool = False,
                 lora=None, dropout=None):
    """HF GPT2Model.forward on inputs_embeds (HF:models/gpt2/modeling_gpt2.py:522-636: + wpe[0..L) :584-585; block
    :262-309; attention :144-236; MLP :238-243; ln_f :628).  `dropout` = the train-mode masks (multiplicative, 0 or
    1/(1-p)): {"embd": [B,L,D] (:612), "attn": [layers][B,H,L,L] (:67-68), "resid_attn" / "resid_mlp": [layers][B,L,D]
    (:233, :243)}; None = eval mode."""
    B, L, D = x.shape
    hd = D // n_heads
    x = x + sd["wpe.weight"][:L][None]
    if dropout is not None:
        x = x * dropout["embd"]
    hidden = [x]
    for i in range(n_layers):
        p = f"h.{i}."
        h = F.layer_norm(x, (D,), sd[p + "ln_1.weight"], sd[p + "ln_1.bias"], eps)
        qkv = conv1d_hf(h, sd[p + "attn.c_attn.weight"], sd[p + "attn.c_attn.bias"])
        if lora is not None:                      # peft default target for GPT-2: the fused c_attn
            qkv = qkv + lora_delta(h, lora, i, 0)
        q, k, v = (t.view(B, L, n_heads, hd).transpose(1, 2) for t in qkv.split(D, dim=-1))
        a = causal_attention(q, k, v, 1.0 / math.sqrt(hd), dropout["attn"][i] if dropout is not None else None)
        a = conv1d_hf(a.transpose(1, 2).reshape(B, L, D), sd[p + "attn.c_proj.weight"], sd[p + "attn.c_proj.bias"])
        x = x + (a * dropout["resid_attn"][i] if dropout is not None else a)
        h = F.layer_norm(x, (D,), sd[p + "ln_2.weight"], sd[p + "ln_2.bias"], eps)
        h = gelu_new(conv1d_hf(h, sd[p + "mlp.c_fc.weight"], sd[p + "mlp.c_fc.bias"]))
        h = conv1d_hf(h, sd[p + "mlp.c_proj.weight"], sd[p + "mlp.c_proj.bias"])
        x = x + (h * dropout["resid_mlp"][i] if dropout is not None else h)
        hidden.append(x)
    out = F.layer_norm(x, (D,), sd["ln_f.weight"], sd["ln_f.bias"], eps)
    return (out, hidden) if return_hidden else out


# ----------------------------------------------------------------------------- whole path
def medtsllm_forward(x_enc, prompt_ids, adapters, backbone_sd, spec, *, training: bool = False,
                     return_stages: bool = False, lora=None, dropout_masks=None):
    """MedTsLLM.forward/predict (models/medtsllm.py:248-261, 321-384): all seven covariate modes, all three
    down-sample modes, optional LoRA; dropout = 0.

    spec keys: task, pred_len, patch_len, stride, d_model (per-feature), d_ff, n_heads, covariate_mode,
    downsample, n_outputs_per_step, backbone ("llama"|"gpt2"), n_layers, llm_heads, eps, rope_theta,
    pad_id, seg_mode, n_classes."""
    if x_enc.ndim == 2:
        x_enc = x_enc.unsqueeze(-1)
    B, T, C = x_enc.shape
    P, S, dm = spec["patch_len"], spec["stride"], spec["d_model"]
    emb_key = "embed_tokens.weight" if spec["backbone"] == "llama" else "wte.weight"
    word_emb = backbone_sd[emb_key]
    stages = {}

    def encode_ts(x, masks=None):
        """models/medtsllm.py:262-297 — also run on every time-series example part of the prompt (:313-319)."""
        bs = x.shape[0]
        mean_, stdev_ = revin_stats(x)
        xn = revin_norm(x, mean_, stdev_).permute(0, 2, 1).contiguous()
        e = token_conv(patchify(xn, P, S), adapters["patch_embedding.value_embedding.tokenConv.weight"])
        n = e.shape[1]
        if masks is not None and "patch" in masks:   # PatchEmbedding.dropout (models/layers/embed.py:197), [B*C, N, dm]
            e = e * masks["patch"]
        pe = e
        if mode == "concat":
            e = e.reshape(bs, C, n, dm).permute(0, 2, 1, 3).reshape(bs, n, C * dm)
        elif mode == "univariate":
            assert C == 1
        elif mode not in ("interleave", "independent", "merge-end", "add", "weighted-average"):
            raise ValueError(mode)
        e = reprogramming(e, source, adapters, spec["n_heads"], attn_mask=masks.get("reprog") if masks is not None else None)
        rp = e
        Dm = e.shape[-1]
        if mode == "add":                                                  # models/medtsllm.py:284-286
            e = e.reshape(bs, C, n, Dm).mean(dim=1)
        elif mode == "weighted-average":                                   # :287-291
            e = F.linear(e.reshape(bs, C, n, Dm).permute(0, 2, 3, 1), adapters["feature_weighting.weight"],
                         adapters["feature_weighting.bias"]).squeeze(-1)
        elif mode == "interleave":                                         # :292-295  (n-major, c-minor)
            e = e.reshape(bs, C, n, Dm).permute(0, 2, 1, 3).reshape(bs, n * C, Dm)
        return e, mean_, stdev_, pe, rp

    mode = spec["covariate_mode"]
    source = mapping(word_emb, adapters["mapping_layer.weight"], adapters["mapping_layer.bias"])
    stages["source_embeddings"] = source
    ex_masks = dropout_masks.get("examples") if dropout_masks is not None else None   # per example part, in prompt order
    ex_counter = [0]

    def encode_example(ts):
        m = ex_masks[ex_counter[0]] if ex_masks is not None else None
        ex_counter[0] += 1
        return encode_ts(ts.unsqueeze(-1) if ts.ndim == 2 else ts, m)[0]

    # prompt first (its example parts go through encode_ts before the window does, models/medtsllm.py:330-341)
    prompt = assemble_prompt(prompt_ids, word_emb, spec["pad_id"], encode_example)
    enc, mean, stdev, pe, rp = encode_ts(x_enc, dropout_masks)
    stages["patch_embedding"] = pe
    stages["reprogramming_layer"] = rp                                 # (the module's own output, pre-merge)
    N = enc.shape[1]

    # prompt + backbone (models/medtsllm.py:330-351)
    if mode in ("independent", "merge-end"):                           # models/medtsllm.py:343-344
        prompt = prompt.repeat_interleave(C, dim=0)
    Bp = enc.shape[0]
    llm_in = torch.cat([prompt, enc], dim=1)
    stages["llm_input"] = llm_in
    bb_masks = dropout_masks.get("backbone") if dropout_masks is not None else None
    if spec["backbone"] == "llama":
        dec, hidden = llama_forward(llm_in, backbone_sd, n_layers=spec["n_layers"], n_heads=spec["llm_heads"],
                                    eps=spec["eps"], theta=spec.get("rope_theta", 10000.0), return_hidden=True, lora=lora,
                                    dropout=bb_masks)
    else:
        dec, hidden = gpt2_forward(llm_in, backbone_sd, n_layers=spec["n_layers"], n_heads=spec["llm_heads"],
                                   eps=spec["eps"], return_hidden=True, lora=lora, dropout=bb_masks)
    stages["llm"] = dec
    stages["llm.hidden_states"] = hidden

    # head (models/medtsllm.py:353-382)
    dec = dec[:, -N:, :]
    if spec["downsample"] == "truncate":
        dec = dec[:, :, : spec["d_ff"]]
    elif spec["downsample"] == "linear":
        dec = F.linear(dec, adapters["embedding_downsample_layer.weight"], adapters["embedding_downsample_layer.bias"])
    elif spec["downsample"] == "average":
        dec = dec.reshape(Bp, N, spec["d_ff"], -1).mean(dim=-1)
    stages["downsample"] = dec
    dec = dec.permute(0, 2, 1).contiguous().flatten(start_dim=-2)          # k = f*N + n
    dec = F.linear(dec, adapters["output_projection.linear.weight"], adapters["output_projection.linear.bias"])
    stages["output_projection"] = dec
    nops = spec["n_outputs_per_step"]
    if mode == "independent":                                          # models/medtsllm.py:369-371
        dec = dec.view(B, C, spec["pred_len"], nops).mean(dim=1)
    elif mode == "merge-end":                                          # :372-375
        dec = dec.view(B, C, spec["pred_len"], nops).permute(0, 2, 3, 1).reshape(B, spec["pred_len"], -1)
        dec = F.linear(dec, adapters["feature_weighting.weight"], adapters["feature_weighting.bias"])
    dec = dec.reshape(B, spec["pred_len"], nops)
    if spec["task"] in ("forecasting", "reconstruction", "anomaly_detection", "pretraining"):
        dec = revin_denorm(dec, mean, stdev)
    else:
        dec = dec.squeeze(-1)
    if not training:  # models/medtsllm.py:251-259
        if spec["task"] == "semantic_segmentation":
            dec = F.softmax(dec, dim=-1) if spec.get("n_classes", 0) > 2 else torch.sigmoid(dec)
        elif spec["task"] == "segmentation" and spec.get("seg_mode") == "boundary-prediction":
            dec = torch.sigmoid(dec)
    stages["revin_mean"], stages["revin_stdev"] = mean, stdev
    stages["output"] = dec
    return (dec, stages) if return_stages else dec
