"""CPU oracle for the mu2-LLM hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain fp32 PyTorch restatement of the reference's forward arithmetic, written functionally over a
state dict that uses the REFERENCE parameter names. Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference leg may import this module; the product path
(u2tokenizer_b200/) never does.

Pinning (see DESIGN.md "Oracle"):
  * mu2-Tokenizer, SpatialPoolingProjector and the multimodal splice are checked against the
    reference modules imported unmodified from /root/reference (tests/test_oracle_pin.py; runs in the
    authoring container where the reference is mounted) and against committed golden tensors
    (tests/golden/, made by tools/make_golden.py from the REFERENCE modules).
  * The ViT3D tower's arithmetic lives in MONAI 1.3.0 (third-party, not vendored, not installed):
    restated from MONAI 1.3.0's published PatchEmbeddingBlock("perceptron") / SABlock / MLPBlock /
    TransformerBlock definitions; the reference holds no test or golden vector for it ->
    "parity unpinned" for that stage (anchored on the call sites
    src/model/multimodal_encoder/vit.py:90-105,114-126,143-158). Cross-checks that do exist (tests/test_oracle_pin.py):
    the block against the installed HF transformers ViTLayer, an independent implementation of the same pre-LN ViT
    block, and the brick gather against einops running MONAI's published pattern string.
  * The decoder is checked against the installed HuggingFace transformers Qwen3/Llama
    implementation (the reference calls it through super().forward, u2llama.py:76-87).

Every function cites the reference file:line it follows (paths relative to the reference repo).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]

# bench.py's same-box GPU comparator ("stock PyTorch eager": cuBLAS + SDPA / flash attention) flips this; the parity
# checks always run with the explicit softmax(QK^T)V below
USE_SDPA = False


def _sdpa(q, k, v, scale, bias=None, causal=False):
    """softmax(q k^T * scale + bias) v for [b, h, s, d] tensors: explicit (the reference's arithmetic) or, for the GPU
    eager baseline, torch's fused scaled_dot_product_attention."""
    if USE_SDPA:
        if k.shape[1] != q.shape[1]:
            return F.scaled_dot_product_attention(q, k, v, attn_mask=bias, is_causal=causal, scale=scale, enable_gqa=True)
        return F.scaled_dot_product_attention(q, k, v, attn_mask=bias, is_causal=causal, scale=scale)
    if k.shape[1] != q.shape[1]:
        rep = q.shape[1] // k.shape[1]
        k, v = k.repeat_interleave(rep, dim=1), v.repeat_interleave(rep, dim=1)
    s = q @ k.transpose(-2, -1) * scale
    if bias is not None:
        s = s + bias
    if causal:
        sq, sk = q.shape[-2], k.shape[-2]
        s = s + torch.full((sq, sk), float("-inf"), device=q.device, dtype=s.dtype).triu(diagonal=sk - sq + 1)
    return torch.softmax(s, dim=-1) @ v


def _lin(x: torch.Tensor, sd: SD, name: str, bias: bool = True) -> torch.Tensor:
    b = sd.get(name + ".bias") if bias else None
    return F.linear(x, sd[name + ".weight"], b)


# ---------------------------------------------------------------------------------------------
# ViT3D tower (MONAI 1.3.0 blocks; call sites src/model/multimodal_encoder/vit.py:90-126)
# ---------------------------------------------------------------------------------------------
def patch_embed(sd: SD, pre: str, x: torch.Tensor, patch_size) -> torch.Tensor:
    """MONAI PatchEmbeddingBlock(pos_embed="perceptron"): einops
    "b c (h p1) (w p2) (d p3) -> b (h w d) (p1 p2 p3 c)" then Linear + learned position embedding
    (vit.py:90-99,115)."""
    b, c, H, W, D = x.shape
    p1, p2, p3 = patch_size
    h, w, d = H // p1, W // p2, D // p3
    x = x.view(b, c, h, p1, w, p2, d, p3).permute(0, 2, 4, 6, 3, 5, 7, 1).reshape(b, h * w * d, p1 * p2 * p3 * c)
    x = _lin(x, sd, pre + "patch_embedding.patch_embeddings.1")
    return x + sd[pre + "patch_embedding.position_embeddings"]


def vit_block(sd: SD, pre: str, x: torch.Tensor, num_heads: int) -> torch.Tensor:
    """MONAI TransformerBlock: x += attn(norm1(x)); x += mlp(norm2(x)); SABlock with fused qkv
    Linear (no bias), einsum attention scaled by head_dim**-0.5, out_proj; MLPBlock Linear-GELU-Linear
    (vit.py:100-105,120-122)."""
    b, s, hid = x.shape
    dh = hid // num_heads
    y = F.layer_norm(x, (hid,), sd[pre + "norm1.weight"], sd[pre + "norm1.bias"], 1e-5)
    qkv = F.linear(y, sd[pre + "attn.qkv.weight"], sd.get(pre + "attn.qkv.bias"))
    # Rearrange("b h (qkv l d) -> qkv b l h d", qkv=3, l=num_heads)
    qkv = qkv.view(b, s, 3, num_heads, dh).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    if USE_SDPA:
        o = _sdpa(q, k, v, dh ** -0.5).permute(0, 2, 1, 3).reshape(b, s, hid)
    else:
        att = torch.softmax(torch.einsum("blxd,blyd->blxy", q, k) * (dh ** -0.5), dim=-1)
        o = torch.einsum("bhxy,bhyd->bhxd", att, v).permute(0, 2, 1, 3).reshape(b, s, hid)
    x = x + _lin(o, sd, pre + "attn.out_proj")
    y = F.layer_norm(x, (hid,), sd[pre + "norm2.weight"], sd[pre + "norm2.bias"], 1e-5)
    y = _lin(F.gelu(_lin(y, sd, pre + "mlp.linear1")), sd, pre + "mlp.linear2")
    return x + y


def vit3d_tower(sd: SD, pre: str, images: torch.Tensor, cfg) -> torch.Tensor:
    """ViT3DTower.forward (vit.py:148-164) over ViT.forward (vit.py:114-126): patch embedding,
    cls token prepended, 12 blocks, final LayerNorm, 'patch' feature selection drops the cls row."""
    vp = pre + "vision_tower."
    x = patch_embed(sd, vp, images, cfg.patch_size)
    cls = sd[vp + "cls_token"].expand(x.shape[0], -1, -1)
    x = torch.cat((cls, x), dim=1)
    n_layers = cfg.vit_layers
    for i in range(n_layers):
        x = vit_block(sd, f"{vp}blocks.{i}.", x, cfg.vit_heads)
    hid = x.shape[-1]
    x = F.layer_norm(x, (hid,), sd[vp + "norm.weight"], sd[vp + "norm.bias"], 1e-5)
    if cfg.vision_select_feature == "patch":
        x = x[:, 1:]
    return x


# ---------------------------------------------------------------------------------------------
# SpatialPoolingProjector (src/model/multimodal_projector/spatial_pooling_projector.py:34-52)
# ---------------------------------------------------------------------------------------------
def spatial_pooling_projector(sd: SD, pre: str, x: torch.Tensor, cfg) -> torch.Tensor:
    b, n, dch = x.shape
    g = [i // p for i, p in zip(cfg.image_size, cfg.patch_size)]
    ps = cfg.proj_pooling_size
    if cfg.proj_pooling_type == "spatial":
        x = x.view(b, g[0], g[1], g[2], dch).permute(0, 4, 1, 2, 3)
        x = F.avg_pool3d(x, kernel_size=ps, stride=ps)
        x = x.permute(0, 2, 3, 4, 1).reshape(b, -1, dch)
    elif cfg.proj_pooling_type == "sequence":
        x = F.avg_pool1d(x.permute(0, 2, 1), kernel_size=ps ** 3, stride=ps ** 3).permute(0, 2, 1)
    depth = int(cfg.proj_layer_num)
    x = _lin(x, sd, pre + "projector.0")
    for i in range(1, depth):
        if cfg.proj_layer_type == "mlp":
            x = F.gelu(x)
            x = _lin(x, sd, pre + f"projector.{2 * i}")
        else:
            x = _lin(x, sd, pre + f"projector.{i}")
    return x


# ---------------------------------------------------------------------------------------------
# mu2-Tokenizer attention flavours
# ---------------------------------------------------------------------------------------------
def _split_heads(x: torch.Tensor, h: int) -> torch.Tensor:
    b, s, e = x.shape
    return x.view(b, s, h, e // h).permute(0, 2, 1, 3)


def rma(sd: SD, pre: str, x: torch.Tensor, h: int, max_seq_len: int = 512) -> torch.Tensor:
    """RelativeMultiheadAttention.forward (src/model/u2tokenizer/rma.py:46-82), self-attention use:
    softmax(QK^T/sqrt(dh) + bias[j-i+max_seq_len-1, head]) V, then dense."""
    b, s, e = x.shape
    dh = e // h
    q = _split_heads(_lin(x, sd, pre + "wq"), h)
    k = _split_heads(_lin(x, sd, pre + "wk"), h)
    v = _split_heads(_lin(x, sd, pre + "wv"), h)
    pos = torch.arange(s, device=x.device)
    idx = pos[None, :] - pos[:, None] + max_seq_len - 1
    if USE_SDPA:
        ctx = _sdpa(q, k, v, 1.0 / math.sqrt(dh), bias=sd[pre + "relative_bias"][idx].permute(2, 0, 1).unsqueeze(0).to(q.dtype))
    else:
        scores = q @ k.transpose(-2, -1) / math.sqrt(dh)
        scores = scores + sd[pre + "relative_bias"][idx].permute(2, 0, 1).unsqueeze(0)
        ctx = torch.softmax(scores, dim=-1) @ v
    ctx = ctx.permute(0, 2, 1, 3).reshape(b, s, e)
    return _lin(ctx, sd, pre + "dense")


def _rotate_half(x):
    d = x.shape[-1] // 2
    return torch.cat((-x[..., d:], x[..., :d]), dim=-1)


def rope_mha(sd: SD, pre: str, x: torch.Tensor, h: int) -> torch.Tensor:
    """RotaryMultiheadAttention.forward (src/model/u2tokenizer/rope.py:62-91): rotate-half RoPE,
    theta 10000, positions 0..S-1 on q and k."""
    b, s, e = x.shape
    dh = e // h
    q = _split_heads(_lin(x, sd, pre + "wq"), h)
    k = _split_heads(_lin(x, sd, pre + "wk"), h)
    v = _split_heads(_lin(x, sd, pre + "wv"), h)
    inv = 1.0 / (10000 ** (torch.arange(0, dh, 2, dtype=torch.float32, device=x.device) / dh))
    fr = torch.outer(torch.arange(s, dtype=torch.float32, device=x.device), inv)
    emb = torch.cat((fr, fr), dim=-1)
    cos, sin = emb.cos()[None, None].to(x.dtype), emb.sin()[None, None].to(x.dtype)
    q = q * cos + _rotate_half(q) * sin
    k = k * cos + _rotate_half(k) * sin
    ctx = torch.softmax(q @ k.transpose(-2, -1) / math.sqrt(dh), dim=-1) @ v
    return _lin(ctx.permute(0, 2, 1, 3).reshape(b, s, e), sd, pre + "dense")


def mha_seq_first(sd: SD, pre: str, x: torch.Tensor, h: int) -> torch.Tensor:
    """The fallback of src/model/u2tokenizer/svr.py:17-18 and tta.py:83-84: any attn_type other than "rma" /
    "rope" builds torch.nn.MultiheadAttention(E, heads) with its default batch_first=False and calls it as
    attn(x, x, x), so dim 0 of x is the SEQUENCE and dim 1 the batch: spatial attention mixes the (batch, frame)
    axis per token, temporal attention mixes (batch, token) per frame, the TTA self-attention mixes the batch
    samples per query. Restated from the published module (packed in_proj_weight [3E, E] = q | k | v, in_proj_bias,
    out_proj, scale 1/sqrt(dh), no mask, dropout 0)."""
    a, bt, e = x.shape
    dh = e // h
    qkv = x @ sd[pre + "in_proj_weight"].to(x.dtype).T + sd[pre + "in_proj_bias"].to(x.dtype)
    q, k, v = (t.reshape(a, bt, h, dh).permute(1, 2, 0, 3) for t in qkv.split(e, dim=-1))  # [bt, h, a, dh]
    if USE_SDPA:
        ctx = _sdpa(q, k, v, 1.0 / math.sqrt(dh))
    else:
        ctx = torch.softmax(q @ k.transpose(-2, -1) / math.sqrt(dh), dim=-1) @ v
    ctx = ctx.permute(2, 0, 1, 3).reshape(a, bt, e)
    return ctx @ sd[pre + "out_proj.weight"].to(x.dtype).T + sd[pre + "out_proj.bias"].to(x.dtype)


def self_attn(sd: SD, pre: str, x: torch.Tensor, h: int, attn_type: str) -> torch.Tensor:
    if attn_type == "rma":
        return rma(sd, pre, x, h)
    if attn_type == "rope":
        return rope_mha(sd, pre, x, h)
    return mha_seq_first(sd, pre, x, h)


def cross_attn(sd: SD, pre: str, q_in: torch.Tensor, kv_in: torch.Tensor, h: int,
               is_compress: bool = False) -> torch.Tensor:
    """MultiHeadCrossAttention.forward (src/model/u2tokenizer/tta.py:42-69): no mask; with
    is_compress the values are the raw kv input and the output projection is skipped."""
    b, sq, e = q_in.shape
    dh = e // h
    q = _split_heads(_lin(q_in, sd, pre + "wq"), h)
    k = _split_heads(_lin(kv_in, sd, pre + "wk"), h)
    v = _split_heads(kv_in if is_compress else _lin(kv_in, sd, pre + "wv"), h)
    if USE_SDPA:
        ctx = _sdpa(q, k, v, 1.0 / math.sqrt(dh))
    else:
        ctx = torch.softmax(q @ k.transpose(-2, -1) / math.sqrt(dh), dim=-1) @ v
    ctx = ctx.permute(0, 2, 1, 3).reshape(b, sq, e)
    return ctx if is_compress else _lin(ctx, sd, pre + "dense")


# ---------------------------------------------------------------------------------------------
# SVR: spatio-temporal refiner, token selection, multi-scale pooling (src/model/u2tokenizer/svr.py)
# ---------------------------------------------------------------------------------------------
def svr_layer(sd: SD, pre: str, x: torch.Tensor, h: int, attn_type: str) -> torch.Tensor:
    """SpatioTemporalAttentionLayer.forward (svr.py:23-40): spatial attention over tokens within a
    frame, then temporal attention over frames per token; no residual / norm / FFN."""
    b, t, n, e = x.shape
    x = self_attn(sd, pre + "spatial_attention.", x.reshape(b * t, n, e), h, attn_type).view(b, t, n, e)
    x = x.permute(0, 2, 1, 3).reshape(b * n, t, e)
    x = self_attn(sd, pre + "temporal_attention.", x, h, attn_type)
    return x.view(b, n, t, e).permute(0, 2, 1, 3).contiguous()


def token_selection_hard(sd: SD, pre: str, x: torch.Tensor, top_k: int) -> torch.Tensor:
    """TokenSelection.forward (svr.py:75-91): Linear(E->1) scores, top-k over frames*tokens, gather."""
    b, t, n, e = x.shape
    scores = _lin(x, sd, pre + "score_net").squeeze(-1).view(b, -1)
    _, idx = torch.topk(scores, top_k, dim=1)
    return x.view(b, t * n, e)[torch.arange(b, device=x.device).unsqueeze(1), idx]


def token_selection_diff(sd: SD, pre: str, x: torch.Tensor, tau: float = 1.0) -> torch.Tensor:
    """DifferentiableTokenSelection.forward (svr.py:101-117). The reference loops over the top_k
    selection heads summing w[:, :, r] * x; that is exactly softmax_tokens(scores)^T @ x."""
    b, t, n, e = x.shape
    scores = _lin(x, sd, pre + "score_net").view(b, t * n, -1)
    w = torch.softmax(scores / tau, dim=1)
    return w.transpose(1, 2) @ x.view(b, t * n, e)


def multi_scale_pool(sd: SD, pre: Optional[str], x: torch.Tensor, dynamic: bool, scales=(1, 2, 4)) -> torch.Tensor:
    """DynamicMultiScalePooling.forward (svr.py:126-151) when `dynamic`, else the plain multi-scale
    concat (svr.py:175-184)."""
    pooled, gates = [], []
    for s in scales:
        if x.size(1) >= s:
            p = F.avg_pool1d(x.permute(0, 2, 1), kernel_size=s, stride=s).permute(0, 2, 1)
            pooled.append(p)
            if dynamic:
                gates.append(_lin(p.mean(dim=1), sd, pre + "gate_fc"))
    if not dynamic:
        return torch.cat(pooled, dim=1)
    w = torch.softmax(torch.cat(gates, dim=1), dim=1)
    return torch.cat([p * w[:, i].view(-1, 1, 1) for i, p in enumerate(pooled)], dim=1)


def svr(sd: SD, pre: str, x: torch.Tensor, cfg) -> torch.Tensor:
    """SpatioTemporalVisualTokenRefinerModel.forward (svr.py:164-188)."""
    for i in range(cfg.u2t_num_layers):
        x = svr_layer(sd, f"{pre}attention_network.layers.{i}.", x, cfg.u2t_num_heads, cfg.attn_type)
    if cfg.enable_diffts:
        x = token_selection_diff(sd, pre + "token_selection.", x)
    else:
        x = token_selection_hard(sd, pre + "token_selection.", x, cfg.u2t_top_k)
    if cfg.use_multi_scale:
        x = multi_scale_pool(sd, pre + "dynamic_pool.", x, cfg.enable_dmtp)
    return x


# ---------------------------------------------------------------------------------------------
# TTA: text-conditioned token aggregator (src/model/u2tokenizer/tta.py:93-139)
# ---------------------------------------------------------------------------------------------
def tta(sd: SD, pre: str, query: torch.Tensor, visual: torch.Tensor, text: torch.Tensor, cfg) -> torch.Tensor:
    h = cfg.u2t_num_heads
    e = query.shape[-1]
    for i in range(cfg.u2t_num_layers):
        lp = f"{pre}layers_vt.{i}."
        s = self_attn(sd, lp + "self_attention.", query, h, cfg.attn_type)
        s = F.layer_norm(query + s, (e,), sd[lp + "norm_self.weight"], sd[lp + "norm_self.bias"], 1e-5)
        v = cross_attn(sd, lp + "visual_cross_attention.", s, visual, h)
        v = F.layer_norm(s + v, (e,), sd[lp + "norm_cross_v.weight"], sd[lp + "norm_cross_v.bias"], 1e-5)
        t = cross_attn(sd, lp + "text_cross_attention.", v, text, h)
        query = F.layer_norm(v + t, (e,), sd[lp + "norm_cross_t.weight"], sd[lp + "norm_cross_t.bias"], 1e-5)
    return cross_attn(sd, pre + "layer_linagg.linear_aggregator.", query, visual, h, is_compress=True)


def u2tokenizer(sd: SD, pre: str, v_token: torch.Tensor, t_token: torch.Tensor, cfg) -> torch.Tensor:
    """u2Tokenizer.forward (src/model/u2tokenizer/u2Tokenizer.py:40-47)."""
    b = v_token.shape[0]
    q = sd[pre + "query_tokens"].expand(b, -1, -1)
    vis = svr(sd, pre + "svt_module.", v_token, cfg)
    return tta(sd, pre + "tta_module.", q, vis, t_token, cfg)


# ---------------------------------------------------------------------------------------------
# multimodal front + splice (src/model/u2_arch.py:96-122)
# ---------------------------------------------------------------------------------------------
def encode_images(sd: SD, images: torch.Tensor, cfg) -> torch.Tensor:
    """u2MetaForCausalLM.encode_images (u2_arch.py:96-99)."""
    f = vit3d_tower(sd, "model.vision_tower.", images, cfg)
    return spatial_pooling_projector(sd, "model.mm_projector.", f, cfg)


def visual_tokens(sd: SD, images: torch.Tensor, question_ids: torch.Tensor, cfg) -> torch.Tensor:
    """The image branch of prepare_inputs_for_multimodal (u2_arch.py:108-117)."""
    if cfg.enable_u2tokenizer:
        B, C, D, H, W = images.shape
        f = encode_images(sd, images.view(B * C, 1, D, H, W), cfg)
        v_tokens = f.view(B, C, f.shape[-2], f.shape[-1])
        t_tokens = F.embedding(question_ids, sd["model.embed_tokens.weight"])
        return u2tokenizer(sd, "model.u2tokenizer.", v_tokens, t_tokens, cfg)
    return encode_images(sd, images, cfg)


def multimodal_embeds(sd: SD, input_ids: torch.Tensor, images: torch.Tensor, question_ids: torch.Tensor, cfg) -> torch.Tensor:
    """prepare_inputs_for_multimodal (u2_arch.py:101-122): the visual tokens overwrite positions
    1..n_vis of the prompt embeddings (position 0 keeps its own embedding)."""
    vis = visual_tokens(sd, images, question_ids, cfg)
    emb = F.embedding(input_ids, sd["model.embed_tokens.weight"])
    return torch.cat((emb[:, :1], vis, emb[:, vis.shape[1] + 1:]), dim=1)


# ---------------------------------------------------------------------------------------------
# decoder (HF transformers Qwen3 / Llama; reached via super().forward, u2llama.py:76-87)
# ---------------------------------------------------------------------------------------------
def _rms(x: torch.Tensor, w: torch.Tensor, eps: float) -> torch.Tensor:
    return w * (x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps))


def rope_inv_freq(cfg) -> torch.Tensor:
    """Default RoPE, or the 'llama3' frequency rescaling used by Llama-3.2 configs
    (base_model_tokenizers/Llama-3.2-1B-Instruct/config.json:49-56)."""
    dh = cfg.head_dim
    inv = 1.0 / (cfg.rope_theta ** (torch.arange(0, dh, 2, dtype=torch.float32) / dh))
    rs = getattr(cfg, "rope_scaling", None)
    if rs and rs.get("rope_type", rs.get("type")) == "llama3":
        factor, lo, hi = rs["factor"], rs["low_freq_factor"], rs["high_freq_factor"]
        old = rs["original_max_position_embeddings"]
        wavelen = 2 * math.pi / inv
        inv_l = torch.where(wavelen > old / lo, inv / factor, inv)
        smooth = (old / wavelen - lo) / (hi - lo)
        smoothed = (1 - smooth) * inv_l / factor + smooth * inv_l
        is_med = ~(wavelen < old / hi) & ~(wavelen > old / lo)
        inv = torch.where(is_med, smoothed, inv_l)
    return inv


def decoder_forward(sd: SD, inputs_embeds: torch.Tensor, cfg, past=None, return_hidden: bool = False):
    """Qwen3 / Llama decoder stack + lm_head, eager fp32, causal, positions = past_len + arange.
    `past`: optional list of (k, v) per layer [B, Hkv, T, dh]; returns (logits, new_past)."""
    b, s, e = inputs_embeds.shape
    hq, hkv, dh = cfg.num_attention_heads, cfg.num_key_value_heads, cfg.head_dim
    eps = cfg.rms_norm_eps
    past_len = 0 if past is None else past[0][0].shape[2]
    dev = inputs_embeds.device
    pos = torch.arange(past_len, past_len + s, dtype=torch.float32, device=dev)
    fr = torch.outer(pos, rope_inv_freq(cfg).to(dev))
    emb = torch.cat((fr, fr), dim=-1)
    cos, sin = emb.cos()[None, None].to(inputs_embeds.dtype), emb.sin()[None, None].to(inputs_embeds.dtype)
    x = inputs_embeds
    new_past = []
    total = past_len + s
    mask = torch.full((s, total), float("-inf"), device=dev, dtype=inputs_embeds.dtype).triu(diagonal=past_len + 1)
    for i in range(cfg.num_hidden_layers):
        lp = f"model.layers.{i}."
        y = _rms(x, sd[lp + "input_layernorm.weight"], eps)
        q = _lin(y, sd, lp + "self_attn.q_proj").view(b, s, hq, dh)
        k = _lin(y, sd, lp + "self_attn.k_proj").view(b, s, hkv, dh)
        v = _lin(y, sd, lp + "self_attn.v_proj").view(b, s, hkv, dh)
        if cfg.qk_norm:
            q = _rms(q, sd[lp + "self_attn.q_norm.weight"], eps)
            k = _rms(k, sd[lp + "self_attn.k_norm.weight"], eps)
        q, k, v = q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2)
        q = q * cos + _rotate_half(q) * sin
        k = k * cos + _rotate_half(k) * sin
        if past is not None:
            k = torch.cat((past[i][0], k), dim=2)
            v = torch.cat((past[i][1], v), dim=2)
        new_past.append((k, v))
        if USE_SDPA:
            o = _sdpa(q, k, v, 1.0 / math.sqrt(dh), causal=(s > 1)).transpose(1, 2).reshape(b, s, hq * dh)
        else:
            kk = k.repeat_interleave(hq // hkv, dim=1)
            vv = v.repeat_interleave(hq // hkv, dim=1)
            att = torch.softmax(q @ kk.transpose(-2, -1) / math.sqrt(dh) + mask, dim=-1)
            o = (att @ vv).transpose(1, 2).reshape(b, s, hq * dh)
        x = x + _lin(o, sd, lp + "self_attn.o_proj")
        y = _rms(x, sd[lp + "post_attention_layernorm.weight"], eps)
        y = _lin(F.silu(_lin(y, sd, lp + "mlp.gate_proj")) * _lin(y, sd, lp + "mlp.up_proj"), sd, lp + "mlp.down_proj")
        x = x + y
    x = _rms(x, sd["model.norm.weight"], eps)
    if return_hidden:
        return x, new_past
    w_head = sd["lm_head.weight"] if "lm_head.weight" in sd else sd["model.embed_tokens.weight"]
    return F.linear(x, w_head), new_past


def forward_logits(sd: SD, input_ids, images, question_ids, cfg) -> torch.Tensor:
    """u2*ForCausalLM.forward with images (u2llama.py:41-87): logits for every prompt position."""
    emb = multimodal_embeds(sd, input_ids, images, question_ids, cfg)
    return decoder_forward(sd, emb, cfg)[0]


@torch.no_grad()
def greedy_generate(sd: SD, input_ids, images, question_ids, cfg, max_new_tokens: int, eos_token_id=None):
    """u2*ForCausalLM.generate(do_sample=False) (u2llama.py:90-127): the vision path runs once, the
    decoder prefills on inputs_embeds, then one token per step with a KV cache; returns the NEW
    token ids only, plus the per-step top-2 logit margin (for margin-aware id comparison)."""
    emb = multimodal_embeds(sd, input_ids, images, question_ids, cfg)
    logits, past = decoder_forward(sd, emb, cfg)
    out, margins = [], []
    b = emb.shape[0]
    done = torch.zeros(b, dtype=torch.bool, device=emb.device)
    for _ in range(max_new_tokens):
        last = logits[:, -1]
        top2 = last.topk(2, dim=-1).values
        margins.append(top2[:, 0] - top2[:, 1])
        nxt = last.argmax(-1)
        out.append(nxt)
        if eos_token_id is not None:
            done |= nxt == eos_token_id
            if bool(done.all()):
                break
        logits, past = decoder_forward(sd, F.embedding(nxt[:, None], sd["model.embed_tokens.weight"]), cfg, past)
    return torch.stack(out, dim=1), torch.stack(margins, dim=1)


# ------------------------------------------------------------------------------------------------
# log-probability head (training / DPO evaluation side of the decoder output)
# ------------------------------------------------------------------------------------------------
def selective_log_softmax(logits: torch.Tensor, index: torch.Tensor) -> torch.Tensor:
    """`trl.trainer.utils.selective_log_softmax` (third-party trl, pinned `trl==0.9.6` in the reference's
    requirements.txt:138 although the function only exists in later trl releases; not installed here): the published
    definition is log_softmax(logits)[index] computed as gather(logits, index) - logsumexp(logits) in fp32.
    Call site: src/train/dpo_u2trainer.py:296. parity unpinned (no reference test or golden vector holds it)."""
    logits = logits.float()
    return logits.gather(-1, index.unsqueeze(-1)).squeeze(-1) - torch.logsumexp(logits, dim=-1)


def dpo_per_token_logps(logits: torch.Tensor, input_ids: torch.Tensor, loss_mask: torch.Tensor):
    """u2DPOTrainer.concatenated_forward, non-padding-free branch (src/train/dpo_u2trainer.py:274-302, 343-350):
    labels = input_ids rolled left by one, masked positions use the dummy label 0 and contribute 0, the result is rolled
    back right by one. Returns (per_token_logps [B, L], all_logps [B], mean of the masked rows' logits)."""
    labels = torch.roll(input_ids, shifts=-1, dims=1)
    mask = torch.roll(loss_mask, shifts=-1, dims=1).bool()
    if logits.shape[:2] != labels.shape[:2]:
        logits = logits[:, -labels.shape[1]:]
    labels = labels.clone()
    labels[~mask] = 0
    ptl = selective_log_softmax(logits, labels)
    ptl[~mask] = 0
    ptl = torch.roll(ptl, shifts=1, dims=1)
    return ptl, ptl.sum(-1), logits.float()[mask].mean()


def causal_lm_loss(logits: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
    """HF ForCausalLMLoss behind forward(labels=...) (u2llama.py:76-87): shift by one, mean NLL over labels != -100."""
    sl = logits[:, :-1].reshape(-1, logits.shape[-1]).float()
    tl = labels[:, 1:].reshape(-1)
    return F.cross_entropy(sl, tl, ignore_index=-100)
