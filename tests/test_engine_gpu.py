"""End-to-end and per-stage parity of the CUDA path (U2Engine over libu2b200.so) against the fp32
oracle on identical seeded weights / volumes / prompts.

Tolerance (stated, bf16): per stage  max|out - ref| / max|ref| <= 3e-2 and cosine >= 0.999
(activations and weights are bf16 on the CUDA side, the oracle computes in fp32 on the same
bf16-rounded weights). Greedy token ids must be exact wherever the oracle's top-1/top-2 logit
margin exceeds 2e-2 * max|logit| (margin-aware)."""
import os
import sys

import pytest
import torch

from common import cosine, fp32_sd, rel_err, tiny_geometry
from oracle import u2_oracle as O
from u2tokenizer_b200.synthetic import synthetic_inputs, synthetic_state_dict

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
TOL, COS = 3e-2, 0.999


def build(g, seed, **head_kw):
    from u2tokenizer_b200.engine import U2Engine
    sd16 = synthetic_state_dict(g, seed=seed, device="cpu", dtype=torch.bfloat16, **head_kw)
    eng = U2Engine(g, sd16, device="cuda")
    return eng, {k: v.float() for k, v in sd16.items()}


def check(out, ref, tol=TOL, what=""):
    out = out.float().cpu()
    e, c = rel_err(out, ref), cosine(out, ref)
    assert e < tol and c > COS, f"{what}: rel_err {e:.4g} cosine {c:.6f}"


@pytest.mark.parametrize("ptype", ["spatial", "sequence"])
def test_encode_images(ptype):
    g = tiny_geometry(proj_pooling_type=ptype)
    eng, sd = build(g, 1)
    imgs = torch.rand(3, 1, *g.image_size, generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        ref = O.encode_images(sd, imgs, g)
    check(eng.encode_images(imgs.cuda()), ref, what="encode_images")


@pytest.mark.parametrize("attn_type,dmtp,multi", [("rma", True, True), ("rope", True, True), ("rma", False, True),
                                                   ("rma", False, False), ("mha", True, True)])
def test_u2tokenizer(attn_type, dmtp, multi):
    g = tiny_geometry(attn_type=attn_type, enable_dmtp=dmtp, use_multi_scale=multi)
    eng, sd = build(g, 2)
    gen = torch.Generator().manual_seed(1)
    v = torch.randn(2, 3, g.tokens_per_frame, g.hidden_size, generator=gen).bfloat16()
    t = torch.randn(2, 5, g.hidden_size, generator=gen).bfloat16()
    with torch.no_grad():
        ref = O.u2tokenizer(sd, "model.u2tokenizer.", v.float(), t.float(), g)
    check(eng.u2tokenizer(v.cuda(), t.cuda()), ref, what="u2tokenizer")


@pytest.mark.parametrize("attn_type", ["rma", "rope", "mha"])  # "mha" = the nn.MultiheadAttention fallback
def test_u2tokenizer_vs_reference_golden(attn_type):
    """CUDA path against the committed outputs of the REFERENCE u2Tokenizer (tests/golden)."""
    from make_golden import golden_geometry
    gold = torch.load(os.path.join(ROOT, "tests", "golden", "u2_reference_outputs.pt"))
    g = golden_geometry()
    g.attn_type = attn_type
    eng, sd = build(g, 11)
    gen = torch.Generator().manual_seed(21)
    v = torch.randn(2, 3, g.tokens_per_frame, g.hidden_size, generator=gen).bfloat16()
    t = torch.randn(2, 24, g.hidden_size, generator=gen).bfloat16()
    check(eng.u2tokenizer(v.cuda(), t.cuda()), gold[f"u2tok_{attn_type}_11"], what="golden u2tokenizer")


@pytest.mark.parametrize("B,C,E", [(1, 4, 128), (3, 2, 512), (1, 1, 128)])
def test_u2tokenizer_mha_fallback_shapes(B, C, E):
    """attn_type outside {"rma", "rope"} -> torch.nn.MultiheadAttention called sequence-first (reference svr.py:17-18,
    tta.py:83-84): one study (no regrouping copy), three studies (frame-major regrouping), head_dim 64 (the fused
    attention kernel on transposed views) and the degenerate single-frame / single-sample sequences of length 1."""
    g = tiny_geometry(attn_type="mha", hidden_size=E, intermediate_size=2 * E, head_dim=E // 4)
    eng, sd = build(g, 5)
    gen = torch.Generator().manual_seed(2)
    v = torch.randn(B, C, g.tokens_per_frame, E, generator=gen).bfloat16()
    t = torch.randn(B, 7, E, generator=gen).bfloat16()
    with torch.no_grad():
        ref = O.u2tokenizer(sd, "model.u2tokenizer.", v.float(), t.float(), g)
    check(eng.u2tokenizer(v.cuda(), t.cuda()), ref, what=f"mha fallback B={B} C={C} E={E}")


@pytest.mark.parametrize("family", ["qwen3", "llama"])
def test_forward_logits_and_generate(family):
    if family == "qwen3":
        g = tiny_geometry()
    else:
        rs = dict(factor=8.0, high_freq_factor=4.0, low_freq_factor=1.0, original_max_position_embeddings=16,
                  rope_type="llama3")
        g = tiny_geometry(qk_norm=False, rope_theta=500000.0, rope_scaling=rs, tie_word_embeddings=True, head_dim=32)
    # untied head (qwen3 variant): bigram-structured weights give decisive top-1 / top-2 margins, so (nearly) every
    # greedy token is actually compared; the tied Llama variant can only use the log-normal row-norm profile
    eng, sd = build(g, 3, **(dict(bigram=1.0) if family == "qwen3" else dict(head_tail=1.0)))
    images, ids, qids = synthetic_inputs(g, batch=2, frames=2, n_question=6, lt=12)
    n_new = 24
    with torch.no_grad():
        ref_emb = O.multimodal_embeds(sd, ids, images, qids, g)
        ref_logits = O.decoder_forward(sd, ref_emb, g)[0]
        ref_ids, margins = O.greedy_generate(sd, ids, images, qids, g, max_new_tokens=n_new)
    emb = eng.multimodal_embeds(ids.cuda(), images.cuda(), qids.cuda())
    check(emb, ref_emb, what="multimodal_embeds")
    hidden = eng.prefill(emb)
    logits = eng.lm_logits(hidden)
    check(logits, ref_logits, what="prefill logits")
    # an argmax flip needs a top-1/top-2 margin below twice the logit error: derive the threshold from the
    # error actually measured on the prefill logits (x4 headroom for the cached decode steps)
    thr = 4.0 * (logits.float().cpu() - ref_logits).abs().max().item()
    assert thr < 0.15 * ref_logits.abs().max().item()
    for use_graph, impl in ((False, "tcgen05"), (True, "tcgen05"), (False, "gemv"), (True, "gemv")):
        eng.decode_impl = impl
        got = eng.generate_greedy(emb, max_new_tokens=n_new, use_graph=use_graph).cpu()
        # compare up to (excluding) the first low-margin step of each sequence: after it the oracle's own
        # choice is not robust to bf16 rounding and the continuations legitimately diverge
        compared = 0
        for b in range(got.shape[0]):
            low = (margins[b] < thr).nonzero()
            upto = int(low[0]) if len(low) else got.shape[1]
            compared += upto
            assert torch.equal(got[b, :upto], ref_ids[b, :upto]), (use_graph, b, got[b], ref_ids[b], margins[b])
        assert got.shape == ref_ids.shape
        print(f"[{family} {impl} graph={use_graph}] greedy ids identical on {compared}/{got.numel()} compared tokens")
        if family == "qwen3":
            assert compared >= 0.9 * got.numel(), (compared, got.numel(), margins, thr)


@pytest.mark.parametrize("impl", ["tcgen05", "gemv"])
def test_decode_matches_prefill(impl):
    """KV-cached decode steps reproduce teacher-forced prefill logits (size-independent property)."""
    g = tiny_geometry()
    eng, sd = build(g, 4)
    eng.decode_impl = impl
    gen = torch.Generator().manual_seed(5)
    emb = (torch.randn(2, 20, g.hidden_size, generator=gen) * 0.5).bfloat16().cuda()
    full = eng.lm_logits(eng.prefill(emb))
    cache = eng.new_cache(2, 32)
    h = eng.prefill(emb[:, :12].contiguous(), cache)
    bufs = eng._decode_buffers(2)
    # feed the remaining embeddings one at a time through the decode path
    from u2tokenizer_b200 import ops
    for t in range(12, 20):
        # inject the embedding directly: emulate ids -> embed by writing x after the gather
        saved = eng.embed
        eng.embed = emb[:, t].contiguous()  # table of 2 rows; ids 0,1 select them
        bufs["ids"].copy_(torch.arange(2, device="cuda").view(2, 1))
        lg = eng.decode_step(cache)
        eng.embed = saved
        e = rel_err(lg.cpu(), full[:, t].float().cpu())
        assert e < 3e-2, (t, e)


def test_full_size_shapes_one_layer():
    """Canonical widths (E=2048: head_dim 256 attention, S=2049 ViT sequence, 1024-way DiffTS, 1792-token
    cross attention) with the depth cut to one layer per stack so the fp32 oracle finishes in seconds."""
    g = tiny_geometry(image_size=[32, 256, 256], patch_size=[4, 16, 16], vit_hidden=768, vit_mlp=3072, vit_layers=1,
                      vit_heads=12, u2t_num_layers=1, u2t_top_k=1024, num_3d_query_token=256, hidden_size=2048,
                      intermediate_size=6144, num_hidden_layers=1, num_attention_heads=16, num_key_value_heads=8,
                      head_dim=128, vocab_size=4096)
    eng, sd = build(g, 7)
    images, ids, qids = synthetic_inputs(g, batch=1, frames=2, n_question=16, lt=64)
    with torch.no_grad():
        feats = O.encode_images(sd, images.view(2, 1, *g.image_size), g)
        ref_vis = O.u2tokenizer(sd, "model.u2tokenizer.", feats.view(1, 2, -1, g.hidden_size),
                                torch.nn.functional.embedding(qids, sd["model.embed_tokens.weight"]), g)
    got_feats = eng.encode_images(images.view(2, 1, *g.image_size).cuda())
    check(got_feats, feats, what="encode_images full width")
    vis = eng.visual_tokens(images.cuda(), qids.cuda())
    check(vis, ref_vis, what="visual tokens full width")


def test_u2tokenizer_hard_selection():
    """enable_diffts=False (reference svr.py:75-91): the top-k SET must agree with the oracle except for
    candidates whose scores are closer than the bf16 noise; the pipeline output is compared on that basis."""
    g = tiny_geometry(enable_diffts=False, enable_dmtp=False, u2t_top_k=8)
    eng, sd = build(g, 8)
    gen = torch.Generator().manual_seed(3)
    v = torch.randn(2, 3, g.tokens_per_frame, g.hidden_size, generator=gen).bfloat16()
    t = torch.randn(2, 5, g.hidden_size, generator=gen).bfloat16()
    with torch.no_grad():
        ref = O.u2tokenizer(sd, "model.u2tokenizer.", v.float(), t.float(), g)
        x = v.float()
        for i in range(g.u2t_num_layers):
            x = O.svr_layer(sd, f"model.u2tokenizer.svt_module.attention_network.layers.{i}.", x, g.u2t_num_heads, g.attn_type)
        scores = (x @ sd["model.u2tokenizer.svt_module.token_selection.score_net.weight"].t()).view(2, -1)
        ref_idx = scores.topk(g.u2t_top_k, dim=1).indices
    got = eng.u2tokenizer(v.cuda(), t.cuda()).float().cpu()
    sel = eng.last_selection.cpu() - torch.arange(2)[:, None] * scores.shape[1]
    # every selected token must be within bf16 noise of the oracle's k-th score
    kth = scores.topk(g.u2t_top_k, dim=1).values[:, -1:]
    picked = torch.gather(scores, 1, sel)
    assert (picked >= kth - 3e-2 * scores.abs().max()).all()
    # downstream of the selection, unconditionally: the oracle continues from the ENGINE's selection (a near-tie at the
    # k-th score may legitimately pick a different token; everything after the pick must still match)
    with torch.no_grad():
        x_sel = x.view(2, -1, g.hidden_size)[torch.arange(2)[:, None], sel]
        vis = O.multi_scale_pool(sd, "model.u2tokenizer.svt_module.dynamic_pool.", x_sel, g.enable_dmtp) \
            if g.use_multi_scale else x_sel
        q = sd["model.u2tokenizer.query_tokens"].expand(2, -1, -1)
        ref_from_sel = O.tta(sd, "model.u2tokenizer.tta_module.", q, vis, t.float(), g)
    check(got, ref_from_sel, what="hard-selection tokenizer (oracle continued from the engine's selection)")
    n_same = int((sel == ref_idx).sum())
    print(f"hard selection: {n_same}/{sel.numel()} indices identical to torch.topk on the fp32 oracle scores")
    if torch.equal(sel, ref_idx):
        check(got, ref, what="hard-selection tokenizer")


@pytest.mark.parametrize("diffts,dmtp", [(True, True), (False, False)])
def test_reference_smoke_shape_svr(diffts, dmtp):
    """The reference's own SVR smoke run (src/model/u2tokenizer/svr.py:190-205): attn_type "rope", E = 512, 8 heads,
    4 layers, top_k 1024, multi-scale, input [1, 64, 256, 512] -> (1, 1792, 512) - one of the only two "known answers"
    the reference holds. 64 frames -> temporal attention over 64 positions; (False, False) is the
    smoke block's own configuration (hard top-k over 16384 tokens, plain multi-scale pooling), (True, True) the
    canonical one (a 16384-token DiffTS softmax). The rest of the tokenizer runs on top with a short question."""
    g = tiny_geometry(hidden_size=512, attn_type="rope", u2t_num_heads=8, u2t_num_layers=4, u2t_top_k=1024,
                      use_multi_scale=True, num_3d_query_token=8, enable_diffts=diffts, enable_dmtp=dmtp)
    eng, sd = build(g, 11)
    gen = torch.Generator().manual_seed(5)
    v = torch.randn(1, 64, 256, 512, generator=gen).bfloat16()
    t = torch.randn(1, 6, 512, generator=gen).bfloat16()
    with torch.no_grad():
        vis = O.svr(sd, "model.u2tokenizer.svt_module.", v.float(), g)
        assert tuple(vis.shape) == (1, 1792, 512)          # the reference's printed shape
        ref = O.u2tokenizer(sd, "model.u2tokenizer.", v.float(), t.float(), g)
    check(eng.u2tokenizer(v.cuda(), t.cuda()), ref, what="svr smoke shape")


def test_reference_smoke_shape_tta():
    """The reference's own TTA smoke run (src/model/u2tokenizer/tta.py:142-151): attn_type "rope", E = 896 (head_dim
    112), 4 layers, query [1, 64, 896], visual [1, 1792, 896], text [1, 755, 896] -> (1, 64, 896)."""
    g = tiny_geometry(hidden_size=896, attn_type="rope", u2t_num_heads=8, u2t_num_layers=4, u2t_top_k=1024,
                      use_multi_scale=True, num_3d_query_token=64)
    eng, sd = build(g, 12)
    gen = torch.Generator().manual_seed(6)
    v = torch.randn(1, 2, 24, 896, generator=gen).bfloat16()   # DiffTS makes 1024 -> multi-scale 1792 visual tokens
    t = torch.randn(1, 755, 896, generator=gen).bfloat16()
    with torch.no_grad():
        assert tuple(O.svr(sd, "model.u2tokenizer.svt_module.", v.float(), g).shape) == (1, 1792, 896)
        ref = O.u2tokenizer(sd, "model.u2tokenizer.", v.float(), t.float(), g)
        assert tuple(ref.shape) == (1, 64, 896)              # the reference's printed shape
    check(eng.u2tokenizer(v.cuda(), t.cuda()), ref, what="tta smoke shape")


def test_cfg1_single_32cube_volume():
    """BASELINE.json configs[0] / SURVEY.md section 8d "cfg 1": ONE 32 x 32 x 32 volume -> images [1, 1, 32, 32, 32],
    image_size (32, 32, 32) -> 32 patches -> 4 tokens per chunk, u2t_top_k = 4, num_3d_query_token = 4, a 2-layer
    Qwen3-shaped stub: the smallest shape the reference path accepts (TokenSelection needs top_k <= C * tokens)."""
    for diffts in (True, False):
        g = tiny_geometry(image_size=[32, 32, 32], u2t_top_k=4, num_3d_query_token=4, enable_diffts=diffts)
        assert g.n_patches == 32 and g.tokens_per_frame == 4
        eng, sd = build(g, 13)
        images, ids, qids = synthetic_inputs(g, batch=1, frames=1, n_question=5, lt=8)
        assert tuple(images.shape) == (1, 1, 32, 32, 32)
        with torch.no_grad():
            ref_emb = O.multimodal_embeds(sd, ids, images, qids, g)
            ref_logits = O.decoder_forward(sd, ref_emb, g)[0]
            ref_ids, margins = O.greedy_generate(sd, ids, images, qids, g, max_new_tokens=6)
        emb = eng.multimodal_embeds(ids.cuda(), images.cuda(), qids.cuda())
        check(emb, ref_emb, what="cfg1 multimodal_embeds")
        logits = eng.lm_logits(eng.prefill(emb))
        check(logits, ref_logits, what="cfg1 logits")
        thr = 4.0 * (logits.float().cpu() - ref_logits).abs().max().item()
        got = eng.generate_greedy(emb, max_new_tokens=6).cpu()
        low = (margins[0] < thr).nonzero()
        upto = int(low[0]) if len(low) else got.shape[1]
        assert torch.equal(got[0, :upto], ref_ids[0, :upto]), (got, ref_ids, margins)
