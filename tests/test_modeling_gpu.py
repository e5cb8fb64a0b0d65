"""The HF-style surface (U2Qwen3ForCausalLM / U2LlamaForCausalLM: forward, generate, the call forms the
reference's trainers and eval scripts use) against the oracle, through the same API a user calls."""
import pytest
import torch

from common import cosine, rel_err, tiny_geometry
from oracle import u2_oracle as O
from u2tokenizer_b200.synthetic import synthetic_inputs, synthetic_state_dict

pytestmark = pytest.mark.gpu


def make(family):
    from u2tokenizer_b200.configuration import U2LlamaConfig, U2Qwen3Config
    from u2tokenizer_b200.geometry import Geometry
    from u2tokenizer_b200.modeling import U2LlamaForCausalLM, U2Qwen3ForCausalLM
    kw = dict(hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2,
              head_dim=32, vocab_size=512, image_size=[16, 64, 64], vit_hidden_size=96, vit_mlp_dim=192, vit_num_layers=2,
              vit_num_heads=4, u2t_num_layers=2, u2t_top_k=8, num_3d_query_token=8, tie_word_embeddings=False,
              rms_norm_eps=1e-6)
    if family == "qwen3":
        cfg = U2Qwen3Config(**kw)
        model = U2Qwen3ForCausalLM(cfg)
    else:
        cfg = U2LlamaConfig(**kw)
        model = U2LlamaForCausalLM(cfg)
    g = Geometry.from_hf(cfg)
    sd16 = synthetic_state_dict(g, seed=9, device="cpu", dtype=torch.bfloat16)
    res = model.load_state_dict(sd16, strict=False)
    assert not res.unexpected_keys and all("rotary" in k or "inv_freq" in k for k in res.missing_keys)
    model = model.to(torch.bfloat16).cuda().eval()
    model.generation_config.eos_token_id = None  # the oracle's greedy loop has no EOS; tested separately below
    return model, g, {k: v.float() for k, v in sd16.items()}


@pytest.mark.parametrize("family", ["qwen3", "llama"])
def test_forward_and_generate_surface(family):
    model, g, sd = make(family)
    images, ids, qids = synthetic_inputs(g, batch=2, frames=2, n_question=6, lt=12)
    labels = ids.clone()
    labels[:, :9] = -100
    with torch.no_grad():
        ref_logits = O.forward_logits(sd, ids, images, qids, g)
        ref_ids, margins = O.greedy_generate(sd, ids, images, qids, g, max_new_tokens=6)
        ref_loss = torch.nn.functional.cross_entropy(ref_logits[:, :-1].reshape(-1, g.vocab_size), labels[:, 1:].reshape(-1),
                                                     ignore_index=-100)
    # trainer-style call (reference train_stage1.py:244-250)
    out = model(images=images.cuda(), input_ids=ids.cuda(), labels=labels.cuda(),
                attention_mask=torch.ones_like(ids).cuda(), question_ids=qids.cuda())
    lg = out.logits.float().cpu()
    assert rel_err(lg, ref_logits) < 3e-2 and cosine(lg, ref_logits) > 0.999
    assert abs(out.loss.item() - ref_loss.item()) < 3e-2 * max(1.0, ref_loss.item())
    thr = 4.0 * (lg - ref_logits).abs().max().item()  # margin below which an argmax flip is within bf16 error
    assert thr < 0.15 * ref_logits.abs().max().item()

    def same(got):
        got = got.cpu()
        assert got.shape == ref_ids.shape
        for b in range(got.shape[0]):
            low = (margins[b] < thr).nonzero()
            upto = int(low[0]) if len(low) else got.shape[1]
            assert torch.equal(got[b, :upto], ref_ids[b, :upto]), (got[b], ref_ids[b], margins[b])
    # eval-style calls (reference eval/mrg.py:74, green_refactored/lu2_model.py:63, dpo_u2trainer.py:71-79)
    same(model.generate(images.cuda(), ids.cuda(), question_ids=qids.cuda(), max_new_tokens=6, do_sample=False))
    same(model.generate(images.cuda(), ids.cuda(), qids.cuda(), max_new_tokens=6, do_sample=False))
    same(model.generate(images=images.cuda(), question_ids=qids.cuda(), input_ids=ids.cuda(),
                        attention_mask=torch.ones_like(ids).cuda(), max_length=ids.shape[1] + 6, do_sample=False))
    with pytest.raises(NotImplementedError):
        model.generate(images.cuda(), ids.cuda(), question_ids=qids.cuda(), inputs_embeds=torch.zeros(1))
    # text-only path (images=None): plain embedding lookup, reference u2llama.py:120-121
    with torch.no_grad():
        ref_txt = O.decoder_forward(sd, torch.nn.functional.embedding(ids, sd["model.embed_tokens.weight"]), g)[0]
    lg = model(input_ids=ids.cuda()).logits.float().cpu()
    assert rel_err(lg, ref_txt) < 3e-2


def test_eos_padding_semantics():
    model, g, sd = make("qwen3")
    images, ids, qids = synthetic_inputs(g, batch=2, frames=2, n_question=6, lt=12)
    free = model.generate(images.cuda(), ids.cuda(), question_ids=qids.cuda(), max_new_tokens=8, do_sample=False).cpu()
    eos = int(free[0, 2])
    got = model.generate(images.cuda(), ids.cuda(), question_ids=qids.cuda(), max_new_tokens=8, do_sample=False,
                         eos_token_id=eos, pad_token_id=0).cpu()
    first = (free[0] == eos).nonzero()[0].item()
    assert torch.equal(got[0, :first + 1], free[0, :first + 1])
    assert (got[0, first + 1:] == 0).all()


def test_sampled_generate_surface():
    """generate(do_sample=True, top_p, temperature) - the call every eval script of the reference makes
    (eval/mrg.py:74-75): runs on the CUDA sampling head, is reproducible for a fixed seed, varies across seeds, and
    collapses to greedy when the nucleus is a single token."""
    model, g, sd = make("qwen3")
    images, ids, qids = synthetic_inputs(g, batch=2, frames=2, n_question=6, lt=12)
    kw = dict(question_ids=qids.cuda(), max_new_tokens=8, do_sample=True, top_p=0.9, temperature=1.0)
    a = model.generate(images.cuda(), ids.cuda(), seed=11, **kw).cpu()
    b = model.generate(images.cuda(), ids.cuda(), seed=11, **kw).cpu()
    st = model.engine()._gen_state
    graph = st["graph"]
    c = model.generate(images.cuda(), ids.cuda(), seed=12, **kw).cpu()
    assert a.shape == (2, 8) and torch.equal(a, b) and not torch.equal(a, c)
    # seed / temperature / top-p live in a device block: a new request replays the SAME captured decode graph
    d = model.generate(images.cuda(), ids.cuda(), seed=12, **dict(kw, temperature=0.5, top_p=0.5)).cpu()
    assert graph is not None and model.engine()._gen_state is st and st["graph"] is graph
    assert d.shape == (2, 8)
    greedy = model.generate(images.cuda(), ids.cuda(), question_ids=qids.cuda(), max_new_tokens=8, do_sample=False).cpu()
    tiny_nucleus = model.generate(images.cuda(), ids.cuda(), question_ids=qids.cuda(), max_new_tokens=8, do_sample=True,
                                  top_p=1e-6, temperature=1.0, seed=5).cpu()
    assert torch.equal(tiny_nucleus, greedy)


def test_multi_sample_generate_shares_the_prefill():
    """generate(do_sample=True, num_return_sequences=n): one vision + prefill pass, the prompt's KV rows replicated into
    the decode cache (the reference's DPO-data workflow draws its 8 samples per study with separate full passes,
    green_refactored/pred_then_green.py:77-83). HF row order: row b*n + s is sample s of prompt b."""
    model, g, sd = make("qwen3")
    images, ids, qids = synthetic_inputs(g, batch=2, frames=2, n_question=6, lt=12)
    args = (images.cuda(), ids.cuda())
    greedy = model.generate(*args, question_ids=qids.cuda(), max_new_tokens=7, do_sample=False).cpu()
    # a single-token nucleus makes every sample the greedy continuation: checks the replicated KV rows exactly,
    # including the 16-sequence chunking of the decode batch (2 prompts x 9 samples = 16 + 2 rows)
    for n in (3, 9):
        out = model.generate(*args, question_ids=qids.cuda(), max_new_tokens=7, do_sample=True, top_p=1e-6,
                             num_return_sequences=n, seed=3).cpu()
        assert out.shape == (2 * n, 7)
        assert torch.equal(out, greedy.repeat_interleave(n, dim=0))
    kw = dict(question_ids=qids.cuda(), max_new_tokens=7, do_sample=True, top_p=0.98, temperature=2.0, num_return_sequences=4)
    a = model.generate(*args, seed=21, **kw).cpu()
    b = model.generate(*args, seed=21, **kw).cpu()
    c = model.generate(*args, seed=22, **kw).cpu()
    assert a.shape == (8, 7) and torch.equal(a, b) and not torch.equal(a, c)
    assert len({tuple(r.tolist()) for r in a[:4]}) > 1      # the samples of one prompt are distinct draws
    assert int(a.min()) >= 0 and int(a.max()) < g.vocab_size
    with pytest.raises(ValueError):
        model.generate(*args, question_ids=qids.cuda(), max_new_tokens=4, do_sample=False, num_return_sequences=2)


def test_forward_graph_replay_matches_eager():
    """Repeated same-shape forwards replay one CUDA graph over static buffers (engine.forward_logits): every replay must
    give what the eager launch sequence gives on the same inputs, for inputs that change from call to call."""
    model, g, sd = make("qwen3")
    eng = model.engine()
    for seed in (1, 2, 3, 4):
        images, ids, qids = synthetic_inputs(g, batch=2, frames=2, n_question=6, lt=12, seed=seed)
        got = model(images=images.cuda(), input_ids=ids.cuda(), question_ids=qids.cuda()).logits
        ref = eng.forward_logits(ids.cuda(), images.cuda(), qids.cuda(), use_graph=False)
        assert got.shape == ref.shape and rel_err(got, ref) < 1e-3, seed
    assert eng._fwd_state["graph"] is not None and eng._fwd_state["n"] > 50
    # a different shape drops the graph and starts over
    images, ids, qids = synthetic_inputs(g, batch=1, frames=2, n_question=6, lt=12, seed=9)
    a = model(images=images.cuda(), input_ids=ids.cuda(), question_ids=qids.cuda()).logits
    b = model(images=images.cuda(), input_ids=ids.cuda(), question_ids=qids.cuda()).logits
    assert a.shape[0] == 1 and rel_err(a, b) < 1e-3
