import sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from test_modeling_gpu import make
from oracle import u2_oracle as O
from u2tokenizer_b200.synthetic import synthetic_inputs
import torch.nn.functional as F
for fam in ("llama", "qwen3"):
    model, g, sd = make(fam)
    images, ids, qids = synthetic_inputs(g, batch=2, frames=2, n_question=6, lt=12)
    eng = model.engine()
    with torch.no_grad():
        emb_ref = O.multimodal_embeds(sd, ids, images, qids, g)
        lg0, past = O.decoder_forward(sd, emb_ref, g)
        nxt = lg0[:, -1].argmax(-1)
        lg1, _ = O.decoder_forward(sd, F.embedding(nxt[:, None], sd["model.embed_tokens.weight"]), g, past)
    emb = eng.multimodal_embeds(ids.cuda(), images.cuda(), qids.cuda())
    for impl in ("tcgen05", "gemv"):
        eng.decode_impl = impl
        cache = eng.new_cache(2, emb.shape[1] + 4)
        h = eng.prefill(emb, cache)
        l0 = eng.lm_logits(h[:, -1].contiguous()).float().cpu()
        bufs = eng._decode_buffers(2)
        bufs["ids"].copy_(nxt.view(2, 1).cuda())
        l1 = eng.decode_step(cache).float().cpu()
        print(fam, impl, "prefill err", (l0 - lg0[:, -1]).abs().max().item(), "decode err", (l1 - lg1[:, -1]).abs().max().item(),
              "scale", lg1.abs().max().item(), "argmax ours", l1.argmax(-1).tolist(), "ref", lg1[:, -1].argmax(-1).tolist())
