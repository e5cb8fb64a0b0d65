"""Fused lm_head + selective log-softmax: the kernel against an fp32 restatement on the same bf16
operands, and the DPO per-token log-probability surface against the oracle (oracle/u2_oracle.py:dpo_per_token_logps)."""
import pytest
import torch

from oracle import u2_oracle as O
from u2tokenizer_b200.synthetic import synthetic_inputs

pytestmark = pytest.mark.gpu


def ref_head(h, w, labels):
    logits = h.float() @ w.float().t()
    lse = torch.logsumexp(logits, -1)
    safe = labels.clamp(min=0)
    lp = logits.gather(-1, safe[:, None])[:, 0] - lse
    return torch.where(labels >= 0, lp, torch.zeros_like(lp)), lse, logits.sum(-1)


@pytest.mark.parametrize("R,V,E", [(128, 512, 64), (200, 1000, 128), (1, 257, 64), (300, 151936, 256), (2048, 4096, 512)])
def test_lmhead_logprob_kernel(R, V, E):
    from u2tokenizer_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(R * 7 + V)
    h = (torch.randn(R, E, device="cuda", generator=g) * 0.7).bfloat16()
    w = (torch.randn(V, E, device="cuda", generator=g) * (2.0 / E ** 0.5)).bfloat16()
    labels = torch.randint(0, V, (R,), device="cuda", generator=g)
    labels[::7] = -1                      # ignored rows
    labels[-1] = V - 1                    # last (possibly partial) tile
    if R > 2:
        labels[1], labels[2] = 0, min(V - 1, 255)
    acc = torch.zeros(2, device="cuda")
    lp, lse, lsum = ops.lmhead_logprob(h, w, labels, want_lse=True, want_logit_sum=True, nll_acc=acc)
    rlp, rlse, rsum = ref_head(h, w, labels)
    assert (lse - rlse).abs().max().item() < 2e-3
    assert (lp - rlp).abs().max().item() < 2e-3
    assert (lsum - rsum).abs().max().item() < 2e-3 * max(1.0, rsum.abs().max().item())
    assert torch.all(lp[labels < 0] == 0)
    n = int((labels >= 0).sum())
    assert abs(acc[1].item() - n) < 0.5
    assert abs(acc[0].item() + rlp.sum().item()) < 2e-3 * max(1.0, abs(rlp.sum().item()))
    # strided hidden rows (a [:, :-1] slice of a [B, L, E] tensor) go through ldh
    if R >= 8:
        big = torch.zeros(R, 2 * E, device="cuda", dtype=torch.bfloat16)
        big[:, :E] = h
        lp2, _, _ = ops.lmhead_logprob(big[:, :E], w, labels)
        assert torch.equal(lp2, lp)


def test_lmhead_logprob_rejects_bad_arguments():
    from u2tokenizer_b200 import ops
    h = torch.zeros(4, 64, device="cuda", dtype=torch.bfloat16)
    w = torch.zeros(32, 64, device="cuda", dtype=torch.bfloat16)
    lab = torch.zeros(4, device="cuda", dtype=torch.int64)
    with pytest.raises(TypeError):
        ops.lmhead_logprob(h.float(), w, lab)
    with pytest.raises(ValueError):
        ops.lmhead_logprob(h, w, lab[:3])
    with pytest.raises(ValueError):
        ops.lmhead_logprob(h, w, lab, ws=torch.empty(16, device="cuda", dtype=torch.uint8))
    with pytest.raises(RuntimeError):
        ops.lmhead_logprob(h.cpu(), w, lab)


@pytest.mark.parametrize("family", ["qwen3", "llama"])
def test_dpo_per_token_logps_surface(family):
    from test_modeling_gpu import make
    model, g, sd = make(family)
    images, ids, qids = synthetic_inputs(g, batch=2, frames=2, n_question=6, lt=12)
    loss_mask = torch.zeros_like(ids)
    loss_mask[:, 10:] = 1                 # completion tokens only, as the DPO collator builds it
    loss_mask[1, -1] = 0                  # a padded tail position
    with torch.no_grad():
        ref_logits = O.forward_logits(sd, ids, images, qids, g)
        ref_ptl, ref_all, ref_mean = O.dpo_per_token_logps(ref_logits, ids, loss_mask)
    out = model.per_token_logps(images=images.cuda(), input_ids=ids.cuda(), question_ids=qids.cuda(), loss_mask=loss_mask.cuda())
    logits = model(images=images.cuda(), input_ids=ids.cuda(), question_ids=qids.cuda()).logits.float().cpu()
    # |d logp| <= 2 * sup|d logits| (log-softmax is 1-Lipschitz in the sup norm, twice: gather and logsumexp)
    bound = 2.0 * (logits - ref_logits).abs().max().item() + 1e-3
    ptl = out["per_token_logps"].cpu()
    assert ptl.shape == ref_ptl.shape
    assert (ptl - ref_ptl).abs().max().item() < bound
    assert torch.equal(ptl == 0, ref_ptl == 0)          # the same positions are masked out
    assert (out["all_logps"].cpu() - ref_all).abs().max().item() < bound * int(loss_mask.sum(1).max())
    assert abs(out["mean_logits"].item() - ref_mean.item()) < bound
    # and against the engine's own materialised logits: only summation order differs
    own_ptl, _, _ = O.dpo_per_token_logps(logits, ids, loss_mask)
    assert (ptl - own_ptl).abs().max().item() < 5e-3
