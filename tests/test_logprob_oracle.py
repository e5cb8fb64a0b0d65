"""The log-probability oracle (oracle/u2_oracle.py: selective_log_softmax, dpo_per_token_logps, causal_lm_loss) against
independent torch formulations on CPU. trl (the third-party home of selective_log_softmax) is not installed: the
restatement is checked against log_softmax + gather, and the roll / mask bookkeeping of
u2DPOTrainer.concatenated_forward (reference src/train/dpo_u2trainer.py:274-302) against a literal per-position loop."""
import torch

from oracle import u2_oracle as O


def test_selective_log_softmax_is_log_softmax_gather():
    g = torch.Generator().manual_seed(0)
    logits = torch.randn(3, 7, 50, generator=g) * 4
    idx = torch.randint(0, 50, (3, 7), generator=g)
    ref = torch.log_softmax(logits.double(), -1).gather(-1, idx[..., None])[..., 0]
    assert torch.allclose(O.selective_log_softmax(logits, idx).double(), ref, atol=1e-5)


def test_dpo_per_token_logps_matches_a_literal_loop():
    g = torch.Generator().manual_seed(1)
    B, L, V = 2, 9, 31
    logits = torch.randn(B, L, V, generator=g) * 3
    ids = torch.randint(1, V, (B, L), generator=g)
    mask = torch.zeros(B, L, dtype=torch.long)
    mask[0, 4:] = 1
    mask[1, 5:8] = 1
    ptl, all_lp, mean_logits = O.dpo_per_token_logps(logits, ids, mask)
    lsm = torch.log_softmax(logits.double(), -1)
    want = torch.zeros(B, L, dtype=torch.float64)
    rows = []
    for b in range(B):
        for t in range(L):
            # position t predicts token t + 1 (wrapping like torch.roll); it counts iff loss_mask[t + 1] is set;
            # the value is stored one step to the right (the final roll back)
            nxt = (t + 1) % L
            if mask[b, nxt]:
                want[b, nxt] = lsm[b, t, ids[b, nxt]]
                rows.append(logits[b, t])
    assert torch.allclose(ptl.double(), want, atol=1e-5)
    assert torch.allclose(all_lp.double(), want.sum(-1), atol=1e-5)
    assert abs(mean_logits.item() - torch.stack(rows).mean().item()) < 1e-6
    assert torch.equal(ptl == 0, want == 0)


def test_causal_lm_loss_matches_manual_shift():
    g = torch.Generator().manual_seed(2)
    logits = torch.randn(2, 6, 17, generator=g)
    labels = torch.randint(0, 17, (2, 6), generator=g)
    labels[:, :3] = -100
    lsm = torch.log_softmax(logits.double(), -1)
    vals = [-lsm[b, t, labels[b, t + 1]] for b in range(2) for t in range(5) if labels[b, t + 1] != -100]
    assert abs(O.causal_lm_loss(logits, labels).item() - torch.stack(vals).mean().item()) < 1e-6
