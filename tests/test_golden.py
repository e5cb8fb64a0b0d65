"""Oracle vs. committed golden outputs of the REFERENCE modules (tests/golden, made by
tools/make_golden.py). CPU-only; this is the pin that travels to machines without /root/reference."""
import os
import sys

import pytest
import torch

from common import fp32_sd, rel_err
from oracle import u2_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from make_golden import golden_geometry  # noqa: E402

GOLD = torch.load(os.path.join(ROOT, "tests", "golden", "u2_reference_outputs.pt"))


def golden_u2tok_inputs(g):
    gen = torch.Generator().manual_seed(21)
    v = torch.randn(2, 3, g.tokens_per_frame, g.hidden_size, generator=gen).bfloat16().float()
    t = torch.randn(2, 24, g.hidden_size, generator=gen).bfloat16().float()
    return v, t


@pytest.mark.parametrize("attn_type,diffts,dmtp", [("rma", True, True), ("rope", True, True), ("rma", False, False),
                                                     ("mha", True, True)])
def test_oracle_u2tokenizer_vs_golden(attn_type, diffts, dmtp):
    g = golden_geometry()
    g.attn_type, g.enable_diffts, g.enable_dmtp = attn_type, diffts, dmtp
    sd = fp32_sd(g, seed=11)
    v, t = golden_u2tok_inputs(g)
    with torch.no_grad():
        got = O.u2tokenizer(sd, "model.u2tokenizer.", v, t, g)
    assert rel_err(got, GOLD[f"u2tok_{attn_type}_{int(diffts)}{int(dmtp)}"]) < 2e-5


def test_oracle_spp_vs_golden():
    g = golden_geometry()
    sd = fp32_sd(g, seed=11)
    gen = torch.Generator().manual_seed(22)
    x = torch.randn(3, g.n_patches, g.vit_hidden, generator=gen).bfloat16().float()
    with torch.no_grad():
        got = O.spatial_pooling_projector(sd, "model.mm_projector.", x, g)
    assert rel_err(got, GOLD["spp"]) < 2e-5
