"""Shared test helpers: small geometries, fp32 state dicts, comparison metrics."""
import torch

from u2tokenizer_b200.geometry import Geometry
from u2tokenizer_b200.synthetic import synthetic_inputs, synthetic_state_dict


def tiny_geometry(**over) -> Geometry:
    """A scaled-down geometry that keeps every structural feature of the canonical one
    (3-D patches, cls token, 2x2x2 pooling, 8 heads, multi-scale, GQA decoder)."""
    kw = dict(
        image_size=[16, 64, 64], patch_size=[4, 16, 16], vit_hidden=96, vit_mlp=192, vit_layers=2, vit_heads=4,
        u2t_num_heads=8, u2t_num_layers=2, u2t_top_k=8, num_3d_query_token=8,
        hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=4,
        num_key_value_heads=2, head_dim=32, vocab_size=512, rms_norm_eps=1e-6, rope_theta=1e6,
        qk_norm=True, tie_word_embeddings=False,
    )
    kw.update(over)
    return Geometry(**kw)


def fp32_sd(g, seed=0):
    """bf16-rounded synthetic weights, held in fp32 (what the oracle computes with)."""
    return {k: v.float() for k, v in synthetic_state_dict(g, seed=seed, device="cpu", dtype=torch.bfloat16).items()}


def rel_err(out: torch.Tensor, ref: torch.Tensor) -> float:
    return ((out.float() - ref.float()).abs().max() / (ref.float().abs().max() + 1e-12)).item()


def cosine(out: torch.Tensor, ref: torch.Tensor) -> float:
    a, b = out.float().flatten(), ref.float().flatten()
    return (a @ b / (a.norm() * b.norm() + 1e-12)).item()
