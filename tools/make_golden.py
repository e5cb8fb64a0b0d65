"""Generate tests/golden/*.pt from the REFERENCE modules (imported unmodified from /root/reference,
MONAI stand-in for the ViT blocks). Run in the authoring container only:

    python tools/make_golden.py

Inputs and weights are NOT stored: they are regenerated from seeds by
u2tokenizer_b200.synthetic (CPU generators are deterministic); only the reference OUTPUTS are
committed, as bf16/fp32 tensors of a few KB.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import refshim  # noqa: E402
from common import fp32_sd, tiny_geometry  # noqa: E402
from u2tokenizer_b200.synthetic import synthetic_inputs  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def golden_geometry():
    """Bigger than the pin-test geometry so that the CUDA kernels see legal tile shapes
    (E=256, 8 heads -> dh 32; 3 frames; 64 tokens/frame; top_k 64; 32 queries)."""
    return tiny_geometry(image_size=[16, 128, 128], patch_size=[4, 16, 16], vit_hidden=128, vit_mlp=256,
                         vit_layers=2, vit_heads=4, u2t_num_layers=2, u2t_top_k=64, num_3d_query_token=32,
                         hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=4,
                         num_key_value_heads=2, head_dim=64, vocab_size=1024)


def _sub(sd, prefix):
    return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}


def main():
    refshim.install()
    from src.model.u2tokenizer.u2Tokenizer import u2Tokenizer
    from src.model.multimodal_projector.spatial_pooling_projector import SpatialPoolingProjector
    os.makedirs(OUT, exist_ok=True)
    gold = {}
    for attn_type, diffts, dmtp in (("rma", True, True), ("rope", True, True), ("rma", False, False), ("mha", True, True)):
        g = golden_geometry()
        g.attn_type, g.enable_diffts, g.enable_dmtp = attn_type, diffts, dmtp
        sd = fp32_sd(g, seed=11)
        ref = u2Tokenizer(embed_size=g.hidden_size, num_heads=g.u2t_num_heads, num_layers=g.u2t_num_layers,
                          top_k=g.u2t_top_k, use_multi_scale=True, num_3d_query_token=g.num_3d_query_token,
                          hidden_size=g.hidden_size, attn_type=attn_type, enable_diffts=diffts, enable_dmtp=dmtp)
        ref.load_state_dict(_sub(sd, "model.u2tokenizer."), strict=True)
        gen = torch.Generator().manual_seed(21)
        v = torch.randn(2, 3, g.tokens_per_frame, g.hidden_size, generator=gen).bfloat16().float()
        t = torch.randn(2, 24, g.hidden_size, generator=gen).bfloat16().float()
        with torch.no_grad():
            gold[f"u2tok_{attn_type}_{int(diffts)}{int(dmtp)}"] = ref(v_token=v, t_token=t).clone()
    g = golden_geometry()
    sd = fp32_sd(g, seed=11)
    spp = SpatialPoolingProjector(image_size=g.image_size, patch_size=g.patch_size, in_dim=g.vit_hidden,
                                  out_dim=g.hidden_size, layer_type="mlp", layer_num=2, pooling_type="spatial",
                                  pooling_size=2)
    spp.load_state_dict(_sub(sd, "model.mm_projector."), strict=True)
    gen = torch.Generator().manual_seed(22)
    x = torch.randn(3, g.n_patches, g.vit_hidden, generator=gen).bfloat16().float()
    with torch.no_grad():
        gold["spp"] = spp(x).clone()
    torch.save(gold, os.path.join(OUT, "u2_reference_outputs.pt"))
    print({k: tuple(v.shape) for k, v in gold.items()})


if __name__ == "__main__":
    main()
