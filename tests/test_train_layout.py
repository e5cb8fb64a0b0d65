"""CPU-side checks of the training layout (u2tokenizer_b200/train.py: Layout): fused groups are adjacent, every
parameter has one slot, and the ZeRO-1 ownership (bucket-interleaved: rank r owns the r-th 1/W slice of every bucket)
partitions the matrix region exactly for every world size - the host logic behind the N > 1 gradient exchange."""
import pytest

from common import tiny_geometry
from u2tokenizer_b200.train import Layout, is_vector_param


@pytest.mark.parametrize("world", [1, 2, 4, 8])
@pytest.mark.parametrize("bucket", [10_000, 50_000, 10**9])
def test_bucket_interleaved_ownership_partitions_the_flat_buffer(world, bucket):
    g = tiny_geometry()
    L = Layout(g, world_size=world, bucket_elems=bucket)
    assert L.mat_total == L.n_buckets * L.bucket and L.bucket == world * L.piece and L.piece % 8 == 0
    assert L.mat_total >= L.mat_used and L.mat_total - L.mat_used < L.n_buckets * 8 * world + 8
    owned = []
    for r in range(world):
        for i in range(L.n_buckets):
            lo = i * L.bucket + r * L.piece
            owned.append((lo, lo + L.piece))
    owned.sort()
    assert owned[0][0] == 0 and owned[-1][1] == L.mat_total
    assert all(a[1] == b[0] for a, b in zip(owned[:-1], owned[1:]))   # disjoint and complete


def test_every_parameter_has_one_slot_and_fused_groups_are_adjacent():
    for kw in (dict(), dict(qk_norm=False, tie_word_embeddings=True), dict(enable_diffts=False, enable_dmtp=False)):
        g = tiny_geometry(**kw)
        L = Layout(g)
        names = set(L.mat_names) | set(L.vec_names)
        assert names == set(L.shapes) and not (set(L.mat_names) & set(L.vec_names))
        for n in L.vec_names:
            assert is_vector_param(n, L.shapes[n])
        spans = sorted((L.mat_off[n], L.mat_off[n] + L._numel(n)) for n in L.mat_names)
        assert all(a[1] <= b[0] for a, b in zip(spans[:-1], spans[1:]))
        l0 = "model.layers.0."
        assert L.adjacent([l0 + "self_attn.q_proj.weight", l0 + "self_attn.k_proj.weight", l0 + "self_attn.v_proj.weight"])
        assert L.adjacent([l0 + "mlp.gate_proj.weight", l0 + "mlp.up_proj.weight"])
        a = "model.u2tokenizer.svt_module.attention_network.layers.0.spatial_attention."
        assert L.adjacent([a + "wq.weight", a + "wk.weight", a + "wv.weight"])
        assert L.adjacent([a + "wq.bias", a + "wk.bias", a + "wv.bias"])
        c = "model.u2tokenizer.tta_module.layers_vt.0.visual_cross_attention."
        assert L.adjacent([c + "wk.weight", c + "wv.weight"]) and L.adjacent([c + "wk.bias", c + "wv.bias"])
