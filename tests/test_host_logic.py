"""Host-side logic of the module surface that needs no GPU: argument checks that guard the fused CUDA path."""
import pytest
import torch


def _cls():
    from u2tokenizer_b200.modeling import U2LlamaForCausalLM
    return U2LlamaForCausalLM


def test_right_padded_masks_are_accepted():
    chk = _cls()._check_right_padded
    chk(None)
    chk(torch.ones(3, 7, dtype=torch.long))
    chk(torch.tensor([[1, 1, 1, 0, 0], [1, 1, 1, 1, 1], [1, 0, 0, 0, 0]]))
    chk(torch.tensor([[True, True, False]]))


@pytest.mark.parametrize("mask", [
    torch.tensor([[0, 0, 1, 1, 1], [1, 1, 1, 1, 1]]),   # left padding: the usual HF layout for batched generation
    torch.tensor([[1, 0, 1, 1, 0]]),                   # a hole
    torch.tensor([[0, 0, 0]]),                         # nothing valid
    torch.ones(2, 3, 4),                               # not [batch, tokens]
])
def test_other_masks_are_refused_loudly(mask):
    """The reference hands the mask to HF (u2llama.py:76-87); the fused path has none, so anything that would change the
    result must raise instead of being ignored."""
    with pytest.raises(NotImplementedError):
        _cls()._check_right_padded(mask)


def test_surface_matches_the_reference_call_forms():
    """forward / generate keep the reference's parameter names (u2llama.py:41-55, 90-96; train_stage1.py:244-250 passes
    images, input_ids, labels, attention_mask, question_ids by keyword)."""
    import inspect
    from u2tokenizer_b200.modeling import U2LlamaForCausalLM, U2Qwen3ForCausalLM
    want = ["images", "input_ids", "labels", "attention_mask", "question_ids", "position_ids", "past_key_values",
            "inputs_embeds", "use_cache", "output_attentions", "output_hidden_states", "return_dict"]
    for cls in (U2LlamaForCausalLM, U2Qwen3ForCausalLM):
        names = [n for n in inspect.signature(cls.forward).parameters if n != "self"]
        assert names[:len(want)] == want, names
        gen = [n for n in inspect.signature(cls.generate).parameters if n != "self"]
        assert gen[:2] == ["images", "inputs"] and "question_ids" in gen, gen
        for attr in ("get_model", "prepare_inputs_for_multimodal", "initialize_vision_tokenizer",
                     "prepare_inputs_for_generation", "per_token_logps"):
            assert hasattr(cls, attr), attr
