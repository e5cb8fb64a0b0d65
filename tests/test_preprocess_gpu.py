"""GPU volume preprocessing (u2_preprocess_volume_f32) against the CPU restatement of u2Transform.adaptive_resize
(oracle/u2_preprocess_oracle.py). Tolerance: the data-dependent integers (foreground box, resized extents) exact, the two
percentiles to 1e-12 relative (float64 on both sides), voxels to 2e-5 absolute on the [0, 1] scale (fp32 separable
filtering + trilinear weights in a different summation order)."""
import pytest
import torch

from oracle import u2_preprocess_oracle as P

pytestmark = pytest.mark.gpu


def body_volume(shape, box, seed, integer=False, air=-1000.0):
    g = torch.Generator().manual_seed(seed)
    vol = torch.full(shape, air)
    (d0, d1), (h0, h1), (w0, w1) = box
    blob = torch.rand(d1 - d0, h1 - h0, w1 - w0, generator=g) * 1400 - 200
    if integer:
        blob = blob.round()          # CT numbers: heavy ties, exactly representable
    vol[d0:d1, h0:h1, w0:w1] = blob
    return vol


CASES = {
    # name: (volume, target, pad_depth)
    "upsample_in_plane": lambda: (torch.rand(40, 96, 80, generator=torch.Generator().manual_seed(1)) * 100 + 1, 256, 64),
    "downsample_in_plane": lambda: (body_volume((48, 300, 280), ((4, 44), (10, 290), (20, 270)), 2), 128, 64),
    "depth_above_padding": lambda: (torch.rand(100, 64, 72, generator=torch.Generator().manual_seed(3)) + 0.25, 128, 64),
    "ct_like_integer_ties": lambda: (body_volume((37, 130, 150), ((5, 30), (17, 120), (9, 141)), 4, integer=True), 64, 32),
    "odd_sizes_no_background": lambda: (torch.randn(33, 61, 47, generator=torch.Generator().manual_seed(5)) * 30, 96, 64),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_preprocess_matches_oracle(name):
    from u2tokenizer_b200 import ops
    vol, target, pad = CASES[name]()
    ref, rinfo = P.adaptive_resize(vol, target=target, padding_size=pad)
    out, info_dev = ops.preprocess_volume(vol.cuda(), target=target, pad_depth=pad)
    info = ops.preprocess_info(info_dev)
    assert info["status"] == 0, info
    assert abs(info["a_min"] - rinfo["a_min"]) <= 1e-12 * max(1.0, abs(rinfo["a_min"]))
    assert abs(info["a_max"] - rinfo["a_max"]) <= 1e-12 * max(1.0, abs(rinfo["a_max"]))
    assert info["lo"] == rinfo["lo"] and info["hi"] == rinfo["hi"]
    oh, ow, od = rinfo["scaling"]
    assert info["out"] == [od, oh, ow]
    assert tuple(out.shape) == tuple(ref.shape) == (pad // 32, 32, target, target)
    err = (out.cpu() - ref).abs().max().item()
    assert err < 2e-5, (name, err, info)
    o3 = out.view(pad, target, target)
    assert float(o3[od:].abs().sum()) == 0 and float(o3[:, oh:].abs().sum()) == 0 and float(o3[:, :, ow:].abs().sum()) == 0


def test_preprocess_flags_degenerate_inputs_and_checks_arguments():
    from u2tokenizer_b200 import ops
    flat = torch.full((8, 16, 16), 3.0, device="cuda")       # a_min == a_max, nothing above it: the reference crashes
    out, info_dev = ops.preprocess_volume(flat, target=32, pad_depth=32)
    info = ops.preprocess_info(info_dev)
    assert info["status"] == 1 and float(out.abs().sum()) == 0
    with pytest.raises(TypeError):
        ops.preprocess_volume(flat.double())
    with pytest.raises(ValueError):
        ops.preprocess_volume(flat, pad_depth=40)
    with pytest.raises(RuntimeError):
        ops.preprocess_volume(flat.cpu())


def test_preprocessed_study_feeds_the_model():
    """preprocess -> [C, 32, H, W] chunks -> the `images` argument of the model surface (cfg-1-sized geometry)."""
    from common import tiny_geometry
    from test_engine_gpu import build
    from u2tokenizer_b200 import ops
    g = tiny_geometry(image_size=[32, 64, 64])
    eng, sd = build(g, 21)
    vol = body_volume((50, 90, 70), ((3, 47), (8, 80), (5, 66)), 7)
    chunks, info_dev = ops.preprocess_volume(vol.cuda(), target=64, pad_depth=64)
    assert tuple(chunks.shape) == (2, 32, 64, 64)
    ref, _ = P.adaptive_resize(vol, target=64, padding_size=64)
    feats = eng.encode_images(chunks.view(2, 1, 32, 64, 64))
    from oracle import u2_oracle as O
    with torch.no_grad():
        rfeats = O.encode_images(sd, ref.view(2, 1, 32, 64, 64), g)
    from common import cosine, rel_err
    assert rel_err(feats.float().cpu(), rfeats) < 3e-2 and cosine(feats.float().cpu(), rfeats) > 0.999
