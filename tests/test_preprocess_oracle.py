"""The preprocessing oracle (oracle/u2_preprocess_oracle.py, a restatement of u2Transform.adaptive_resize + the MONAI 1.3.0
transforms it calls) checked against independent formulations on CPU. MONAI itself is not installed: "parity unpinned"."""
import numpy as np
import torch

from oracle import u2_preprocess_oracle as P


def test_percentile_scaling_matches_torch_quantile_and_clips():
    g = torch.Generator().manual_seed(0)
    v = torch.randn(7, 9, 11, generator=g) * 50 + 10
    scaled, a_min, a_max = P.scale_intensity_range_percentiles(v, 0.5, 99.5)
    q = torch.quantile(v.double().flatten(), torch.tensor([0.005, 0.995], dtype=torch.float64))
    assert abs(a_min - q[0].item()) < 1e-9 and abs(a_max - q[1].item()) < 1e-9
    assert scaled.dtype == torch.float32 and scaled.min() == 0 and scaled.max() == 1
    ref = ((v.double() - a_min) / (a_max - a_min)).clamp(0, 1).float()
    assert torch.equal(scaled, ref)


def test_gaussian_taps():
    assert torch.equal(P.gaussian_1d(0.0), torch.tensor([0.0, 1.0, 0.0]))      # sigma 0: identity, tail 1
    k = P.gaussian_1d(0.75)
    assert k.numel() == 2 * 3 + 1 and torch.allclose(k, k.flip(0)) and 0.999 < k.sum() <= 1.0
    assert P.gaussian_1d(2.0).numel() == 2 * 8 + 1


def test_resize_is_identity_at_equal_size_and_exact_on_linear_ramps():
    g = torch.Generator().manual_seed(1)
    x = torch.rand(1, 6, 7, 5, generator=g)
    assert torch.allclose(P.monai_resize(x, [6, 7, 5]), x, atol=1e-6)
    # upsampling (no anti-aliasing) of a linear ramp with align_corners=True stays a linear ramp
    ramp = torch.linspace(0, 1, 5).view(1, 5, 1, 1).expand(1, 5, 3, 2).contiguous()
    up = P.monai_resize(ramp, [9, 3, 2])
    assert torch.allclose(up[0, :, 0, 0], torch.linspace(0, 1, 9), atol=1e-6)
    # shrinking smooths first: a constant interior stays constant away from the zero-padded faces
    c = torch.ones(1, 40, 40, 4)
    dn = P.monai_resize(c, [10, 10, 4])
    assert torch.allclose(dn[0, 3:7, 3:7], torch.ones(4, 4, 4), atol=2e-3) and dn[0, 0, 0, 0] < 0.9


def test_adaptive_resize_layout_and_padding():
    g = torch.Generator().manual_seed(2)
    vol = torch.full((20, 50, 40), -1000.0)                      # air
    vol[3:17, 5:45, 8:36] = torch.rand(14, 40, 28, generator=g) * 400 + 1   # the body
    out, info = P.adaptive_resize(vol, target=64, padding_size=32)
    assert tuple(out.shape) == (1, 32, 64, 64) and out.dtype == torch.float32
    assert info["lo"] == [3, 5, 8] and info["hi"] == [17, 45, 36]
    assert info["scaling"] == [64, int(28 * (64 / 40)), 14]      # larger in-plane side -> 64, depth kept
    oh, ow, od = info["scaling"]
    vol3 = out.reshape(32, 64, 64)
    assert vol3[od:].abs().sum() == 0 and vol3[:, oh:].abs().sum() == 0 and vol3[:, :, ow:].abs().sum() == 0
    assert 0 <= vol3.min() and vol3.max() <= 1 and vol3[:od, :oh, :ow].mean() > 0.2
    # depth larger than the padding size is resized down to it (u2Transform.py:95-97)
    tall = torch.rand(70, 16, 16, generator=g) + 0.5
    out2, info2 = P.adaptive_resize(tall, target=32, padding_size=64)
    assert tuple(out2.shape) == (2, 32, 32, 32) and info2["scaling"][2] == 64
