"""CPU oracle for the volume preprocessing that feeds the hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Restates `u2Transform.adaptive_resize` (reference src/utils/u2Transform.py:62-122, validation pipeline :47-56):

    ScaleIntensityRangePercentiles(0.5, 99.5, b_min 0, b_max 1, clip)  ->  CropForeground  ->  permute to (H, W, D)
    -> in-plane scale so that the larger in-plane side becomes 256 (anti-aliased trilinear, align_corners=True; the depth
       keeps its size when it is <= 256 and is resized to 256 otherwise)  ->  zero-pad to [256, 256, 256]
    -> permute to (D, H, W)  ->  view(-1, 32, 256, 256)

The arithmetic of the first two steps and of the resize lives in MONAI (third-party, `monai==1.3.0`,
requirements.txt:52; not vendored, not installed in this image), so this file restates MONAI 1.3.0's published
definitions and says which ones:
  * `monai.transforms.utils_pytorch_numpy_unification.percentile`: arrays of more than 1e6 elements (and all numpy
    inputs) go through `np.percentile` (linear interpolation, float64 here because `nib...get_fdata()` yields float64);
  * `ScaleIntensityRange.__call__`: (img - a_min) / (a_max - a_min) * (b_max - b_min) + b_min, clip, cast to float32;
  * `CropForeground` (array version; the reference's stray `source_key=` lands in the unused pad kwargs): bounding box of
    `img > 0` over the spatial axes, margin 0;
  * `monai.transforms.spatial.functional.resize` with `anti_aliasing=True`: per-axis factor in/out (float32), sigma =
    max((factor - 1) / 2, 0), `GaussianSmooth(sigma)` = separable zero-padded convolution with
    `gaussian_1d(sigma, truncated=4.0, approx="erf", normalize=False)`, only when some axis shrinks; then
    `torch.nn.functional.interpolate(mode="trilinear", align_corners=True)`.
**parity unpinned**: the reference holds no test or golden vector for this path and MONAI cannot be imported here; the
anchor is the call site above. Only tests/ may import this module.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def scale_intensity_range_percentiles(vol: torch.Tensor, lower: float = 0.5, upper: float = 99.5) -> tuple:
    """u2Transform.py:51 -> MONAI ScaleIntensityRangePercentiles(lower, upper, b_min=0, b_max=1, clip=True).
    vol: any float dtype; returns (float32 scaled volume, a_min, a_max)."""
    x = vol.double().numpy()
    a_min, a_max = (float(v) for v in np.percentile(x, [lower, upper]))
    if a_max - a_min == 0.0:
        out = x - a_min  # MONAI warns "Divide by zero (a_min == a_max)" and returns img - a_min + b_min
    else:
        out = np.clip((x - a_min) / (a_max - a_min), 0.0, 1.0)
    return torch.from_numpy(out.astype(np.float32)), a_min, a_max


def foreground_box(scaled: torch.Tensor):
    """u2Transform.py:52 -> MONAI generate_spatial_bounding_box(select_fn = x > 0, margin 0): per axis
    [first, last + 1) of the voxels that are positive. Raises when nothing is (the reference crashes further on)."""
    pos = scaled > 0
    if not bool(pos.any()):
        raise ValueError("no foreground voxel (all intensities at or below the lower percentile)")
    lo, hi = [], []
    for ax in range(3):
        other = tuple(a for a in range(3) if a != ax)
        idx = pos.any(dim=other).nonzero()
        lo.append(int(idx[0]))
        hi.append(int(idx[-1]) + 1)
    return lo, hi


def gaussian_1d(sigma: float, truncated: float = 4.0) -> torch.Tensor:
    """MONAI 1.3.0 `monai.networks.layers.convutils.gaussian_1d(approx="erf", normalize=False)`."""
    s = torch.as_tensor(sigma, dtype=torch.float32)
    tail = int(max(float(s) * truncated, 0.5) + 0.5)
    x = torch.arange(-tail, tail + 1, dtype=torch.float32)
    t = 0.70710678 / torch.abs(s)
    out = 0.5 * ((t * (x + 0.5)).erf() - (t * (x - 0.5)).erf())
    return out.clamp(min=0)


def monai_resize(img: torch.Tensor, out_size) -> torch.Tensor:
    """MONAI 1.3.0 `resize(img[C, *spatial], out_size, mode="bilinear" (-> trilinear for 3-D), align_corners=True,
    anti_aliasing=True, anti_aliasing_sigma=None)`, u2Transform.py:81-92."""
    in_size = list(img.shape[1:])
    x = img.float()
    if any(o < i for o, i in zip(out_size, in_size)):
        factors = torch.div(torch.Tensor(in_size), torch.Tensor(list(out_size)))
        sigma = torch.maximum(torch.zeros(factors.shape), (factors - 1) / 2).tolist()
        y = x.unsqueeze(0)  # [1, C, *spatial]
        for ax, s in enumerate(sigma):
            k = gaussian_1d(s)
            shape = [1, 1, 1, 1, 1]
            shape[2 + ax] = k.numel()
            pad = [0, 0, 0]
            pad[ax] = (k.numel() - 1) // 2
            y = F.conv3d(y, k.view(shape), padding=pad)  # zero padding (separable_filtering mode "zeros")
        x = y[0]
    return F.interpolate(x.unsqueeze(0), size=list(out_size), mode="trilinear", align_corners=True)[0]


def adaptive_resize(vol_dhw: torch.Tensor, target: int = 256, padding_size: int = 256, lower: float = 0.5,
                    upper: float = 99.5):
    """u2Transform.adaptive_resize on an in-memory volume. `vol_dhw` is the reference's `data[0]` right after
    `get_fdata().transpose(2, 0, 1)` (u2Transform.py:67): axis order (D, H, W). Returns (chunks
    [padding_size / 32, 32, target, target] float32, info dict)."""
    scaled, a_min, a_max = scale_intensity_range_percentiles(vol_dhw, lower, upper)
    lo, hi = foreground_box(scaled)
    crop = scaled[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]]
    data = crop.permute(1, 2, 0)                                  # :70  (H, W, D)
    shape = list(data.shape)
    ratio = min(target / shape[i] for i in range(2))             # :74
    scaling = [int(shape[i] * ratio) for i in range(2)]          # :75
    scaling.append(shape[2] if padding_size >= shape[2] else padding_size)   # :79-80 / :96-97
    res = monai_resize(data.unsqueeze(0), scaling)                # [1, h, w, d]
    res = F.pad(res, (0, padding_size - scaling[2], 0, target - scaling[1], 0, target - scaling[0]))
    out = res.permute(0, 3, 1, 2).reshape(-1, 32, target, target)  # :116-119
    return out, dict(a_min=a_min, a_max=a_max, lo=lo, hi=hi, scaling=scaling)
