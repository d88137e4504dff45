"""Device-side `PatchPreprocessor` (reference: data/preprocessing/patches.py:14-107).

Same constructor fields and `initialise` statistics as the reference class; `__call__` takes the raw spectra of a
batch and returns `(patches [B, P, patch_size] fp32, attention_mask [B, P] bool)` - computed by one CUDA kernel on
the device (standardise + the interpolation, which is a slice because both wavenumber grids share their knots +
trim + patch), batch-first, so the collator no longer makes a host pass over B x 1791 floats nor a transpose.
`derivative=True` appends the patches of torch.gradient(raw spectrum) (patches.py:91-95) from the same launch pair.
With `masking=True` a patch is flagged when its fp32 sum is exactly 0 (patches.py:98-100); a gradient patch telescopes to
about 0, so for those patches the flag is a function of the summation order and can differ from a torch-CPU run on near-ties
(as it does between torch's own CPU and CUDA sums).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import ops

# interpolate(): old grid 400..3980 (1791 points) or 400..3998 (1800 points), step 2 -> new grid 650..3898, step 2
INTERP_OFFSET = (650 - 400) // 2
INTERP_POINTS = (3900 - 650) // 2


@dataclass
class DevicePatchPreprocessor:
    patch_size: int
    masking: bool = False
    interpolation: bool = False
    overlap: int = 1
    derivative: bool = False
    encoding_type: str = ""
    mean: float = field(init=False, default=0.0)
    std: float = field(init=False, default=1.0)
    mean_deriv: Optional[float] = field(init=False, default=None)  # computed like the reference's, and like there unused
    std_deriv: Optional[float] = field(init=False, default=None)

    def initialise(self, spectra: Union[np.ndarray, Sequence[Sequence[float]]], modality: Optional[str] = None) -> None:
        """Mean / std over the non-zero points of the sampled spectra (patches.py:37-39).  Accepts the array itself or
        a `datasets.Dataset`-like mapping with the modality column."""
        if modality is not None and not isinstance(spectra, np.ndarray):
            spectra = spectra[modality]
        arr = np.array(spectra)
        self.mean = float(arr[arr != 0].mean())
        self.std = float(arr[arr != 0].std())
        if self.derivative:  # patches.py:41-46 (torch.gradient, unbiased std)
            g = np.gradient(arr.astype(np.float32), axis=-1)
            self.mean_deriv, self.std_deriv = float(g.mean()), float(g.std(ddof=1))

    def __call__(self, spectra: Union[torch.Tensor, List[Optional[List[float]]]], device="cuda") -> Tuple[torch.Tensor, torch.Tensor]:
        missing = None
        if not isinstance(spectra, torch.Tensor):
            sizes = [len(s) if s is not None else -1 for s in spectra]
            n = max(sizes) if max(sizes) != -1 else 500
            missing = torch.tensor([s == -1 for s in sizes], dtype=torch.uint8)
            spectra = torch.tensor([s if s is not None else [0.0] * n for s in spectra], dtype=torch.float32)
        raw = spectra.to(device=device, dtype=torch.float32, non_blocking=True)
        if raw.stride(1) != 1:
            raw = raw.contiguous()
        B, n_pts = raw.shape
        offset, n_use = (INTERP_OFFSET, INTERP_POINTS) if self.interpolation else (0, n_pts)
        if self.interpolation and n_pts not in (1791, 1800):
            raise ValueError(f"interpolation expects 1791 or 1800 points, got {n_pts}")
        n_patches = n_use // self.patch_size
        hop = self.patch_size // self.overlap
        P = n_patches if self.overlap == 1 else (n_patches * self.patch_size - self.patch_size) // hop + 1
        Pd = n_patches if self.derivative else 0
        out = torch.empty(B, P + Pd, self.patch_size, dtype=torch.float32, device=raw.device)
        pad = torch.empty(B, P + Pd, dtype=torch.uint8, device=raw.device)
        miss_dev = None if missing is None else missing.to(raw.device, non_blocking=True)
        if self.derivative:
            ops.patchify_deriv(raw, out, self.mean, self.std, Pd, offset=offset, n_use=n_use, hop=hop, pad=pad,
                               missing=miss_dev, masking=self.masking)
        else:
            ops.patchify(raw, out, self.mean, self.std, offset=offset, hop=hop, pad=pad, missing=miss_dev,
                         masking=self.masking)
        return out, pad.bool()
