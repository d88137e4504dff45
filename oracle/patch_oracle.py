"""CPU restatement of the reference's PatchPreprocessor (TEST INFRASTRUCTURE ONLY - never imported by product code).

Follows data/preprocessing/patches.py: `initialise` :29-46 (mean / std over non-zero points), `interpolate` :48-52
(scipy interp1d from the 400..3980|3998 cm^-1 grid, step 2, onto 650..3898, step 2), `__call__` :54-107 (None ->
zeros, interpolate, standardise, trim to whole patches, view / unfold, attention mask).  numpy only; parity pinned by
tests/golden/patches.pt, which tests/golden/make_patches_golden.py produced by running the reference class itself.
"""
import numpy as np


def statistics(spectra):
    arr = np.array(spectra)
    nz = arr[arr != 0]
    return nz.mean(), nz.std()


def interpolate(spectrum):
    """Linear interpolation old grid -> new grid, written out (no scipy): every new abscissa lies on an old knot."""
    spectrum = np.asarray(spectrum, dtype=np.float64)
    old_x = np.arange(400, 4000 if len(spectrum) == 1800 else 3982, 2)
    new_x = np.arange(650, 3900, 2)
    lo = np.clip(np.searchsorted(old_x, new_x, side="left"), 1, len(old_x) - 1)
    x0, x1 = old_x[lo - 1], old_x[lo]
    y0, y1 = spectrum[lo - 1], spectrum[lo]
    return y0 + (y1 - y0) / (x1 - x0) * (new_x - x0)


def gradient(x):
    """torch.gradient(x, dim=-1)[0] for unit spacing: central differences, one-sided at both ends (patches.py:92)."""
    g = np.empty_like(x)
    g[:, 1:-1] = (x[:, 2:] - x[:, :-2]) / np.float32(2)
    g[:, 0] = x[:, 1] - x[:, 0]
    g[:, -1] = x[:, -1] - x[:, -2]
    return g


def patch_preprocess(spectra, mean, std, patch_size, masking=False, interpolation=False, overlap=1, derivative=False):
    """-> (patches float32 [B, P, patch_size], attention_mask bool [B, P])."""
    sizes = [len(s) if s is not None else -1 for s in spectra]
    n = max(sizes) if max(sizes) != -1 else 500
    rows = [s if s is not None else [0] * n for s in spectra]
    if interpolation:
        rows = [interpolate(r) for r in rows]
    x = np.asarray(rows, dtype=np.float32)                      # torch.Tensor(spectra): float32
    raw = x
    x = (x - np.float32(mean)) / np.float32(std)
    n_patches = x.shape[1] // patch_size
    x = x[:, : n_patches * patch_size]
    if overlap == 1:
        p = x.reshape(-1, n_patches, patch_size)
    else:
        hop = patch_size // overlap
        count = (x.shape[1] - patch_size) // hop + 1
        p = np.stack([x[:, i * hop: i * hop + patch_size] for i in range(count)], axis=1)
    if derivative:  # patches.py:91-95: gradient of the raw (not standardised) tensor, never overlapping
        g = gradient(raw)[:, : n_patches * patch_size].reshape(-1, n_patches, patch_size)
        p = np.concatenate([p, g], axis=1)
    if masking:
        mask = p.sum(-1) == 0
    else:
        mask = np.repeat(np.asarray([s == -1 for s in sizes])[:, None], p.shape[1], axis=1)
    return p, mask
