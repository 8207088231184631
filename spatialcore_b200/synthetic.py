"""Seeded synthetic Xenium / CosMx-shaped inputs.  Coordinates are generated on the
host with numpy; expression is generated on the device in gene tiles (20 GB at 5M x 1k would take
minutes with numpy) and is deterministic for a given seed and GPU model."""

from __future__ import annotations

import math


import numpy as np
import torch


def coords_uniform(n: int, extent: float, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return rng.uniform(0.0, extent, (n, 2))


def coords_mixture(n: int, extent: float, seed: int, blob_frac: float = 0.3, n_blobs: int = 40) -> np.ndarray:
    """70 % uniform + 30 % Gaussian blobs (tissue-like density contrast), shuffled."""
    rng = np.random.default_rng(seed)
    nb = int(blob_frac * n)
    uni = rng.uniform(0.0, extent, (n - nb, 2))
    centers = rng.uniform(0.1 * extent, 0.9 * extent, (n_blobs, 2))
    which = rng.integers(0, n_blobs, nb)
    blob = centers[which] + rng.normal(0.0, 0.02 * extent, (nb, 2))
    c = np.concatenate([uni, np.clip(blob, 0.0, extent)])
    return c[rng.permutation(n)]


def radius_for_mean_degree(n: int, extent: float, degree: float) -> float:
    return math.sqrt(degree * extent * extent / (math.pi * n))


def expression_device(coords: np.ndarray, g: int, seed: int, device="cuda", smooth_frac: float = 0.25,
                      tile: int = 64) -> torch.Tensor:
    """log1p of Poisson counts (mean ~0.4, ~70-80 % zeros); a quarter of the genes carry a smooth
    spatial pattern (random low-frequency cosines).  float32 [n, g] on the device."""
    n = coords.shape[0]
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    c = torch.from_numpy(np.ascontiguousarray(coords, dtype=np.float32)).to(device)
    ext = float(coords.max() - coords.min()) or 1.0
    X = torch.empty((n, g), dtype=torch.float32, device=device)
    n_smooth = int(smooth_frac * g)
    for g0 in range(0, g, tile):
        g1 = min(g, g0 + tile)
        w = g1 - g0
        rate = torch.full((n, w), 0.4, dtype=torch.float32, device=device)
        ns = max(0, min(g1, n_smooth) - g0)
        if ns > 0:
            freq = (torch.rand((2, ns), generator=gen, device=device) * 6.0 + 1.0) * (2.0 * math.pi / ext)
            phase = torch.rand((1, ns), generator=gen, device=device) * (2.0 * math.pi)
            field = torch.cos(c[:, 0:1] * freq[0:1] + c[:, 1:2] * freq[1:2] + phase)
            rate[:, :ns] = 0.4 * (1.0 + 0.8 * field)
        X[:, g0:g1] = torch.log1p(torch.poisson(rate, generator=gen))
    return X


def patchy_labels(coords: np.ndarray, n_types: int, seed: int, n_seeds: int = 300, noise: float = 0.2) -> np.ndarray:
    """Cell types in spatial patches: nearest of ``n_seeds`` random seeds picks the type, plus noise."""
    from scipy.spatial import cKDTree

    rng = np.random.default_rng(seed)
    lo, hi = coords.min(0), coords.max(0)
    seeds = rng.uniform(lo, hi, (n_seeds, 2))
    seed_type = rng.integers(0, n_types, n_seeds)
    _, nearest = cKDTree(seeds).query(coords, k=1, workers=-1)
    lab = seed_type[nearest]
    flip = rng.random(coords.shape[0]) < noise
    return np.where(flip, rng.integers(0, n_types, coords.shape[0]), lab).astype(np.int32)
