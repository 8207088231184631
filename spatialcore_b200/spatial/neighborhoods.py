"""Drop-in ``compute_neighborhood_profile`` on B200 [R src/spatialcore/spatial/neighborhoods.py:48-296].

The kNN / radius query and the per-cell cell-type histogram run fused in one kernel
(``sc_grid_knn`` / ``sc_grid_radius_count`` with the composition epilogue): neighbour indices are
never materialised.
"""

from __future__ import annotations

from typing import Literal, Optional

import numpy as np
import torch

from spatialcore_b200 import engine
from spatialcore_b200.core.logging import get_logger
from spatialcore_b200.core.metadata import update_metadata

logger = get_logger(__name__)


def compute_neighborhood_profile(
    adata,
    celltype_column: str,
    method: Literal["knn", "radius"] = "knn",
    k: int = 15,
    radius: Optional[float] = None,
    normalize: bool = True,
    spatial_key: str = "spatial",
    key_added: str = "neighborhood_profile",
    copy: bool = False,
    *,
    device="cuda",
):
    if spatial_key not in adata.obsm:
        raise ValueError(
            f"adata.obsm['{spatial_key}'] not found. "
            "Spatial coordinates are required for neighborhood computation."
        )
    if celltype_column not in adata.obs.columns:
        raise ValueError(
            f"Column '{celltype_column}' not found in adata.obs. "
            f"Available columns: {list(adata.obs.columns)[:10]}..."
        )
    if method not in ["knn", "radius"]:
        raise ValueError(f"Invalid method: '{method}'. Must be 'knn' or 'radius'.")
    n = adata.n_obs
    if method == "knn" and k < 1:
        raise ValueError(f"k must be >= 1, got {k}")
    if method == "knn" and k >= n:
        raise ValueError(f"k must be < number of cells ({n}), got {k}")
    if method == "radius":
        if radius is None:
            raise ValueError("'radius' must be provided when method='radius'.")
        if radius <= 0:
            raise ValueError(f"radius must be > 0, got {radius}")
    adata = adata.copy() if copy else adata

    series = adata.obs[celltype_column]
    if series.isna().any():
        raise ValueError(
            f"{int(series.isna().sum())} cells have missing labels in '{celltype_column}'. "
            "Fill or remove missing labels before computing neighborhoods."
        )
    celltypes = sorted(series.unique())
    n_types = len(celltypes)
    if n_types < 2:
        raise ValueError(f"At least 2 unique cell types required, found {n_types}. Check column '{celltype_column}'.")
    logger.info(f"Computing neighborhood profiles: {n:,} cells, {n_types} cell types, method={method}")

    lookup = {ct: i for i, ct in enumerate(celltypes)}
    codes = np.fromiter((lookup[v] for v in series.values), dtype=np.int32, count=n)
    labels = torch.from_numpy(codes).to(device)
    if method == "knn":
        _, _, profile = engine.knn_graph(adata.obsm[spatial_key], k, labels=labels, n_types=n_types, want_idx=False, device=device)
    else:
        _, profile = engine.radius_graph(adata.obsm[spatial_key], radius, labels=labels, n_types=n_types, want_graph=False, device=device)
    n_empty = engine.profile_normalize(profile, normalize)
    if n_empty > 0:
        raise ValueError(
            f"{n_empty} cells have empty neighborhood profiles. "
            "Increase radius, switch to knn, or pre-filter isolated cells before profiling."
        )
    adata.obsm[key_added] = profile.cpu().numpy()
    adata.uns[f"{key_added}_celltypes"] = list(celltypes)
    logger.info(f"Stored neighborhood profiles in adata.obsm['{key_added}'] (shape: {tuple(profile.shape)})")
    update_metadata(
        adata,
        function_name="compute_neighborhood_profile",
        parameters={
            "celltype_column": celltype_column,
            "method": method,
            "k": k if method == "knn" else None,
            "radius": radius if method == "radius" else None,
            "normalize": normalize,
            "spatial_key": spatial_key,
        },
        outputs={"obsm": key_added, "uns": f"{key_added}_celltypes", "n_celltypes": n_types, "n_cells": n},
    )
    return adata
