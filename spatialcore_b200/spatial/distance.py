"""Drop-in ``calculate_domain_distances`` / ``get_distance_matrix`` on B200
[R src/spatialcore/spatial/distance.py:46-500].

Same arguments, outputs (``adata.obs`` distance / nearest-domain columns, ``adata.uns
['domain_distances']``), control flow and error messages as the reference; the compiled routines it
calls are replaced: ``cKDTree(target).query(source, k=1)`` by ``sc_cross_nn`` (grid-hashed exact 1-NN)
and ``scipy.spatial.distance.cdist(a, b).min() / .mean()`` by ``sc_pairwise_reduce`` (FP64 brute force
on the device).  The per-domain bookkeeping stays pandas on the host, like the reference.
"""

from __future__ import annotations

from typing import List, Optional

import numpy as np
import pandas as pd
import torch

from spatialcore_b200 import engine
from spatialcore_b200.core.logging import get_logger
from spatialcore_b200.core.metadata import update_metadata

logger = get_logger(__name__)


def _dev(coords: np.ndarray, device) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(coords[:, :2], dtype=np.float64)).to(device)


def calculate_domain_distances(
    adata,
    source_domain_column: str,
    target_domain_column: str,
    source_domain_subset: Optional[List[str]] = None,
    target_domain_subset: Optional[List[str]] = None,
    distance_metric: str = "minimum",
    output_mode: str = "both",
    output_distance_column: str = "distance_to_target",
    output_nearest_column: str = "nearest_target_domain",
    copy: bool = False,
    *,
    device="cuda",
):
    if "spatial" not in adata.obsm:
        raise ValueError(f"adata.obsm['spatial'] not found. Available keys: {list(adata.obsm.keys())}")
    if source_domain_column not in adata.obs.columns:
        raise ValueError(
            f"Source column '{source_domain_column}' not found in adata.obs. Available columns: {list(adata.obs.columns)}"
        )
    if target_domain_column not in adata.obs.columns:
        raise ValueError(
            f"Target column '{target_domain_column}' not found in adata.obs. Available columns: {list(adata.obs.columns)}"
        )
    if distance_metric not in ["minimum", "centroid", "mean"]:
        raise ValueError(f"Invalid distance_metric: '{distance_metric}'. Must be 'minimum', 'centroid', or 'mean'.")
    if output_mode not in ["cell", "matrix", "both"]:
        raise ValueError(f"Invalid output_mode: '{output_mode}'. Must be 'cell', 'matrix', or 'both'.")
    adata = adata.copy() if copy else adata
    logger.info(
        f"Calculating domain distances: {source_domain_column} → {target_domain_column} "
        f"(metric={distance_metric}, mode={output_mode})"
    )
    source_domains = adata.obs[source_domain_column].dropna().unique().tolist()
    target_domains = adata.obs[target_domain_column].dropna().unique().tolist()
    if source_domain_subset:
        source_domains = [d for d in source_domains if d in source_domain_subset]
    if target_domain_subset:
        target_domains = [d for d in target_domains if d in target_domain_subset]
    if not source_domains:
        raise ValueError(f"No valid source domains found in '{source_domain_column}'")
    if not target_domains:
        raise ValueError(f"No valid target domains found in '{target_domain_column}'")

    distance_matrix = pd.DataFrame(index=source_domains, columns=target_domains, dtype=float)
    spatial = np.asarray(adata.obsm["spatial"])
    same_column = source_domain_column == target_domain_column
    src_vals = adata.obs[source_domain_column].values
    tgt_vals = adata.obs[target_domain_column].values

    # device copies of each domain's coordinates, made on first use
    cache = {}

    def dom_coords(column_vals, name, tag):
        key = (tag, name)
        if key not in cache:
            mask = np.asarray(column_vals == name)
            cache[key] = (_dev(spatial[mask], device) if mask.any() else None)
        return cache[key]

    def per_cell_nearest():
        """Nearest target cell of every source cell (the cKDTree branch of the reference)."""
        target_mask = adata.obs[target_domain_column].isin(target_domains)
        target_indices = np.where(target_mask.values)[0]
        target_coords = spatial[target_indices]
        target_domains_arr = adata.obs[target_domain_column].iloc[target_indices].values
        source_mask = adata.obs[source_domain_column].isin(source_domains)
        source_indices = np.where(source_mask.values)[0]
        source_coords = spatial[source_indices]
        if len(source_coords) == 0 or len(target_coords) == 0:
            return None
        distances, nearest_idx = engine.cross_nn(target_coords, source_coords, device=device)
        nearest_domains = target_domains_arr[nearest_idx]
        dist_col_idx = adata.obs.columns.get_loc(output_distance_column)
        nearest_col_idx = adata.obs.columns.get_loc(output_nearest_column)
        adata.obs.iloc[source_indices, dist_col_idx] = distances
        adata.obs.iloc[source_indices, nearest_col_idx] = nearest_domains
        return source_indices, source_coords, target_coords, target_domains_arr, distances, nearest_domains

    if output_mode in ["cell", "both"]:
        adata.obs[output_distance_column] = np.nan
        adata.obs[output_nearest_column] = None

    if distance_metric == "minimum" and output_mode in ["cell", "both"]:
        res = per_cell_nearest()
        if res is not None:
            source_indices, source_coords, target_coords, target_domains_arr, distances, nearest_domains = res
            source_domains_arr = adata.obs[source_domain_column].iloc[source_indices].values
            for src in source_domains:
                src_mask = np.asarray(source_domains_arr == src)
                if not src_mask.any():
                    continue
                src_distances = distances[src_mask]
                src_nearest = nearest_domains[src_mask]
                for tgt in target_domains:
                    if src == tgt and same_column:
                        distance_matrix.loc[src, tgt] = 0.0
                        continue
                    tgt_mask_local = np.asarray(src_nearest == tgt)
                    if tgt_mask_local.any():
                        distance_matrix.loc[src, tgt] = src_distances[tgt_mask_local].min()
                    else:
                        tgt_cell_mask = np.asarray(target_domains_arr == tgt)
                        if tgt_cell_mask.any():
                            dmin, _ = engine.pairwise_reduce(_dev(source_coords[src_mask], device),
                                                             _dev(target_coords[tgt_cell_mask], device))
                            distance_matrix.loc[src, tgt] = dmin

    elif distance_metric == "centroid":
        source_centroids, target_centroids = {}, {}
        for src in source_domains:
            coords = spatial[np.asarray(src_vals == src)]
            if len(coords) > 0:
                source_centroids[src] = coords.mean(axis=0)
        for tgt in target_domains:
            coords = spatial[np.asarray(tgt_vals == tgt)]
            if len(coords) > 0:
                target_centroids[tgt] = coords.mean(axis=0)
        for src in source_domains:
            if src not in source_centroids:
                continue
            for tgt in target_domains:
                if src == tgt and same_column:
                    distance_matrix.loc[src, tgt] = 0.0
                    continue
                if tgt not in target_centroids:
                    continue
                distance_matrix.loc[src, tgt] = np.linalg.norm(source_centroids[src] - target_centroids[tgt])
        if output_mode in ["cell", "both"]:
            # nearest target CENTROID of every source cell (the reference loops over rows in Python)
            source_mask = adata.obs[source_domain_column].isin(source_domains).values
            src_idx = np.where(source_mask)[0]
            names = list(target_centroids.keys())
            if len(src_idx) and names:
                cent = np.stack([target_centroids[t] for t in names])
                d = np.linalg.norm(spatial[src_idx][:, None, :] - cent[None, :, :], axis=2)
                if same_column:  # a cell never measures to its own domain's centroid
                    d[np.asarray(src_vals[src_idx])[:, None] == np.asarray(names, dtype=object)[None, :]] = np.inf
                best = d.argmin(1)  # first minimum, like the reference's strict `<` scan in dict order
                bestd = d[np.arange(len(src_idx)), best]
                found = np.isfinite(bestd)
                dist_col_idx = adata.obs.columns.get_loc(output_distance_column)
                nearest_col_idx = adata.obs.columns.get_loc(output_nearest_column)
                adata.obs.iloc[src_idx, dist_col_idx] = bestd
                adata.obs.iloc[src_idx[found], nearest_col_idx] = np.asarray(names, dtype=object)[best[found]]

    elif distance_metric == "mean":
        for src in source_domains:
            a = dom_coords(src_vals, src, "s")
            if a is None:
                continue
            for tgt in target_domains:
                if src == tgt and same_column:
                    distance_matrix.loc[src, tgt] = 0.0
                    continue
                b = dom_coords(tgt_vals, tgt, "t")
                if b is None:
                    continue
                _, dsum = engine.pairwise_reduce(a, b)
                distance_matrix.loc[src, tgt] = dsum / (a.shape[0] * b.shape[0])
        if output_mode in ["cell", "both"]:
            per_cell_nearest()  # the reference falls back to the minimum distance per cell

    else:  # minimum, matrix only
        for src in source_domains:
            a = dom_coords(src_vals, src, "s")
            if a is None:
                continue
            for tgt in target_domains:
                if src == tgt and same_column:
                    distance_matrix.loc[src, tgt] = 0.0
                    continue
                b = dom_coords(tgt_vals, tgt, "t")
                if b is None:
                    continue
                dmin, _ = engine.pairwise_reduce(a, b)
                distance_matrix.loc[src, tgt] = dmin

    valid = distance_matrix.values[~np.isnan(distance_matrix.values.astype(float))].astype(float)
    summary = {
        "min_distance": float(valid.min()) if len(valid) > 0 else None,
        "max_distance": float(valid.max()) if len(valid) > 0 else None,
        "mean_distance": float(valid.mean()) if len(valid) > 0 else None,
        "median_distance": float(np.median(valid)) if len(valid) > 0 else None,
    }
    if len(valid) > 0:
        logger.info(
            f"Distance statistics: min={summary['min_distance']:.1f}, "
            f"max={summary['max_distance']:.1f}, mean={summary['mean_distance']:.1f}"
        )
    if output_mode in ["matrix", "both"]:
        adata.uns["domain_distances"] = {
            "source_domain_column": source_domain_column,
            "target_domain_column": target_domain_column,
            "distance_metric": distance_metric,
            "source_domains": source_domains,
            "target_domains": target_domains,
            "summary_statistics": summary,
            "distance_matrix": distance_matrix.to_dict(orient="index"),
        }
    outputs = {"summary_statistics": summary}
    if output_mode in ["cell", "both"]:
        outputs["obs_distance"] = output_distance_column
        outputs["obs_nearest"] = output_nearest_column
    if output_mode in ["matrix", "both"]:
        outputs["uns"] = "domain_distances"
    update_metadata(
        adata,
        function_name="calculate_domain_distances",
        parameters={
            "source_domain_column": source_domain_column,
            "target_domain_column": target_domain_column,
            "source_domain_subset": source_domain_subset,
            "target_domain_subset": target_domain_subset,
            "distance_metric": distance_metric,
            "output_mode": output_mode,
        },
        outputs=outputs,
    )
    return adata


def get_distance_matrix(adata, key: str = "domain_distances") -> pd.DataFrame:
    """Distance matrix (source domains x target domains) as a DataFrame [R distance.py:452-500]."""
    if key not in adata.uns:
        raise KeyError(
            f"'{key}' not found in adata.uns. "
            "Run calculate_domain_distances() with output_mode='matrix' or 'both' first."
        )
    data = adata.uns[key]
    if "distance_matrix" not in data:
        raise KeyError(f"'distance_matrix' not found in adata.uns['{key}']")
    return pd.DataFrame(data["distance_matrix"]).T
