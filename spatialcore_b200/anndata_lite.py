"""Minimal AnnData duck type.

The reference API takes ``anndata.AnnData`` [R src/spatialcore/spatial/autocorrelation.py:421-432].
``anndata`` is not installable in this image, so the drop-in functions accept any
object exposing the slots the hot path touches (``X, layers, obs, var_names, obsm,
obsp, uns, n_obs, copy()``); a real ``anndata.AnnData`` satisfies the same protocol.
``AnnDataLite`` is the stand-in used by the tests, the bench and the oracle shim.
"""

from __future__ import annotations

import copy as _copy
from typing import Any, Dict, Optional, Sequence

import numpy as np
import pandas as pd


class AnnDataLite:
    """Holds exactly the AnnData slots ``spatialcore.spatial`` reads and writes."""

    def __init__(
        self,
        X: Any,
        obs: Optional[pd.DataFrame] = None,
        var_names: Optional[Sequence[str]] = None,
        obsm: Optional[Dict[str, Any]] = None,
        obsp: Optional[Dict[str, Any]] = None,
        uns: Optional[Dict[str, Any]] = None,
        layers: Optional[Dict[str, Any]] = None,
        var: Optional[pd.DataFrame] = None,
    ) -> None:
        self.X = X
        n_obs, n_vars = X.shape
        self._obs = obs  # built on first access: a 5 M-row string index costs ~1 s and most calls never read it
        if var_names is None:
            var_names = [f"g{i}" for i in range(n_vars)]
        self.var_names = pd.Index(list(var_names))
        if len(self.var_names) != n_vars:
            raise ValueError(f"var_names has {len(self.var_names)} entries, X has {n_vars} columns")
        self.var = var if var is not None else pd.DataFrame(index=self.var_names)
        self.obsm = dict(obsm) if obsm else {}
        self.obsp = dict(obsp) if obsp else {}
        self.uns = dict(uns) if uns else {}
        self.layers = dict(layers) if layers else {}

    @property
    def obs(self) -> pd.DataFrame:
        if self._obs is None:
            self._obs = pd.DataFrame(index=pd.RangeIndex(self.X.shape[0]).astype(str))
        return self._obs

    @obs.setter
    def obs(self, value: pd.DataFrame) -> None:
        self._obs = value

    @property
    def n_obs(self) -> int:
        return self.X.shape[0]

    @property
    def n_vars(self) -> int:
        return self.X.shape[1]

    @property
    def shape(self):
        return self.X.shape

    @property
    def obs_names(self) -> pd.Index:
        return self.obs.index

    def copy(self) -> "AnnDataLite":
        return AnnDataLite(
            X=self.X.copy(),
            obs=self.obs.copy(),
            var_names=list(self.var_names),
            obsm={k: _copy.copy(v) for k, v in self.obsm.items()},
            obsp={k: v.copy() for k, v in self.obsp.items()},
            uns=_copy.deepcopy(self.uns),
            layers={k: v.copy() for k, v in self.layers.items()},
            var=self.var.copy(),
        )

    def __getitem__(self, key) -> "AnnDataLite":
        """Supports the one indexing form the reference uses: ``adata[:, gene_names]``
        [R autocorrelation.py:573]."""
        if not (isinstance(key, tuple) and len(key) == 2):
            raise IndexError("AnnDataLite supports adata[:, genes] only")
        rows, cols = key
        if not (isinstance(rows, slice) and rows == slice(None)):
            raise IndexError("AnnDataLite supports adata[:, genes] only")
        if isinstance(cols, str):
            cols = [cols]
        col_idx = np.asarray([self.var_names.get_loc(c) for c in cols], dtype=np.int64)
        return AnnDataLite(
            X=self.X[:, col_idx],
            obs=self.obs,
            var_names=[self.var_names[i] for i in col_idx],
            obsm=self.obsm,
            obsp=self.obsp,
            uns=self.uns,
            layers={k: v[:, col_idx] for k, v in self.layers.items()},
            var=self.var.iloc[col_idx],
        )

    def __repr__(self) -> str:
        return f"AnnDataLite object with n_obs x n_vars = {self.n_obs} x {self.n_vars}"
