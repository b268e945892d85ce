"""AnnData access for the drop-in API.

The reference takes ``anndata.AnnData`` objects (``main.py:6, 84``).  ``anndata`` is not installed in the build
image nor on the GPU box and cannot be installed (no network), so when it is missing a minimal stand-in with the
attributes the reference's hot-path API touches is used instead: ``X``, ``obs``, ``var_names``, ``obs_names``,
``shape``, ``obsm``, ``varm``, ``layers``, and row subsetting ``adata[idx]`` (used by ComponentOptimizer's folds,
optimization.py:243-244).  With the real package installed, the real class is used and the stand-in is ignored.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import pandas as pd

try:  # pragma: no cover - not installed in this image
    import anndata as _ad

    AnnData = _ad.AnnData
    HAVE_ANNDATA = True
except Exception:  # ImportError or a broken install
    HAVE_ANNDATA = False

    class AnnData:  # type: ignore[no-redef]
        """Minimal stand-in: a cells x genes matrix with per-cell ``obs`` and per-gene ``var`` annotations."""

        def __init__(self, X, obs: Optional[pd.DataFrame] = None, var: Optional[pd.DataFrame] = None):
            self.X = X
            n_obs, n_var = X.shape
            self.obs = obs if obs is not None else pd.DataFrame(index=[str(i) for i in range(n_obs)])
            self.var = var if var is not None else pd.DataFrame(index=[str(i) for i in range(n_var)])
            if len(self.obs) != n_obs or len(self.var) != n_var:
                raise ValueError("obs/var length does not match X")
            self.obsm: dict = {}
            self.varm: dict = {}
            self.layers: dict = {}
            self.uns: dict = {}

        @property
        def shape(self):
            return self.X.shape

        @property
        def n_obs(self) -> int:
            return self.X.shape[0]

        @property
        def n_vars(self) -> int:
            return self.X.shape[1]

        @property
        def obs_names(self) -> pd.Index:
            return self.obs.index

        @property
        def var_names(self) -> pd.Index:
            return self.var.index

        def copy(self) -> "AnnData":
            X = self.X.copy() if is_sparse(self.X) else np.array(self.X, copy=True)
            out = AnnData(X, self.obs.copy(), self.var.copy())
            out.obsm = {k: np.array(v, copy=True) for k, v in self.obsm.items()}
            out.varm = {k: (v.copy() if hasattr(v, "copy") else v) for k, v in self.varm.items()}
            out.layers = {k: np.array(v, copy=True) for k, v in self.layers.items()}
            return out

        def __getitem__(self, idx) -> "AnnData":
            """Row (cell) subsetting, as ``adata[train_idx]`` in optimization.py:243-244."""
            if isinstance(idx, tuple):
                raise NotImplementedError("the AnnData stand-in only supports cell subsetting")
            idx = np.asarray(idx)
            if idx.dtype == bool:
                idx = np.nonzero(idx)[0]
            out = AnnData(self.X[idx], self.obs.iloc[idx].copy(), self.var.copy())
            out.obsm = {k: np.asarray(v)[idx] for k, v in self.obsm.items()}
            out.varm = dict(self.varm)
            out.layers = {k: np.asarray(v)[idx] for k, v in self.layers.items()}
            return out

        def __len__(self) -> int:
            return self.X.shape[0]

        def __repr__(self) -> str:
            return f"AnnData stand-in with n_obs x n_vars = {self.X.shape[0]} x {self.X.shape[1]}"


def is_sparse(x) -> bool:
    """scipy.sparse matrix / array?  Decided from the type's module so that dense fits never import scipy (0.2 s)."""
    if isinstance(x, np.ndarray):
        return False
    return (type(x).__module__ or "").startswith("scipy.sparse")
