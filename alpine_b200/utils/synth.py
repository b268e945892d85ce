"""Seeded synthetic inputs of the shapes BASELINE.json names.

Used by tests, ``bench.py`` and ``__graft_entry__.smoke()``; there is no
network, so no real single-cell data set is available.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np


def make_counts(
    n_cells: int,
    n_genes: int,
    seed: int = 0,
    rank: int = 0,
    noise: float = 0.25,
    dtype=np.float32,
) -> np.ndarray:
    """cells x genes non-negative matrix (what ``adata.X`` holds).

    ``rank == 0``: i.i.d. Gamma(0.3, 2.0) (the survey's probe data).
    ``rank > 0``: low-rank Gamma factors plus Gamma noise, so that loss curves
    have an elbow and gene rankings have structure.
    """
    rng = np.random.default_rng(seed)
    if rank <= 0:
        return rng.gamma(0.3, 2.0, size=(n_cells, n_genes)).astype(dtype)
    Hc = rng.gamma(0.5, 1.0, size=(n_cells, rank))
    Wg = rng.gamma(0.4, 1.0, size=(rank, n_genes))
    X = Hc @ Wg + noise * rng.gamma(0.3, 2.0, size=(n_cells, n_genes))
    return X.astype(dtype)


def make_poisson_counts(n_cells: int, n_genes: int, seed: int = 0, rank: int = 8, depth: float = 3.0) -> np.ndarray:
    """cells x genes raw counts (uint16): Poisson draws around a low-rank Gamma mean, like UMI count matrices.
    Every value is an integer < 2048, i.e. exactly representable in tf32 (the count-matrix kernel variant)."""
    rng = np.random.default_rng(seed)
    Hc = rng.gamma(0.6, 1.0, size=(n_cells, rank))
    Wg = rng.gamma(0.4, 1.0, size=(rank, n_genes))
    mean = Hc @ Wg
    mean *= depth / mean.mean()
    return np.minimum(rng.poisson(mean), 2047).astype(np.uint16)


def make_labels(
    n_cells: int,
    n_categories: Sequence[int],
    seed: int = 0,
    nan_fraction: float = 0.0,
) -> List[np.ndarray]:
    """One object-dtype label vector per covariate (obs columns must be dtype 'O')."""
    rng = np.random.default_rng(seed + 1)
    out = []
    for ci, c in enumerate(n_categories):
        codes = rng.integers(0, c, size=n_cells)
        lab = np.array([f"c{ci}_{v}" for v in codes], dtype=object)
        if nan_fraction > 0:
            mask = rng.random(n_cells) < nan_fraction
            lab[mask] = np.nan
        out.append(lab)
    return out


def labels_to_dummies(labels: Sequence[np.ndarray], dtype=np.float32) -> Tuple[List[np.ndarray], List[List[str]]]:
    """cells x categories one-hot per covariate, NaN -> all-zero row.

    Mirrors the reference's FeatureEncoders.fit_transform (encoder.py:17-38)
    without sklearn: categories are the sorted unique non-null labels.
    """
    mats, names = [], []
    for lab in labels:
        lab = np.asarray(lab, dtype=object)
        isna = np.array([(v is None) or (isinstance(v, float) and v != v) for v in lab])
        cats = sorted(set(lab[~isna].tolist()))
        index = {c: i for i, c in enumerate(cats)}
        m = np.zeros((len(lab), len(cats)), dtype=dtype)
        rows = np.nonzero(~isna)[0]
        cols = np.fromiter((index[lab[r]] for r in rows), dtype=np.int64, count=len(rows))
        m[rows, cols] = 1
        mats.append(m)
        names.append([str(c) for c in cats])
    return mats, names
