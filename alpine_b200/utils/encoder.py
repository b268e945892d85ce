"""One-hot encoding of covariates; mirrors the reference's ``alpine/utils/encoder.py:11-60``.

Same class name, methods and attributes (``encoders``, ``encoded_labels``); NaN labels give all-zero rows
(encoder.py:27-37), categories unseen at ``fit_transform`` time are ignored at ``transform`` time
(``handle_unknown="ignore"``, encoder.py:23-25), column names are ``"<key>_<category>"`` as
``OneHotEncoder.get_feature_names_out`` produces them (encoder.py:36).  Host-side preprocessing: runs once per fit,
not part of the GPU hot path.  Written without scikit-learn so that the GPU box needs nothing beyond numpy/pandas.
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np
import numpy.typing as npt
import pandas as pd

Float32Array = npt.NDArray[np.float32]


def _missing(v) -> bool:
    return v is None or (isinstance(v, float) and v != v)


class _CategoryEncoder:
    """The part of sklearn's OneHotEncoder the reference uses: sorted categories, unknown -> zero row.  Labels are
    hashed once per column (``pandas.factorize``, C speed) and the handful of distinct values is mapped to the
    category indices, instead of one Python dict lookup per cell."""

    def __init__(self, key: str):
        self.key = key
        self.categories_: List = []

    def fit(self, values: np.ndarray) -> "_CategoryEncoder":
        self.fit_transform(values)
        return self

    def _set_categories(self, uniques) -> None:
        self.categories_ = sorted(v for v in uniques if not _missing(v))
        self._index = {c: i for i, c in enumerate(self.categories_)}

    def _one_hot(self, raw: np.ndarray, uniques) -> Float32Array:
        lut = np.fromiter((self._index.get(u, -1) for u in uniques), dtype=np.int64, count=len(uniques))
        cols = np.append(lut, -1)[raw]  # raw == -1 (missing) -> -1; unknown labels -> -1
        out = np.zeros((len(raw), len(self.categories_)), dtype=np.float32)
        rows = np.nonzero(cols >= 0)[0]
        out[rows, cols[rows]] = 1.0
        return out

    def fit_transform(self, values: np.ndarray) -> Float32Array:
        raw, uniques = pd.factorize(values)  # one hash pass; -1 for NaN / None
        uniques = uniques.tolist()
        self._set_categories(uniques)
        return self._one_hot(raw, uniques)

    def transform(self, values: np.ndarray) -> Float32Array:
        raw, uniques = pd.factorize(values)
        return self._one_hot(raw, uniques.tolist())

    def get_feature_names_out(self) -> np.ndarray:
        return np.asarray([f"{self.key}_{c}" for c in self.categories_], dtype=object)


class FeatureEncoders:
    def __init__(self, covariate_keys: List[str]):
        self.covariate_keys: List[str] = covariate_keys
        self.encoders: Dict[str, _CategoryEncoder] = {}
        self.encoded_labels: Dict[str, List[str]] = {}

    def fit_transform(self, df: pd.DataFrame) -> List[Float32Array]:
        if not isinstance(df, pd.DataFrame):
            raise TypeError("adata.obs must be a pandas DataFrame.")
        transformed_matrices: List[Float32Array] = []
        for key in self.covariate_keys:
            values = df[key].to_numpy()
            encoder = _CategoryEncoder(key)
            transformed = encoder.fit_transform(values)  # missing labels -> all-zero rows (encoder.py:27-37)
            self.encoders[key] = encoder
            self.encoded_labels[key] = encoder.get_feature_names_out().tolist()
            transformed_matrices.append(transformed)
        return transformed_matrices

    def transform(self, df: pd.DataFrame) -> List[Float32Array]:
        if not isinstance(df, pd.DataFrame):
            raise TypeError("adata.obs must be a pandas DataFrame.")
        transformed_matrices: List[Float32Array] = []
        for key in self.covariate_keys:
            if key in self.encoders:
                transformed_matrices.append(self.encoders[key].transform(df[key].to_numpy()))
        return transformed_matrices
