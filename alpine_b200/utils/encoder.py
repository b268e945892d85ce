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


class _CategoryEncoder:
    """The part of sklearn's OneHotEncoder the reference uses: sorted categories, unknown -> zero row."""

    def __init__(self, key: str):
        self.key = key
        self.categories_: List = []

    def fit(self, values: np.ndarray) -> "_CategoryEncoder":
        self.categories_ = sorted(set(values.tolist()))
        self._index = {c: i for i, c in enumerate(self.categories_)}
        return self

    def transform(self, values: np.ndarray) -> Float32Array:
        out = np.zeros((len(values), len(self.categories_)), dtype=np.float32)
        cols = np.fromiter((self._index.get(v, -1) for v in values.tolist()), dtype=np.int64, count=len(values))
        rows = np.nonzero(cols >= 0)[0]
        out[rows, cols[rows]] = 1.0
        return out

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
            col = df[key]
            na_mask = col.isna().to_numpy()
            encoder = _CategoryEncoder(key).fit(col.to_numpy()[~na_mask])
            transformed = np.zeros((len(col), len(encoder.categories_)), dtype=np.float32)
            transformed[~na_mask, :] = encoder.transform(col.to_numpy()[~na_mask])
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
                col = df[key]
                na_mask = col.isna().to_numpy()
                encoder = self.encoders[key]
                transformed = np.zeros((len(col), len(encoder.categories_)), dtype=np.float32)
                transformed[~na_mask, :] = encoder.transform(col.to_numpy()[~na_mask])
                transformed_matrices.append(transformed)
        return transformed_matrices
